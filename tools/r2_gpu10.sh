set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_lpr1.log 2>&1
echo pytest rc=$?
tail -30 gpurun_out/r2_pytest_lpr1.log
timeout 600 python bench.py --no-extras --skip-cpu --steps 10 > gpurun_out/r2c_bench_lpr1.json 2> gpurun_out/r2c_bench_lpr1.err
echo bench rc=$?
tail -3 gpurun_out/r2c_bench_lpr1.err | cut -c1-400
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_lpr1.json')); print('LPR1 value %.1f M  kernel %.2f ms  e2e %.1f M' % (d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6), d['parity'])"
