set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_lpr1.log 2>&1
echo pytest rc=$?
tail -5 gpurun_out/r2_pytest_lpr1.log
for v in 1 2; do
timeout 600 python bench.py --no-extras --skip-cpu --steps 10 --set lanes_per_read=$v > gpurun_out/r2c_bench_lpr$v.json 2> gpurun_out/r2c_bench_lpr$v.err
echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_lpr$v.json')); print('LPR$v value %.1f M  kernel %.2f ms  e2e %.1f M' % (d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6), d['parity'])"
done
