set -x
timeout 900 python -m pytest tests/test_gpu_wire.py -m gpu -x -q > gpurun_out/r2_pytest_wire.log 2>&1
echo pytest rc=$?
tail -12 gpurun_out/r2_pytest_wire.log
timeout 900 python bench.py > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo bench rc=$?
tail -3 gpurun_out/r2g_bench_n1.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_n1.json')); print('value %.1f M kernel %.2f ms e2e %.1f M (%.2f ms) d2h %d' % (d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step']), d['parity'], d['e2e']['record_choice'], d['e2e']['api'])"
