set -x
timeout 1500 python -m pytest tests/test_gpu_sector_kernels.py -m gpu -x -q > gpurun_out/r2_pytest_split.log 2>&1
echo pytest rc=$?
tail -15 gpurun_out/r2_pytest_split.log
for v in 4 2; do
timeout 600 python bench.py --no-extras --skip-cpu --steps 10 --set lanes_per_read=$v > gpurun_out/ab_lpr$v.json 2> gpurun_out/ab_lpr$v.err
echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/ab_lpr$v.json')); print('lanes_per_read=$v: %.1f M reads/s, seed kernel %.2f ms' % (d['value']/1e6, d['roofline']['kernel_ms']), d['parity']['bit_exact'])"
done
