set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo pytest rc=$?
tail -12 gpurun_out/r2_pytest_gpu.log
timeout 900 python tools/dropin_bench.py --ref-bp 100000000 --reads 2000000 --threads 16 --handles 8 --out gpurun_out/r2_dropin_100Mbp.json > /dev/null 2> gpurun_out/r2_dropin_100Mbp.err
echo dropin rc=$?
grep "^\[dropin\]" gpurun_out/r2_dropin_100Mbp.err | cut -c1-700
