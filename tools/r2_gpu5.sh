set -x
timeout 900 python tools/dropin_bench.py --ref-bp 100000000 --reads 2000000 --threads 16 --handles 8 --out gpurun_out/r2_dropin_100Mbp.json > /dev/null 2> gpurun_out/r2_dropin_100Mbp.err
echo dropin rc=$?
grep "^\[dropin\]" gpurun_out/r2_dropin_100Mbp.err | cut -c1-700
timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -3
