set -x
timeout 1200 python tools/dropin_bench.py --ref-bp 100000000 --reads 1000000 --threads 16 --batches 64,256,1024 --handles 0 --out gpurun_out/r2_dropin_small.json > /dev/null 2> gpurun_out/r2_dropin_small.err
echo dropin rc=$?
grep "^\[dropin\]" gpurun_out/r2_dropin_small.err | cut -c1-900
