# Round 2, third GPU pass: GPU tests, drop-in throughput with the leader breakdown, kernel variants, ncu capture of the seed kernel.
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo pytest rc=$?
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 900 python tools/dropin_bench.py --ref-bp 100000000 --reads 1000000 --threads 16 --handles 4,8 --out gpurun_out/r2_dropin_100Mbp.json > /dev/null 2> gpurun_out/r2_dropin_100Mbp.err
echo dropin rc=$?
grep "^\[dropin\]" gpurun_out/r2_dropin_100Mbp.err | cut -c1-700
bash tools/variants.sh run > gpurun_out/r2_variants.txt 2>&1
cat gpurun_out/r2_variants.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_kernel -s 2 -c 1 -o gpurun_out/r2_seed python bench.py --skip-cpu --no-extras --steps 1 --warmup 3 > gpurun_out/r2_ncu_seed.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/r2_seed.ncu-rep --page raw --csv > gpurun_out/r2_seed_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_seed.ncu-rep --page source --csv > gpurun_out/r2_seed_source.csv 2>/dev/null
ls -la gpurun_out/
