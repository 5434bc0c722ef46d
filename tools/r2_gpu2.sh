# Round 2, second GPU pass: rest of the GPU tests, launch list of a bench step, drop-in throughput (config 2 size).
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo pytest rc=$?
tail -8 gpurun_out/r2_pytest_gpu.log
timeout 900 python tools/dropin_bench.py --ref-bp 100000000 --reads 1000000 --out gpurun_out/r2_dropin_100Mbp.json > /dev/null 2> gpurun_out/r2_dropin_100Mbp.err
echo dropin rc=$?
grep "^\[dropin\]" gpurun_out/r2_dropin_100Mbp.err | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:seed_kernel|pack_|unpack2|window_flags|amb_patch|compact|scan_|scatter_counts|intv16' -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --skip-cpu --no-extras --steps 3 --warmup 3 > gpurun_out/r2_ncu_launch.log 2>&1
echo ncu rc=$?
tail -40 gpurun_out/r2_launches.csv | cut -c1-200
