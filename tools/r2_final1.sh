# Round 2, single-GPU evidence run: GPU tests, the bench line, the reference arm, launch list of a bench step, full ncu capture of the seed kernel.
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo pytest rc=$?
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_arm.json 2> gpurun_out/r2_bench_ref_arm.err
echo ref arm rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:seed_kernel|pack_|unpack2|window_flags|amb_patch|compact|scan_|publish|tiny_tail|intv1|seed_expand|seed_count|chain' -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --skip-cpu --no-extras --steps 3 --warmup 3 > gpurun_out/r2_ncu_launch.log 2>&1
echo ncu launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_kernel -s 5 -c 1 -o gpurun_out/r2_seed python bench.py --skip-cpu --no-extras --steps 1 --warmup 3 > gpurun_out/r2_ncu_seed.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/r2_seed.ncu-rep --page raw --csv > gpurun_out/r2_seed_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_seed.ncu-rep --page source --csv > gpurun_out/r2_seed_source.csv 2>/dev/null
rm -f gpurun_out/r2_seed.ncu-rep
