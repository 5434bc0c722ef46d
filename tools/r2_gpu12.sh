set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_kernel -s 5 -c 1 -o gpurun_out/r2_seed_lpr1 python bench.py --skip-cpu --no-extras --steps 1 --warmup 3 > gpurun_out/r2_ncu_seed_lpr1.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/r2_seed_lpr1.ncu-rep --page raw --csv > gpurun_out/r2_seed_lpr1_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_seed_lpr1.ncu-rep --page source --csv > gpurun_out/r2_seed_lpr1_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
