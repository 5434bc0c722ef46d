set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo pytest rc=$?
tail -8 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_n1.json')); print('value %.1f M  kernel %.2f ms  e2e %.1f M (%.2f ms)' % (d['value']/1e6, d['roofline']['kernel_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']), d['parity'], d['roofline']['frac'], d['config4_250bp']['value'], d['e2e_seeds']['value'], d['e2e_chains']['value'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_kernel -s 5 -c 1 -o gpurun_out/r2_seed python bench.py --skip-cpu --no-extras --steps 1 --warmup 3 > gpurun_out/r2_ncu_seed.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/r2_seed.ncu-rep --page raw --csv > gpurun_out/r2_seed_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_seed.ncu-rep --page source --csv > gpurun_out/r2_seed_source.csv 2>/dev/null
rm -f gpurun_out/r2_seed.ncu-rep
