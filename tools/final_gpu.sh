set -x
python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err
echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_bench_ref.json 2> gpurun_out/r1f_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --skip-cpu --steps 2 --warmup 3 > gpurun_out/r1f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seed_kernel -s 2 -c 1 -o gpurun_out/r1f_seed_uw python bench.py --skip-cpu --steps 1 --warmup 3 > gpurun_out/r1f_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
