# Round-end measurement pass on one B200 (run under gpurun): configs 4 and 2, the default bench line, the reference arm,
# the launch list of our kernels.  Outputs under gpurun_out/.
set -x
python bench.py --ref-bp 3100000000 --reads 1000000 --read-len 250 --err 0.02 --steps 5 > gpurun_out/r1f_cfg4.json 2> gpurun_out/r1f_cfg4.err
python bench.py --ref-bp 100000000 --reads 1000000 --full-compare --steps 5 > gpurun_out/r1f_cfg2.json 2> gpurun_out/r1f_cfg2.err
python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err
echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:seed_kernel|pack_|compact|scan_|rf_|repack|fsa_build|gather_probe|scatter_counts' -c 400 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --skip-cpu --steps 5 --warmup 3 > gpurun_out/r1f_ncu_launch.log 2>&1
tail -c 600 gpurun_out/r1f_bench_n1.json
