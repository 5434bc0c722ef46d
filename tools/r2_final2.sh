# Round 2: smoke, BASELINE config 2 with every interval compared (100 Mbp, 1 M reads), repeat-rich robustness run (400 Mbp, 30 % repeat families)
set -x
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --ref-bp 100000000 --reads 1000000 --full-compare --no-extras > gpurun_out/r2_config2_full_compare.json 2> gpurun_out/r2_config2.err; echo cfg2 rc=$?
grep "full compare" gpurun_out/r2_config2.err
timeout 1200 python bench.py --ref-bp 400000000 --reads 1000000 --repeat-frac 0.3 --no-extras > gpurun_out/r2_repeat_rich_400Mbp.json 2> gpurun_out/r2_repeat.err; echo repeat rc=$?
tail -3 gpurun_out/r2_repeat.err | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo bench rc=$?
