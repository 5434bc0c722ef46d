# Round 2: smoke, BASELINE config 2 with every interval compared (100 Mbp, 1 M reads), repeat-rich robustness run (400 Mbp, 30 % repeat
# families), drop-in throughput through the reference's own bwa mem (tools/dropin_bench.py)
set -x
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --ref-bp 100000000 --reads 1000000 --full-compare --no-extras > gpurun_out/r2_config2_full_compare.json 2> gpurun_out/r2_config2.err; echo cfg2 rc=$?
grep "full compare" gpurun_out/r2_config2.err
timeout 1200 python bench.py --ref-bp 400000000 --reads 1000000 --repeat-frac 0.3 --no-extras > gpurun_out/r2_repeat_rich_400Mbp.json 2> gpurun_out/r2_repeat.err; echo repeat rc=$?
timeout 1200 python tools/dropin_bench.py --ref-bp 100000000 --reads 2000000 --threads 16 --batches 64,256,1024,16384,65536 --handles 0 --out gpurun_out/r2_dropin_100Mbp.json > /dev/null 2> gpurun_out/r2_dropin_100Mbp.err
echo dropin rc=$?
grep "^\[dropin\]" gpurun_out/r2_dropin_100Mbp.err | cut -c1-420
