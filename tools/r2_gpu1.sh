# Round 2, first GPU pass on one B200 (run under gpurun): GPU test suite, then the default bench line.
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo pytest rc=$?
tail -15 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo bench rc=$?
tail -30 gpurun_out/r2_bench_n1.err
tail -c 1500 gpurun_out/r2_bench_n1.json
