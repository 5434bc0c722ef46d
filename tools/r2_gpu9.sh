set -x
timeout 900 python -m pytest tests/test_gpu_wire.py -m gpu -x -q > gpurun_out/r2_pytest_wire.log 2>&1
echo pytest rc=$?
tail -30 gpurun_out/r2_pytest_wire.log
timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
echo bench rc=$?
tail -5 gpurun_out/r2b_bench_n1.err | cut -c1-600
