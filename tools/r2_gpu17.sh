set -x
timeout 1500 python -m pytest tests/test_sa.py tests/test_chains.py tests/test_gpu_config1.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest_sa.log 2>&1
echo pytest rc=$?
tail -5 gpurun_out/r2_pytest_sa.log
timeout 900 python bench.py --skip-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo bench rc=$?
grep -E "^\[bench\] e2e, (seeds|chains)" gpurun_out/r2d_bench.err | cut -c1-400
