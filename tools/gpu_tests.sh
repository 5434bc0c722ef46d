set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo pytest rc=$?
tail -15 gpurun_out/r2_pytest_gpu.log
