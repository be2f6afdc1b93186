# CUSTOMPROP probing + shim call sequence from a compiled C host: whole GPU suite
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_g.log
tail -15 gpurun_out/r2_pytest_gpu_g.log
