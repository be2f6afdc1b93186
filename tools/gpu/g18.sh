# 2-GPU parity tests (fixed partitions, adaptive replicated layout incl. the CUSTOMPROP callback case)
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_dist.py -x -q -s > gpurun_out/r2_pytest_dist_2gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dist_2gpu_b.log
grep -E "CUSTOMPROP|passed|failed|OK|FAILED|rc=" gpurun_out/r2_pytest_dist_2gpu_b.log | tail -20
