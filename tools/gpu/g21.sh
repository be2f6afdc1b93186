# OFFDIAG as one record per state (SSA walks / coef gather read all R values of a random state): parity + phase times
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_i.log
tail -5 gpurun_out/r2_pytest_gpu_i.log
timeout 600 python tools/phase_breakdown.py goutsias repressilator toggle > gpurun_out/r2_phases_aos.txt 2>&1
grep -v "expm n=" gpurun_out/r2_phases_aos.txt
