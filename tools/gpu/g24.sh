# SSA walks evaluate out-of-projection states through the factored tables: parity + phase times (A/B KFSP_SSA_FAC)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_k.log
tail -5 gpurun_out/r2_pytest_gpu_k.log
timeout 600 python tools/phase_breakdown.py goutsias repressilator toggle > gpurun_out/r2_phases_fac1.txt 2>&1
KFSP_SSA_FAC=0 timeout 600 python tools/phase_breakdown.py goutsias > gpurun_out/r2_phases_fac0.txt 2>&1
grep -v "expm n=" gpurun_out/r2_phases_fac1.txt gpurun_out/r2_phases_fac0.txt
