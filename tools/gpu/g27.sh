# SSA walk with the reaction count at compile time: parity + phase times
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_m.log
tail -5 gpurun_out/r2_pytest_gpu_m.log
timeout 900 python tools/phase_breakdown.py goutsias repressilator toggle transcr6d > gpurun_out/r2_phases_rt.txt 2>&1
grep -v "expm n=" gpurun_out/r2_phases_rt.txt
