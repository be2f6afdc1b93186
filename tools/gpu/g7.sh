# whole GPU suite on the current build, Pade kernel timing, phase times of configs 1-3
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_b.log
timeout 300 python tools/expm_timing.py > gpurun_out/r2_expm_timing_cluster.txt 2>&1
timeout 600 python tools/phase_breakdown.py toggle repressilator goutsias > gpurun_out/r2_phases_b.txt 2>&1
tail -4 gpurun_out/r2_pytest_gpu_b.log; cat gpurun_out/r2_expm_timing_cluster.txt gpurun_out/r2_phases_b.txt
