# cooperative mid-size sweep: parity (whole GPU suite) + phase times with and without
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_h.log
tail -5 gpurun_out/r2_pytest_gpu_h.log
timeout 600 python tools/phase_breakdown.py goutsias repressilator > gpurun_out/r2_phases_coop1.txt 2>&1
KFSP_COOP_SWEEP=0 timeout 600 python tools/phase_breakdown.py goutsias repressilator > gpurun_out/r2_phases_coop0.txt 2>&1
grep -v "expm n=" gpurun_out/r2_phases_coop1.txt gpurun_out/r2_phases_coop0.txt
