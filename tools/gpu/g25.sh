# CUSTOMPROP with bilinear reactions recognised by probing (transcr6d on the device): parity + config 4 phase times
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_l.log
tail -5 gpurun_out/r2_pytest_gpu_l.log
timeout 900 python tools/phase_breakdown.py transcr6d goutsias > gpurun_out/r2_phases_probe1.txt 2>&1
KFSP_CUSTOM_PROBE=0 timeout 900 python tools/phase_breakdown.py transcr6d > gpurun_out/r2_phases_probe0.txt 2>&1
grep -v "expm n=" gpurun_out/r2_phases_probe1.txt gpurun_out/r2_phases_probe0.txt
