# shared-memory-resident sweep + new Pade kernel: parity (whole suite), phase times of configs 1-3 with and without it
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_c.log
timeout 600 python tools/phase_breakdown.py toggle repressilator goutsias driver_toggle > gpurun_out/r2_phases_c.txt 2>&1
KFSP_SMEM_SWEEP=0 timeout 600 python tools/phase_breakdown.py toggle driver_toggle > gpurun_out/r2_phases_c_nosmem.txt 2>&1
tail -4 gpurun_out/r2_pytest_gpu_c.log; grep -v "expm n=" gpurun_out/r2_phases_c.txt gpurun_out/r2_phases_c_nosmem.txt
