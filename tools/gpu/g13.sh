# SSA candidates emitted in the counting pass: parity (whole suite), phase times with and without
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_e.log
timeout 600 python tools/phase_breakdown.py goutsias repressilator toggle > gpurun_out/r2_phases_f.txt 2>&1
KFSP_SSA_EMIT=0 timeout 600 python tools/phase_breakdown.py goutsias > gpurun_out/r2_phases_f_replay.txt 2>&1
tail -4 gpurun_out/r2_pytest_gpu_f.log; grep -v "expm n=" gpurun_out/r2_phases_f.txt gpurun_out/r2_phases_f_replay.txt
