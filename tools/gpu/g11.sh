# state-major SSA records: parity (whole suite), SSA phase of configs 3-4 with and without
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_d.log
timeout 600 python tools/phase_breakdown.py goutsias repressilator transcr6d > gpurun_out/r2_phases_d.txt 2>&1
KFSP_SSA_REC=0 timeout 600 python tools/phase_breakdown.py goutsias > gpurun_out/r2_phases_d_norec.txt 2>&1
tail -4 gpurun_out/r2_pytest_gpu_d.log; grep -v "expm n=" gpurun_out/r2_phases_d.txt gpurun_out/r2_phases_d_norec.txt
