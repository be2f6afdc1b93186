# SSA walk as one flattened jump loop per lane: parity + phase times (emit and replay paths, CUSTOMPROP config 4)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_j.log
tail -5 gpurun_out/r2_pytest_gpu_j.log
timeout 600 python tools/phase_breakdown.py goutsias repressilator > gpurun_out/r2_phases_flat.txt 2>&1
KFSP_SSA_EMIT=0 timeout 600 python tools/phase_breakdown.py goutsias > gpurun_out/r2_phases_flat_replay.txt 2>&1
timeout 900 python tools/phase_breakdown.py transcr6d > gpurun_out/r2_phases_flat_transcr6d.txt 2>&1
grep -v "expm n=" gpurun_out/r2_phases_flat.txt gpurun_out/r2_phases_flat_replay.txt gpurun_out/r2_phases_flat_transcr6d.txt
