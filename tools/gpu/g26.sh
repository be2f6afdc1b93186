# Goutsias / transcr6d full horizon: explicit vs index-only SpMV with the cooperative sweep
set -x
mkdir -p gpurun_out
KFSP_VARIANT=2 timeout 900 python tools/phase_breakdown.py goutsias repressilator > gpurun_out/r2_phases_idx_coop.txt 2>&1
KFSP_VARIANT=2 KFSP_COOP_SWEEP=0 timeout 900 python tools/phase_breakdown.py goutsias > gpurun_out/r2_phases_idx_nocoop.txt 2>&1
grep -v "expm n=" gpurun_out/r2_phases_idx_coop.txt gpurun_out/r2_phases_idx_nocoop.txt
