# ncu evidence for the other kernels of round 2: index-only SpMV (full capture), Goutsias launch list (large-N end of the solve),
# cluster Pade kernel (full capture).  Each ncu run only after the same command exited 0 without ncu.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
CMD="python bench.py --spmv-variant 2 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity"
timeout 300 $CMD > gpurun_out/plain_idx.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_idx -s 12 -c 1 -o /tmp/r2_idx_full $CMD > gpurun_out/ncu_idx.log 2>&1
ncu -i /tmp/r2_idx_full.ncu-rep --page raw --csv > gpurun_out/r2_spmv_idx_full_raw.csv 2>/dev/null
ncu -i /tmp/r2_idx_full.ncu-rep --page details > gpurun_out/r2_spmv_idx_full.txt 2>/dev/null
CMD2="python tools/phase_breakdown.py goutsias"
timeout 300 $CMD2 > gpurun_out/plain_goutsias.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80000 -c 8000 --csv --log-file gpurun_out/r2_launches_goutsias_tail.csv $CMD2 > gpurun_out/ncu_goutsias.log 2>&1
CMD3="python tools/phase_breakdown.py toggle"
timeout 300 $CMD3 > gpurun_out/plain_toggle.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_expm_cluster -s 400 -c 1 -o /tmp/r2_expm_full $CMD3 > gpurun_out/ncu_expm.log 2>&1
ncu -i /tmp/r2_expm_full.ncu-rep --page raw --csv > gpurun_out/r2_expm_cluster_full_raw.csv 2>/dev/null
ncu -i /tmp/r2_expm_full.ncu-rep --page details > gpurun_out/r2_expm_cluster_full.txt 2>/dev/null
tail -3 gpurun_out/r2_pytest_gpu_final.log; ls -la gpurun_out | tail -12
