set -x
mkdir -p gpurun_out
for v in 0 1 0 1; do
  KFSP_EXPM_SMALL_THREADS=$v timeout 300 python tools/phase_breakdown.py toggle repressilator goutsias 2>&1 | sed "s/^/small=$v /" | cut -c1-330
done | tee gpurun_out/r2_expm_small_threads_ab.txt
for v in 0 1; do KFSP_EXPM_SMALL_THREADS=$v timeout 200 python tools/expm_timing.py 2>&1 | grep "n= 32" | sed "s/^/small=$v /"; done | tee -a gpurun_out/r2_expm_small_threads_ab.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu_n.log
