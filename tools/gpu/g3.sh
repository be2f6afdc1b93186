# index-only SpMV after the load-grouping change + launch list of the latency-bound toggle solve (config 1)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_index_only.py tests/test_gpu_kernels.py -x -q > gpurun_out/r2_pytest_idx2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_idx2.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-companion --no-parity"
timeout 600 $B --spmv-variant 2 > gpurun_out/r2_bench_n1_idx.json 2> gpurun_out/r2_bench_n1_idx.err
KFSP_IDX_DREC=1 timeout 600 $B --spmv-variant 2 > gpurun_out/r2_bench_n1_idx_drec.json 2> gpurun_out/r2_bench_n1_idx_drec.err
for v in 0 2; do KFSP_VARIANT=$v timeout 600 python tools/phase_breakdown.py goutsias > gpurun_out/r2_phases_goutsias_v$v.txt 2>&1; done
timeout 300 python tools/phase_breakdown.py toggle > gpurun_out/plain_toggle.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r2_launches_toggle_full.csv python tools/phase_breakdown.py toggle > gpurun_out/ncu_toggle.log 2>&1
ls -la gpurun_out
