# index-only SpMV (spmv_variant 2): parity tests, config-5 bench in natural and scattered order, configs 1-3 phase times
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_index_only.py -x -q -s > gpurun_out/r2_pytest_idx.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_idx.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-companion --no-parity"
timeout 600 $B --spmv-variant 2 > gpurun_out/r2_bench_n1_idx.json 2> gpurun_out/r2_bench_n1_idx.err
KFSP_IDX_DREC=1 timeout 600 $B --spmv-variant 2 > gpurun_out/r2_bench_n1_idx_drec.json 2> gpurun_out/r2_bench_n1_idx_drec.err
timeout 900 $B --spmv-variant 2 --scattered > gpurun_out/r2_bench_n1_idx_scattered.json 2> gpurun_out/r2_bench_n1_idx_scattered.err
timeout 900 $B --spmv-variant 0 --scattered > gpurun_out/r2_bench_n1_explicit_scattered.json 2> gpurun_out/r2_bench_n1_explicit_scattered.err
for v in 0 2; do KFSP_VARIANT=$v timeout 600 python tools/phase_breakdown.py goutsias repressilator toggle > gpurun_out/r2_phases_v$v.txt 2>&1; done
KFSP_VARIANT=2 KFSP_IDX_DREC=1 timeout 600 python tools/phase_breakdown.py goutsias > gpurun_out/r2_phases_v2_drec.txt 2>&1
ls -la gpurun_out
