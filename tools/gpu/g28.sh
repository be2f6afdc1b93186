# final round-2 evidence on one GPU: the driver's commands (pytest -m gpu, smoke, bench.py) + the ncu launch list of the bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-companion --no-parity"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_1gpu_final.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu_final.log gpurun_out/r2_smoke.log; cut -c1-600 gpurun_out/r2_bench_n1_final.json; cut -c1-400 gpurun_out/r2_bench_reference_final.json
