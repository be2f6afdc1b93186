# round-2 evidence run: GPU tests, the bench line, the ncu launch list and one full capture of the column kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-companion --no-parity"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_1gpu.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_bd2 -s 15 -c 1 -o /tmp/r2_bd2_full $CMD > gpurun_out/ncu2.log 2>&1
ncu -i /tmp/r2_bd2_full.ncu-rep --page raw --csv > gpurun_out/r2_bd2_full_raw.csv 2>/dev/null
ncu -i /tmp/r2_bd2_full.ncu-rep --page details > gpurun_out/r2_bd2_full_details.txt 2>/dev/null
ncu -i /tmp/r2_bd2_full.ncu-rep --page source --csv > gpurun_out/r2_bd2_full_source.csv 2>/dev/null
gzip -c /tmp/r2_bd2_full.ncu-rep > /tmp/rep.gz; ls -la /tmp/rep.gz; [ $(stat -c %s /tmp/rep.gz) -lt 40000000 ] && cp /tmp/rep.gz gpurun_out/r2_bd2_full.ncu-rep.gz
du -sh gpurun_out; ls -la gpurun_out
