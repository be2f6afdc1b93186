# low-latency reduction exchange: 2-GPU parity (dist tests) + bench with exchange statistics, LL on / off
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -s > gpurun_out/r2_pytest_dist_2gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dist_2gpu_c.log
grep -E "passed|failed|OK|FAILED|rc=" gpurun_out/r2_pytest_dist_2gpu_c.log | tail -6
for ll in 1 0; do
KFSP_DIST_LL=$ll timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$ll bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-companion --e2e-steps 0 > gpurun_out/r2_bench_n2_ll$ll.json 2> gpurun_out/r2_bench_n2_ll$ll.err
done
python - <<'PY'
import json
for ll in (1,0):
    d=json.loads(open("gpurun_out/r2_bench_n2_ll%d.json"%ll).read().strip().splitlines()[-1])
    print("LL",ll,"ms",d["ms_per_step"],"exchange_us",d["dist"].get("exchange_us"),d["dist"].get("exchange_us_worst_single"),{k:round(v["avg_ms"],4) for k,v in d["kernels"].items()}, d.get("dist_parity",{}).get("dist_bit_identical"))
PY
