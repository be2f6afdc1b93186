# where does kfsp_solve spend its time beyond the resident solve on a partitioned handle? (2 GPUs)
set -x
mkdir -p gpurun_out
KFSP_DEBUG_E2E=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-companion --no-parity > gpurun_out/r2_e2e_dbg_n2.json 2> gpurun_out/r2_e2e_dbg_n2.err
grep kfsp_solve gpurun_out/r2_e2e_dbg_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_e2e_dbg_n2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("dist"))
PY
