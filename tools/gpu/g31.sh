# A/B in one run: tables staged before griddepcontrol.wait (new) vs after (old library), 1 GPU, alternating
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-companion --no-parity"
for i in 1 2; do
timeout 300 $CMD > gpurun_out/ab_new$i.json 2>/dev/null
KFSP_LIB=$PWD/krylovfspssa_b200/libkfsp_old.so timeout 300 $CMD > gpurun_out/ab_old$i.json 2>/dev/null
done
python - <<'PY'
import json
for f in ("new1","old1","new2","old2"):
    d=json.loads(open("gpurun_out/ab_%s.json"%f).read().strip().splitlines()[-1])
    print(f, "ms %.2f"%d["ms_per_step"], "sweep avg %.4f"%d["roofline"]["avg_launch_ms"], {k:round(v["avg_ms"],4) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"])
PY
timeout 600 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q 2>&1 | tail -2
