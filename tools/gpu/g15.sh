# where do the ~60 us per column launch go between the per-launch kernel times and the in-situ sweep time?  PDL on/off A/B
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-companion --no-parity"
timeout 300 $CMD > gpurun_out/r2_ab_pdl1.json 2> gpurun_out/r2_ab_pdl1.err
KFSP_PDL=0 timeout 300 $CMD > gpurun_out/r2_ab_pdl0.json 2> gpurun_out/r2_ab_pdl0.err
timeout 300 $CMD > gpurun_out/r2_ab_pdl1b.json 2> gpurun_out/r2_ab_pdl1b.err
python - <<'PY'
import json
for f in ("pdl1","pdl0","pdl1b"):
    try:
        d=json.loads(open("gpurun_out/r2_ab_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms_per_step %.2f"%d["ms_per_step"], "sweep avg launch %.4f"%d["roofline"]["avg_launch_ms"], {k:round(v["avg_ms"],4) for k,v in d["kernels"].items()}, d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
