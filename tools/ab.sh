# A/B helper for gpurun: run <tag> VAR=value ... -> one line with the per-class kernel times of a short bench run
run() { tag=$1; shift; env "$@" python bench.py --steps 2 --warmup 2 --e2e-steps 0 --no-cpu-baseline --no-companion --no-parity > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$tag.json"))
    k=d["kernels"]
    print("$tag", "ms/solve %.2f"%d["ms_per_step"], " ".join("%s=%.4f"%(n,k[n]["avg_ms"]) for n in ("spmv_dot","spmv_fin_dot","spmv_fin_nrm","axpy_dot","combine") if n in k), "plain=%.4f"%d["roofline"]["plain_spmv"]["avg_ms"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag FAILED", e)
PY
}
dram() { tag=$1; shift; env "$@" ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_spmv_bd2 -s 15 -c 1 --csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-companion --no-parity 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F, -v t=$tag '{print t, $(NF-2), $(NF-1), $NF}'; }
