# ncu --set full of one late SSA counting pass of the Goutsias solve (source-level stall sampling)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ssa_walk -s 200 -c 1 -o /tmp/r2_ssa_full2 python tools/phase_breakdown.py goutsias > gpurun_out/ncu_ssa.log 2>&1
ncu -i /tmp/r2_ssa_full2.ncu-rep --page raw --csv > gpurun_out/r2_ssa_walk2_raw.csv 2>/dev/null
ncu -i /tmp/r2_ssa_full2.ncu-rep --page details > gpurun_out/r2_ssa_walk2_full.txt 2>/dev/null
ncu -i /tmp/r2_ssa_full2.ncu-rep --page source --csv > gpurun_out/r2_ssa_walk2_source.csv 2>/dev/null
ls -la gpurun_out/r2_ssa_walk2*; tail -3 gpurun_out/ncu_ssa.log
