# concurrent parameter sweeps on one GPU: spinning vs blocking host waits
set -x
mkdir -p gpurun_out
for mode in spin block; do
SWEEP_SYNC=$mode timeout 900 python tools/sweep_throughput.py toggle 64 100 > gpurun_out/r2_sweep_throughput_$mode.txt 2>&1
SWEEP_SYNC=$mode timeout 600 python tools/sweep_throughput.py repressilator 64 2 >> gpurun_out/r2_sweep_throughput_$mode.txt 2>&1
done
cat gpurun_out/r2_sweep_throughput_spin.txt gpurun_out/r2_sweep_throughput_block.txt
