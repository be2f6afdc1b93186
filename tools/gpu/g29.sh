# 8 GPUs at HEAD: bench (lattice-vs-explicit parity at 1e8 states, partitioned-vs-solo check and exchange statistics inside)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
cut -c1-300 gpurun_out/r2_bench_n8_final.json
