# 8 GPUs: bench (with lattice-vs-explicit parity at 1e8 states and the partitioned-vs-solo check inside), adaptive check at 8 ranks
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
KFSP_REPL_MIN_ROWS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tests/dist_adaptive_check.py > gpurun_out/r2_dist_adaptive_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_adaptive_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 10 --warmup 3 --no-parity --no-companion > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
grep -c "bit-identical=True" gpurun_out/r2_dist_adaptive_8gpu.log; tail -2 gpurun_out/r2_dist_adaptive_8gpu.log
