# 2 GPUs: fixed-set and ADAPTIVE partitioned solves (bit-identity), full-horizon configs 1-3 on 2 GPUs, bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_gpu_dist.py -x -q -s > gpurun_out/r2_pytest_dist2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dist2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_adaptive_check.py full > gpurun_out/r2_dist_adaptive_full_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_adaptive_full_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_pytest_dist2.log gpurun_out/r2_dist_adaptive_full_2gpu.log
