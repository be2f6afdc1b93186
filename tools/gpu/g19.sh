# compute-sanitizer memcheck over the fast GPU tests (state space, golden fixtures, kernels, customprop, lattice)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_golden.py tests/test_gpu_state_space.py tests/test_gpu_kernels.py tests/test_gpu_lattice.py -m gpu -x -q > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
tail -25 gpurun_out/r2_memcheck.log
