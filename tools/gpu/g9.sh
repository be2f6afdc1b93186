# register-window LU / substitution of the Pade kernel (debug build with cycle stamps): parity, then the stamps
set -x
mkdir -p gpurun_out
export KFSP_LIB=$PWD/krylovfspssa_b200/libkfsp_dbg.so
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_full_configs.py -x -q 2>&1 | grep -v "^expm n=" | tail -5
timeout 300 python tools/expm_timing.py 2>&1 | grep -E "cycles|per call" | awk "NR%31<2" | head -30
