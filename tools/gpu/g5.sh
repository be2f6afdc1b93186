# cluster Pade kernel + merged reductions of the single-CTA sweep: parity tests, then phase times of configs 1-3
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_full_configs.py tests/test_gpu_state_space.py -x -q > gpurun_out/r2_pytest_expm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_expm.log
timeout 600 python tools/phase_breakdown.py toggle repressilator goutsias > gpurun_out/r2_phases_cluster_expm.txt 2>&1
KFSP_EXPM_CLUSTER=1000 timeout 600 python tools/phase_breakdown.py toggle repressilator > gpurun_out/r2_phases_single_cta_expm.txt 2>&1
tail -5 gpurun_out/r2_pytest_expm.log; cat gpurun_out/r2_phases_cluster_expm.txt gpurun_out/r2_phases_single_cta_expm.txt
