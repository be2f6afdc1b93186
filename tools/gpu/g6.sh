# Pade kernel timing by phase (ns = 0: everything but the squarings), CUSTOMPROP configs with the persistent side cache
set -x
mkdir -p gpurun_out
timeout 300 python tools/expm_timing.py > gpurun_out/r2_expm_timing_cluster.txt 2>&1
KFSP_EXPM_CLUSTER=1000 timeout 300 python tools/expm_timing.py > gpurun_out/r2_expm_timing_single.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_customprop.py tests/test_gpu_full_configs.py -x -q -s > gpurun_out/r2_pytest_customprop.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_customprop.log
timeout 600 python tools/phase_breakdown.py transcr6d driver_repressilator driver_toggle > gpurun_out/r2_phases_customprop.txt 2>&1
KFSP_PROP_CACHE_STATES=0 timeout 600 python tools/phase_breakdown.py transcr6d > gpurun_out/r2_phases_customprop_nocache.txt 2>&1
cat gpurun_out/r2_expm_timing_cluster.txt gpurun_out/r2_expm_timing_single.txt gpurun_out/r2_phases_customprop.txt gpurun_out/r2_phases_customprop_nocache.txt; tail -12 gpurun_out/r2_pytest_customprop.log
