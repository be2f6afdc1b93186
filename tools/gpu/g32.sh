# L2 eviction hints for the generator in the cooperative sweep: Goutsias / transcr6d phase times for several keep sizes
set -x
mkdir -p gpurun_out
python -c "
import ctypes
rt=ctypes.CDLL('libcudart.so')
v=ctypes.c_int()
for a,name in ((108,'MaxPersistingL2CacheSize'),(109,'MaxAccessPolicyWindowSize'),(38,'L2CacheSize')):
    rt.cudaDeviceGetAttribute(ctypes.byref(v),a,0); print(name,v.value)
" 2>&1 | tail -3
for mb in 0 64 88 104; do
KFSP_L2_KEEP_MB=$mb timeout 600 python tools/phase_breakdown.py goutsias 2>&1 | grep -v "expm n=" | sed "s/^/keep=$mb /" | cut -c1-260
done
KFSP_L2_KEEP_MB=88 KFSP_L2_NO_SETASIDE=1 timeout 600 python tools/phase_breakdown.py goutsias 2>&1 | grep -v "expm n=" | sed "s/^/keep=88 no set-aside /" | cut -c1-260
