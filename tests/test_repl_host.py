"""Host logic of ADAPTIVE multi-GPU solves on the CPU (gloo, world size 2): the replicated-layout partition arithmetic of the
product (kfsp_repl_partition = Engine::repartition) and the slice-compute / in-place-gather scheme, against the oracle."""
import ctypes as C
import os
import subprocess
import sys

from krylovfspssa_b200._lib import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_repl_partition_arithmetic():
    L = lib()
    for n in (0, 1, 5, 1000, 4194303, 4194304, 10 ** 7):
        for p in (1, 2, 3, 8):
            prev = 0
            for r in range(p):
                lo, hi, whole = C.c_int64(), C.c_int64(), C.c_int32()
                assert L.kfsp_repl_partition(n, p, r, 1 << 22, C.byref(lo), C.byref(hi), C.byref(whole)) == 0
                if p == 1 or n < (1 << 22):
                    assert whole.value == 1 and lo.value == 0 and hi.value == n        # nothing is split below the threshold
                else:
                    assert whole.value == 0 and lo.value == prev and hi.value - lo.value in (n // p, n // p + 1)
                    prev = hi.value
            if p > 1 and n >= (1 << 22):
                assert prev == n
    # more ranks than rows: empty slices are legal (the kernels run with zero rows and contribute zeros to the reductions)
    lo, hi, whole = C.c_int64(), C.c_int64(), C.c_int32()
    assert L.kfsp_repl_partition(3, 8, 7, 0, C.byref(lo), C.byref(hi), C.byref(whole)) == 0
    assert whole.value == 0 and lo.value == hi.value == 3
    assert L.kfsp_repl_partition(3, 8, 8, 0, C.byref(lo), C.byref(hi), C.byref(whole)) < 0


def test_two_rank_replicated_layout_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "repl_host_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "HOST REPL OK" in r.stdout
