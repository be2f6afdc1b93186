"""N > 1 host logic on the CPU (gloo, world size 2): partition arithmetic of the product and the halo plan."""
import ctypes as C
import os
import subprocess
import sys

import pytest

from krylovfspssa_b200._lib import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_rows_exactly():
    L = lib()
    for n in (7, 100, 1001, 10 ** 8):
        for p in (1, 2, 3, 4, 8):
            if n < p:
                continue
            prev = 0
            for r in range(p):
                lo, hi = C.c_int64(), C.c_int64()
                assert L.kfsp_dist_partition(n, p, r, C.byref(lo), C.byref(hi)) == 0
                assert lo.value == prev and hi.value - lo.value in (n // p, n // p + 1)
                prev = hi.value
                for g in {lo.value, hi.value - 1, (lo.value + hi.value) // 2}:
                    o = C.c_int32()
                    assert L.kfsp_dist_owner(n, p, g, C.byref(o)) == 0 and o.value == r
            assert prev == n


def test_two_rank_halo_plan_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_host_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "HOST DIST OK" in r.stdout
