"""Worker of tests/test_repl_host.py (world_size-2 gloo, CPU only): the host logic of ADAPTIVE solves on several GPUs
(csrc/engine.cuh: repl_enter / repartition).  Every rank holds the whole state space in the single-GPU
layout; in the Krylov loop rank r computes rows [lo_r, hi_r) of y = A x, gathering x by GLOBAL index (rows of other ranks come
from the owner's copy at the same offset); before expansion / pruning the ranks' slices of W are gathered in place.  Here gloo
stands in for NVLink/NCCL and numpy for the kernels: the sliced product, reassembled, must equal the oracle's FMATVEC bit for
bit while the state set grows through ONESTEP_EXTENDER, across the whole -> split transition."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from gpu_common_cases import CASES  # noqa: E402
from krylovfspssa_b200._lib import lib  # noqa: E402


_libm = C.CDLL("libm.so.6")
_libm.fma.restype = C.c_double
_libm.fma.argtypes = [C.c_double, C.c_double, C.c_double]


def plan(n, world, rank, min_rows):
    lo, hi, whole = C.c_int64(), C.c_int64(), C.c_int32()
    assert lib().kfsp_repl_partition(n, world, rank, min_rows, C.byref(lo), C.byref(hi), C.byref(whole)) == 0
    return lo.value, hi.value, bool(whole.value)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    fname, params, x0 = CASES["goutsias"]
    om = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", fname), params)
    f = oracle.Fsp(om, reproducible=1)
    f.set_states([x0]); f.matrix_starter()
    min_rows = 400                                             # the set starts whole and is split once it has grown
    rng = np.random.default_rng(3)
    seen_whole = seen_split = False
    for step in range(7):
        f.onestep()                                            # identically on every rank: the state space is replicated
        n = f.size
        lo, hi, whole = plan(n, world, rank, min_rows)
        bounds = [plan(n, world, r, min_rows) for r in range(world)]
        if whole:
            seen_whole = True
            assert all(b[0] == 0 and b[1] == n for b in bounds)
        else:
            seen_split = True
            assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
        # replicated W (same seed on every rank), FMATVEC by slices with global-index gathers from the "owner's copy"
        x = np.random.default_rng(100 + step).standard_normal(n)
        yref = f.matvec(x)
        get = f.get()
        adj, off, diag = get["adj"], get["offdiag"], get["diag"]      # column form: ADJ(k,i) 1-based successor, OFFDIAG(k,i) = a_k(x_i)
        y = np.zeros(n)
        # gather form of rows [lo,hi): y_i = -d_i x_i + sum_k a_k(x_j) x_j over predecessors j with ADJ(k,j) = i, in reaction order
        R = adj.shape[1]
        terms = [[] for _ in range(n)]
        for j in range(n):
            for k in range(R):
                i = adj[j, k] - 1
                if i >= 0:
                    terms[i].append((k, j))
        for i in range(lo, hi):
            sv = -(diag[i] * x[i])
            for k, j in sorted(terms[i]):
                sv = _libm.fma(float(off[j, k]), float(x[j]), float(sv))     # the kernels' fused multiply-add, reaction order
            y[i] = sv
        # in-place all-gather of the slices (grouped broadcasts, one per owner), as gather_rows does with ncclBroadcast
        yt = torch.from_numpy(y)
        for r in range(world):
            blo, bhi, _ = bounds[r]
            if whole:
                break
            if bhi > blo:
                piece = yt[blo:bhi].clone()
                dist.broadcast(piece, r)
                yt[blo:bhi] = piece
        assert np.array_equal(yt.numpy(), yref), "sliced FMATVEC differs from the oracle at n=%d" % n
    assert seen_whole and seen_split
    dist.barrier()
    if rank == 0:
        print("HOST REPL OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
