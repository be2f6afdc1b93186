"""Worker of tests/test_dist_host.py (world_size-2 gloo, CPU only): exercises the product's partition
arithmetic (kfsp_dist_partition / kfsp_dist_owner) and the halo-exchange PLAN of the multi-GPU SpMV --
which rows each rank must send to whom -- with gloo standing in for NCCL and numpy for the kernels."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402
from krylovfspssa_b200._lib import lib  # noqa: E402


def exchange(out_list, in_list, rank):
    """grouped send/recv, the shape of the product's ncclGroupStart/ncclSend/ncclRecv/ncclGroupEnd step"""
    reqs = []
    for p, t in enumerate(out_list):
        if p != rank and t.numel():
            reqs.append(dist.isend(t.contiguous(), p))
    for p, t in enumerate(in_list):
        if p != rank and t.numel():
            reqs.append(dist.irecv(t, p))
    for r in reqs:
        r.wait()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    L = lib()
    bx, by = 37, 29
    states, p0 = bench.synthetic(bx, by)
    n = len(p0)
    om = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", "toggle_test.input"), bench.PARAMS)
    f = oracle.Fsp(om, max_size=n + 8, reproducible=1)
    f.set_states(states)
    f.matrix_starter()
    d = f.get()
    lo, hi = C.c_int64(), C.c_int64()
    assert L.kfsp_dist_partition(n, world, rank, C.byref(lo), C.byref(hi)) == 0
    lo, hi = lo.value, hi.value
    bounds = []
    for r in range(world + 1):
        a, b = C.c_int64(), C.c_int64()
        L.kfsp_dist_partition(n, world, min(r, world - 1), C.byref(a), C.byref(b))
        bounds.append(a.value if r < world else b.value)
    for g in (0, lo, hi - 1, n - 1, n // 2):
        o = C.c_int32()
        assert L.kfsp_dist_owner(n, world, g, C.byref(o)) == 0
        assert bounds[o.value] <= g < bounds[o.value + 1]
    # gather form of this rank's rows from the column form: row j receives offdiag[i,k] from i = pred
    pred = -np.ones((n, om.R), dtype=np.int64)
    coef = np.zeros((n, om.R))
    for i in range(n):
        for kk in range(om.R):
            j = d["adj"][i, kk]
            if j > 0:
                pred[j - 1, kk] = i
                coef[j - 1, kk] = d["offdiag"][i, kk]
    mine = pred[lo:hi]
    remote = np.unique(mine[(mine >= 0) & ((mine < lo) | (mine >= hi))])          # ascending => grouped by owner
    want = [remote[(remote >= bounds[o]) & (remote < bounds[o + 1])] for o in range(world)]
    # exchange request lists (counts, then indices)
    cnt_out = torch.tensor([len(w) for w in want], dtype=torch.int64)
    cnt_in = torch.zeros(world, dtype=torch.int64)
    dist.all_to_all_single(cnt_in, cnt_out)
    req_out = [torch.from_numpy(w.astype(np.int64)) for w in want]
    req_in = [torch.zeros(int(c), dtype=torch.int64) for c in cnt_in]
    exchange(req_out, req_in, rank)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n)
    x_loc = x[lo:hi]
    # one SpMV exchange step: pack what the peers asked for, receive our halo
    send = [torch.from_numpy(x_loc[(r.numpy() - lo)]) for r in req_in]
    recv = [torch.zeros(len(w), dtype=torch.float64) for w in want]
    exchange(send, recv, rank)
    halo = np.concatenate([r.numpy() for r in recv]) if remote.size else np.zeros(0)
    pos = {int(g): q for q, g in enumerate(remote)}
    y = np.zeros(hi - lo)
    for il in range(hi - lo):
        s = -(d["diag"][lo + il] * x_loc[il])
        for kk in range(om.R):
            g = mine[il, kk]
            if g >= 0:
                xv = x_loc[g - lo] if lo <= g < hi else halo[pos[int(g)]]
                s = coef[lo + il, kk] * xv + s
        y[il] = s
    ref = f.matvec(x)[lo:hi]
    err = np.abs(y - ref).max()
    ok = err <= 1e-12 * max(1.0, np.abs(ref).max())
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("HOST DIST OK" if int(flag) == 1 else "HOST DIST FAILED", "halo", len(remote), "err", err)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
