"""Worker of tests/test_lattice_host.py (world_size-2 gloo, CPU only): the slab partition of the lattice variant
(kfsp_lattice_partition), its halo plan -- ONE plane of the slowest species per neighbour -- and an independent numpy
restatement of the matrix-free row formula of csrc/lattice.cuh, checked against the oracle's FMATVEC on the explicit matrix."""
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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    L = lib()
    bx, by = 37, 29
    states, p0 = bench.synthetic(bx, by)
    om = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", "toggle_test.input"), bench.PARAMS)
    f = oracle.Fsp(om, max_size=bx * by + 8, reproducible=1)
    f.set_states(states)
    f.matrix_starter()
    x = np.random.default_rng(3).standard_normal(bx * by)
    y_ref = f.matvec(x).reshape(by, bx)

    zlo, zhi = C.c_int32(), C.c_int32()
    assert L.kfsp_lattice_partition(by, world, rank, C.byref(zlo), C.byref(zhi)) == 0
    zlo, zhi = zlo.value, zhi.value
    xs = x.reshape(by, bx)                                  # [z, c]: slowest species first
    mine = xs[zlo:zhi].copy()
    # halo: one plane from each neighbour (what the kernel loads straight from the owner's basis column)
    below = torch.zeros(bx, dtype=torch.float64)
    above = torch.zeros(bx, dtype=torch.float64)
    reqs = []
    if rank > 0:
        reqs += [dist.isend(torch.from_numpy(mine[0].copy()), rank - 1), dist.irecv(below, rank - 1)]
    if rank < world - 1:
        reqs += [dist.isend(torch.from_numpy(mine[-1].copy()), rank + 1), dist.irecv(above, rank + 1)]
    for r in reqs:
        r.wait()
    ext = np.vstack([below.numpy()[None, :], mine, above.numpy()[None, :]])      # rows zlo-1 .. zhi
    # tables a_k(count) over the one species each propensity reads; stoichiometry from the model
    stoich = om.stoich                                       # [reaction, species]
    sp = [1, 0, 0, 1]                                        # kx/(1+Y^2.5), ky/(1+X^1.5), dx*X, dy*Y
    tab = []
    for k in range(4):
        st = np.zeros((max(bx, by), 2), dtype=np.int32)
        st[:, sp[k]] = np.arange(max(bx, by))
        tab.append(np.array([om.propensity(s, k + 1) for s in st]))
    cc = np.arange(bx)[None, :].repeat(zhi - zlo, 0)
    zz = np.arange(zlo, zhi)[:, None].repeat(bx, 1)
    cnt = [cc, zz]
    d = np.zeros((zhi - zlo, bx))
    for k in range(4):
        d = d + tab[k][cnt[sp[k]]]
    y = -(d * mine)
    for k in range(4):
        pc, pz = cc - stoich[k, 0], zz - stoich[k, 1]        # predecessor x - nu_k
        ok = (pc >= 0) & (pc < bx) & (pz >= 0) & (pz < by)
        pcnt = [pc, pz]
        a = np.where(ok, tab[k][np.clip(pcnt[sp[k]], 0, max(bx, by) - 1)], 0.0)
        xv = ext[np.clip(pz - (zlo - 1), 0, zhi - zlo + 1), np.clip(pc, 0, bx - 1)]
        y = y + np.where(ok, a * xv, 0.0)
    err = np.abs(y - y_ref[zlo:zhi]).max() / np.abs(y_ref).max()
    ok = err < 1e-13
    # slabs tile the box
    slabs = [None] * world
    dist.all_gather_object(slabs, (zlo, zhi))
    ok = ok and slabs[0][0] == 0 and slabs[-1][1] == by and all(slabs[r][1] == slabs[r + 1][0] for r in range(world - 1))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("rank %d slab [%d,%d) halo planes %d rel err %.2e" % (rank, zlo, zhi, (rank > 0) + (rank < world - 1), err), flush=True)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("LATTICE HOST FAILED")
    if rank == 0:
        print("LATTICE HOST OK")


if __name__ == "__main__":
    main()
