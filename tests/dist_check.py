"""Multi-GPU parity check, run under torchrun on >= 2 GPUs of one box:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py

Every rank solves the fixed-state-set synthetic toggle (config 5, small rectangle) twice: alone on its own
GPU, and row-partitioned over all ranks (NCCL halo exchange + double-double all-gather reductions).  The
partitioned result must be BIT-IDENTICAL to the single-GPU one: same decision trace, same counters, and
this rank's slice of the probability vector equal bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import krylovfspssa_b200 as k  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [(int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]))] if len(sys.argv) > 3 else \
        [(301, 257, 0.05), (97, 1031, 0.02), (1200, 800, 0.03), (301, 257, 0.05)]
    all_ok = True
    for bx, by, t_final in cases:
        all_ok = run_case(rank, world, local, bx, by, t_final) and all_ok
    dist.destroy_process_group()
    if not all_ok:
        raise SystemExit("DIST CHECK FAILED")
    if rank == 0:
        print("DIST CHECK OK")


def run_case(rank, world, local, bx, by, t_final, quiet=False):
    states, p0 = bench.synthetic(bx, by)
    n = len(p0)
    model = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle_test.input"))
    model.reset_parameters(bench.PARAMS)
    opts = dict(max_states=n + 64, m_max=30, m_min=10, n_init_onestep=0, enable_drop=0, enable_expand=0, device=local)

    solo = k.KrylovFspHandle(model, **opts)
    solo.fsp_init(states)
    solo.set_vector(p0)
    rc1, st1 = solo.solve_resident(t_final, 1e-6, 1e-8)
    ref = solo.get(matrix=False)["vector"]
    tr1 = solo.trace()

    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(k.KrylovFspHandle.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    uid2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid2.copy_(torch.frombuffer(bytearray(k.KrylovFspHandle.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid2, 0)
    part = k.KrylovFspHandle(model, **opts)
    part.dist_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    part.fsp_init(states)
    part.set_vector(p0)
    rc2, st2 = part.solve_resident(t_final, 1e-6, 1e-8)
    info = part.dist_info()
    mine = part.get(matrix=False)["vector"]
    tr2 = part.trace()

    # matrix-free lattice variant, partitioned in slabs of the slowest species: same bits again
    box = k.KrylovFspHandle(model, spmv_variant=1, **opts)
    box.dist_init(rank, world, bytes(uid2.cpu().numpy().tobytes()))
    box.fsp_init_box([bx, by])
    box.set_vector(p0)
    rc3, st3 = box.solve_resident(t_final, 1e-6, 1e-8)
    binfo = box.dist_info()
    bmine = box.get(matrix=False)["vector"]
    tr3 = box.trace()
    okb = rc3 == 0 and np.array_equal(tr1["i"], tr3["i"]) and np.array_equal(tr1["d"], tr3["d"])
    okb = okb and len(bmine) == binfo["hi"] - binfo["lo"] and np.array_equal(bmine, ref[binfo["lo"]:binfo["hi"]])
    box.close()

    # the reference-facing call on partitioned handles (global list and vector in, this rank's rows out), and FMATVEC on an
    # operand that is not a column of the basis (staged for the peer-memory halo) -- both variants
    okc = True
    xg = np.random.default_rng(7).standard_normal(n)
    yref = solo.matvec(xg)
    for variant in (0, 1):
        uid3 = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid3.copy_(torch.frombuffer(bytearray(k.KrylovFspHandle.dist_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid3, 0)
        hh = k.KrylovFspHandle(model, spmv_variant=variant, **opts)
        hh.dist_init(rank, world, bytes(uid3.cpu().numpy().tobytes()))
        out = hh.solve(t_final, states, p0, 1e-6, 1e-8)
        inf = hh.dist_info()
        lo, hi = inf["lo"], inf["hi"]
        okc = okc and out["iflag"] == 0 and len(out["vector"]) == hi - lo
        okc = okc and np.array_equal(out["vector"], ref[lo:hi]) and np.array_equal(out["states"], states[lo:hi])
        okc = okc and np.array_equal(out["trace"]["i"], tr1["i"])
        okc = okc and np.array_equal(hh.matvec(xg[lo:hi]), yref[lo:hi])
        hh.close()

    ok = rc1 == rc2 == 0 and okb and okc
    ok = ok and np.array_equal(tr1["i"], tr2["i"]) and np.array_equal(tr1["d"], tr2["d"])
    ok = ok and all(st1[key] == st2[key] for key in ("nmult", "nexph", "nscale", "nstep", "nreject"))
    ok = ok and len(mine) == info["hi"] - info["lo"] and np.array_equal(mine, ref[info["lo"]:info["hi"]])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not quiet:
      print("%dx%d rank %d/%d rows [%d,%d) halo %d send %d steps %d nmult %d bit-identical=%s lattice[%d,%d)=%s kfsp_solve+matvec=%s max|diff|=%.3e" %
          (bx, by, rank, world, info["lo"], info["hi"], info["n_halo"], info["n_send"], st2["nstep"], st2["nmult"], ok,
           binfo["lo"], binfo["hi"], okb, okc,
           float(np.abs(mine - ref[info["lo"]:info["hi"]]).max()) if len(mine) == info["hi"] - info["lo"] else -1.0), flush=True)
    solo.close()
    part.close()
    return int(flag.item()) == 1


if __name__ == "__main__":
    main()
