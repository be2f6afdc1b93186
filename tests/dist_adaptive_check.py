"""Adaptive Krylov-FSP-SSA on several GPUs, run under torchrun on >= 2 GPUs of one box:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/dist_adaptive_check.py [full]

DGEXPV_FSP with expansion (SSA_EXTENDER + ONESTEP_EXTENDER) and pruning (DROP_STATES) enabled
(src/fsp/KrylovSolver.f90:509-534, src/state_space/StateSpace.f90:347-630) on a handle that was partitioned with
kfsp_dist_init: the rows of every N-sized operation of the Krylov loop are split over the ranks, the state space is
expanded / pruned identically on every rank.  State list, decision trace and probability vector must equal the committed
oracle fixtures (tests/golden/*.npz; with `full`, the SHA-256 digests of the full-horizon runs of BASELINE configs 1-3)
bit for bit on every rank, for both SpMV variants that support irregular sets (explicit, index-only)."""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import krylovfspssa_b200 as k  # noqa: E402
from gpu_common_cases import CASES, GOLDEN_RUNS  # noqa: E402

FULL_RUNS = {
    "toggle_full": ("toggle", 1000.0, 1e-4, 1e-10, 400000),
    "repressilator_full": ("repressilator", 10.0, 1e-4, 1e-10, 2000000),
    "goutsias_full": ("goutsias", 300.0, 1e-6, 1e-8, 6291469),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def new_uid(rank):
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(k.KrylovFspHandle.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    return bytes(uid.cpu().numpy().tobytes())


def handle(name, rank, world, local, variant, cap, seed):
    fname, params, x0 = CASES[name]
    model = k.CME_MODEL().load(os.path.join(k.models_dir(), fname))
    model.reset_parameters(params)
    h = k.KrylovFspHandle(model, max_states=cap, seed=seed, spmv_variant=variant, device=local)
    h.dist_init(rank, world, new_uid(rank))
    return h, x0


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    ok_all = True
    if not full:
        for tag in sorted(GOLDEN_RUNS):
            name, t, ftol, ktol, seed = GOLDEN_RUNS[tag]
            g = np.load(os.path.join(HERE, "golden", tag + ".npz"))
            for variant in (0, 2):
                h, x0 = handle(name, rank, world, local, variant, 400000, seed)
                out = h.solve(t, [x0], [1.0], ftol, ktol)
                info = h.dist_info()
                ok = out["iflag"] == 0 and np.array_equal(out["states"], g["states"]) and np.array_equal(out["vector"], g["vector"])
                # FMATVEC on the final (irregular) state set: whole vector in and out on every rank
                x = np.random.default_rng(5).standard_normal(len(out["vector"]))
                y = h.matvec(x)
                ok = ok and np.isfinite(y).all()
                print("%s variant %d rank %d/%d: N=%d steps=%d expansions=%d drops=%d rows [%d,%d) bit-identical=%s" %
                      (tag, variant, rank, world, len(out["vector"]), out["stats"]["nstep"], out["stats"]["n_expand"],
                       out["stats"]["n_drop"], info["lo"], info["hi"], ok), flush=True)
                ok_all = ok_all and ok
                h.close()
        # CUSTOMPROP (examples/transcr6d.f90: an opaque host function reading two species in reactions 5 and 7) on partitioned
        # handles: every rank evaluates its own copy of the callback while the state space is expanded identically everywhere.
        # Must equal the same solve on an unpartitioned handle of this rank's GPU bit for bit.
        from krylovfspssa_b200 import examples
        d = examples.DRIVERS["transcr6d"]
        os.environ["KFSP_CUSTOM_PROBE"] = "0"               # the callbacks stay host functions (read when a handle is created)
        solo = k.KrylovFspHandle(examples.driver_model("transcr6d"), max_states=400000, seed=4242, device=local)
        ref = solo.solve(30.0, [d["x0"]], [1.0], d["fsp_tol"], d["exp_tol"])
        solo.close()
        for probe in ("0", "1"):                            # "1": the probed model is evaluated on the device -- same bits
            os.environ["KFSP_CUSTOM_PROBE"] = probe
            h = k.KrylovFspHandle(examples.driver_model("transcr6d"), max_states=400000, seed=4242, device=local)
            h.dist_init(rank, world, new_uid(rank))
            assert h.model_info()["n_host_evaluated"] == (d["R"] if probe == "0" else 0)
            out = h.solve(30.0, [d["x0"]], [1.0], d["fsp_tol"], d["exp_tol"])
            ok = out["iflag"] == 0 and ref["iflag"] == 0 and np.array_equal(out["states"], ref["states"])
            ok = ok and np.array_equal(out["vector"], ref["vector"]) and np.array_equal(out["trace"]["i"], ref["trace"]["i"])
            print("transcr6d CUSTOMPROP probe=%s rank %d/%d: N=%d steps=%d expansions=%d host evaluations=%d bit-identical=%s" %
                  (probe, rank, world, len(out["vector"]), out["stats"]["nstep"], out["stats"]["n_expand"],
                   int(h.phase_seconds()["host_propensity_evals"]), ok), flush=True)
            ok_all = ok_all and ok
            h.close()
        del os.environ["KFSP_CUSTOM_PROBE"]
    else:
        db = json.load(open(os.path.join(HERE, "golden", "full_digests.json")))
        for tag, (name, t, ftol, ktol, cap) in sorted(FULL_RUNS.items()):
            if tag not in db:
                continue
            g = db[tag]
            h, x0 = handle(name, rank, world, local, 0, cap, 12345)
            out = h.solve(t, [x0], [1.0], ftol, ktol)
            ok = out["iflag"] == 0 and len(out["vector"]) == g["n"] and sha(out["states"].astype(np.int32)) == g["states_sha256"]
            ok = ok and sha(out["vector"].astype(np.float64)) == g["vector_sha256"]
            ok = ok and sha(out["trace"]["i"].astype(np.int32)) == g["trace_i_sha256"]
            print("%s rank %d/%d: N=%d steps=%d nmult=%d device %.2f s phases %s bit-identical=%s" %
                  (tag, rank, world, len(out["vector"]), out["stats"]["nstep"], out["stats"]["nmult"], out["stats"]["device_seconds"],
                   " ".join("%s=%.2f" % kv for kv in list(h.phase_seconds().items())[:5]), ok), flush=True)
            ok_all = ok_all and ok
            h.close()
    flag = torch.tensor([1 if ok_all else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("DIST ADAPTIVE CHECK FAILED")
    if rank == 0:
        print("DIST ADAPTIVE CHECK OK")


if __name__ == "__main__":
    main()
