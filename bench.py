#!/usr/bin/env python
"""bench.py -- the reference's headline metric on its headline configuration.

Metric (BASELINE.json): expv wall time to t_final + generator-SpMV GB/s vs HBM peak.
Workload (config 5): synthetic toggle switch with copy-number bounds scaled to
~1e8 FSP states -- rectangle [0,Bx) x [0,By), reactions and propensity strings of
krylovfspssa_b200/models/toggle_test.input with (kx,ky,dx,dy) = (5000,1600,1,1), p0 = product of
two discretised Gaussians centred mid-box (sigma = B/16), fixed state set (FSP adaptivity off),
KRYTOL 1e-8, Krylov dimension adapting in [10, 30].

The generator SpMV runs matrix-free on the lattice by default (--spmv-variant 1, csrc/lattice.cuh: FMATVEC recomputed
from the integer state, bit-identical to the explicit matrix); the same line carries a companion measurement
of the identical solves on the explicit gather-ELL matrix (--spmv-variant 0 makes that one the main measurement)
and a `parity` object: the two final vectors compared bit for bit at the benchmarked size.

A "step" is one complete adaptive expv solve exp(t_final*A) p0 over the whole state space.
`value` = generator state updates per second = N * NMULT / time with the state space and p0
resident in HBM (kfsp_solve_resident); `e2e` is the same quantity through the reference-facing
C-ABI call kfsp_solve() with HOST buffers (pinned): H2D of states and p0, MATRIX_STARTER on the
device, the solve, D2H of states and p -- all inside the timed region, at every GPU count.

`--impl reference` times the CPU restatement of the reference (oracle/, netlib-order arithmetic,
1 thread -- the reference is serial) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PARAMS = [5000.0, 1600.0, 1.0, 1.0]
METRIC = "expv_generator_state_updates_per_s"
UNIT = "state-updates/s"
R_TOGGLE = 4
# Algorithmic bytes per state of FMATVEC alone, per SpMV variant: explicit gather-ELL 12R+24, matrix-free lattice
# x + y = 16.  What each launch of the solve must move (the SpMV's operands and results plus those of the vector work fused
# into the same pass) is counted by the library per launch (kfsp_profile_get).
SPMV_ONLY_BYTES = {0: 12 * R_TOGGLE + 24, 1: 16, 2: 4 * R_TOGGLE + 4 * 2 + 24}
SPMV_CLASSES = ("spmv_plain", "spmv_dot", "spmv_nrm", "spmv_fin_dot", "spmv_fin_nrm")
KERNEL_NAME = {0: "k_spmv (generator SpMV, explicit gather ELL, inner products of the IOP window fused; per GPU, rank 0)",
               2: "k_spmv_idx (index-only generator SpMV: pred + integer state streamed, a_k(x - nu_k) recomputed from factored tables, inner "
                  "products of the IOP window fused; per GPU, rank 0)",
               1: "k_spmv_bd2 (matrix-free generator SpMV on the lattice = one whole Arnoldi column per launch: the previous column's two "
                  "DAXPYs + DNRM2 in its load stage, FMATVEC, and the inner products of the IOP window in its epilogue; per GPU, rank 0)"}
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")     # dram__bytes per launch from the round's ncu --set full pass


def synthetic(bx, by):
    """States in index order i = x + Bx*y (x fastest) and the Gaussian p0."""
    x = np.tile(np.arange(bx, dtype=np.int32), by)
    y = np.repeat(np.arange(by, dtype=np.int32), bx)
    states = np.empty((bx * by, 2), dtype=np.int32)
    states[:, 0] = x
    states[:, 1] = y

    def g(n):
        return np.exp(-0.5 * ((np.arange(n) - n / 2.0) / (n / 16.0)) ** 2)
    p0 = np.outer(g(by), g(bx)).ravel()
    p0 /= p0.sum()
    return states, p0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_traffic(variant, states):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the ncu pass recorded in
    profiles/traffic.json (bytes per state at the recorded size, scaled to this run's local states); None if absent."""
    try:
        with open(TRAFFIC_FILE) as fh:
            t = json.load(fh)[str(variant)]
        return t["dram_bytes_per_state"] * states, t["source"]
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [c.strip() for c in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_sample(bx, by, m):
    """FMATVEC + one IOP-2 Arnoldi sweep of the CPU restatement on a bounded rectangle."""
    import oracle
    om = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", "toggle_test.input"), PARAMS)
    states, p0 = synthetic(bx, by)
    n = len(p0)
    f = oracle.Fsp(om, max_size=n + 64)
    f.set_states(states)
    f.matrix_starter()
    work = np.zeros(n * (m + 2))
    nm = C.c_int32()
    t_sweep = oracle.lib().ko_arnoldi_sweep(f.h, p0.ctypes.data_as(C.POINTER(C.c_double)), m,
                                            work.ctypes.data_as(C.POINTER(C.c_double)), None, C.byref(nm))
    y = np.zeros(n)
    t_mv = oracle.lib().ko_time_matvec(f.h, p0.ctypes.data_as(C.POINTER(C.c_double)),
                                       y.ctypes.data_as(C.POINTER(C.c_double)), 5)
    return n, nm.value, t_sweep, t_mv


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement of the reference's own path, serial like the reference, same t_final and Krylov
    dimension range as the repo arm, on a bounded rectangle of the same workload."""
    if rank != 0:
        return
    import oracle
    bx, by = args.ref_bx, args.ref_by
    om = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", "toggle_test.input"), PARAMS)
    states, p0 = synthetic(bx, by)
    n = len(p0)
    times, mults, setups = [], [], []
    for it in range(args.warmup_ref + args.steps):
        out = oracle.solve(om, states, p0, args.t_final, 1e-6, 1e-8, max_size=n + 64, m_max=args.m_max, m_min=10,
                           n_init_onestep=0, enable_drop=0, enable_expand=0)
        if it >= args.warmup_ref:
            # the time-stepping loop only: MATRIX_STARTER (the reference builds its hash table of big-integer keys there) is left
            # out of the reference's time although this repo's e2e figure includes its own set-up -- on a bounded sample the
            # set-up would otherwise dominate and flatter the comparison
            times.append(out["stats"]["wall_seconds"] - out["stats"]["setup_seconds"])
            setups.append(out["stats"]["setup_seconds"])
            mults.append(out["stats"]["nmult"])
    total = sum(times)
    value = n * sum(mults) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup_ref, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic toggle rectangle %dx%d (bounded sample of config 5), expv to t_final=%g, KRYTOL 1e-8, "
                               "Krylov dimension in [10,%d], fixed state set" % (bx, by, args.t_final, args.m_max),
                   "states": n, "m_range": [10, args.m_max]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "setup_seconds_not_counted": sum(setups) / len(setups),
                         "sample": "time-stepping loop of a full expv solve (MATRIX_STARTER excluded) on a %dx%d rectangle (%d states), %d SpMVs per solve; oracle port of the "
                                   "serial Fortran reference (no Fortran compiler in this image or on the GPU box: "
                                   "profiles/r2_fortran_probe_gpu_box.txt), 1 of %d host cores" % (bx, by, n, mults[0], os.cpu_count())},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def sha_of(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).view(np.uint8)).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bx", type=int, default=10000)
    ap.add_argument("--by", type=int, default=10000)
    ap.add_argument("--t-final", type=float, default=0.01)
    ap.add_argument("--m-max", type=int, default=30)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-bx", type=int, default=3200)
    ap.add_argument("--cpu-by", type=int, default=3200)
    ap.add_argument("--ref-bx", type=int, default=2500)
    ap.add_argument("--ref-by", type=int, default=2500)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the bit-for-bit comparisons (profiling runs)")
    ap.add_argument("--scattered", action="store_true",
                    help="SURVEY 8d scattered-order variant: the same state set with its indices permuted by a fixed Philox permutation "
                         "(seed 12345) -- the irregular gather a real FSP ordering produces; explicit matrix only")
    ap.add_argument("--no-companion", action="store_true",
                    help="skip the explicit-matrix companion measurement that a --spmv-variant 1 run adds to its line")
    ap.add_argument("--spmv-variant", type=int, default=1, choices=[0, 1, 2],
                    help="0: explicit gather-ELL matrix (the reference's data model); 1: matrix-free lattice SpMV "
                         "(bit-identical results, 16 instead of 72 bytes per state); 2: index-only SpMV for any state set "
                         "(coefficients recomputed from the integer state, 48 instead of 72 bytes per state)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import krylovfspssa_b200 as k
    from krylovfspssa_b200._lib import Stats, check, lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: krylovfspssa_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        tt = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def new_uid():
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(k.KrylovFspHandle.dist_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return uid.cpu().numpy().tobytes()

    L = lib()
    bx, by = args.bx, args.by
    states_np, p0_np = synthetic(bx, by)
    n = len(p0_np)
    if args.scattered:
        if args.spmv_variant == 1:
            raise SystemExit("--scattered needs --spmv-variant 0 or 2: the lattice variant requires the natural order")
        perm = np.random.Generator(np.random.Philox(12345)).permutation(n)
        states_np = np.ascontiguousarray(states_np[perm])
        p0_np = np.ascontiguousarray(p0_np[perm])
        del perm
    # pinned host buffers (inputs and outputs of the C-ABI call)
    states_h = torch.from_numpy(states_np).pin_memory()
    p0_h = torch.from_numpy(p0_np).pin_memory()
    p_out = torch.empty(n, dtype=torch.float64).pin_memory()
    del states_np

    model = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle_test.input"))
    model.reset_parameters(PARAMS)
    hopts = dict(max_states=n + 64, m_max=args.m_max, m_min=10, n_init_onestep=0, enable_drop=0, enable_expand=0, device=local_rank)
    h = k.KrylovFspHandle(model, spmv_variant=args.spmv_variant, **hopts)
    variant = args.spmv_variant
    i32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    fsp_tol, kry_tol = 1e-6, 1e-8
    lo, hi = 0, n
    if world > 1:
        # rows of the state space are block-partitioned over the ranks (one process per GPU); the NCCL
        # communicator of the library is bootstrapped through torch.distributed
        h.dist_init(rank, world, new_uid())

    # ---- resident setup (not timed for `value`): MATRIX_STARTER on the device, p0 in HBM -------------
    t_setup = time.time()
    check(L.kfsp_fsp_init(h._h, n, C.cast(states_h.data_ptr(), i32p)), "MATRIX_STARTER")
    if world > 1:
        info = h.dist_info()                              # this rank's rows (blocks of rows / slabs of the slowest species)
        lo, hi = info["lo"], info["hi"]
    nloc = hi - lo
    p0_dev = C.c_void_p()
    check(L.kfsp_device_alloc(h._h, 8 * nloc, C.byref(p0_dev)))
    check(L.kfsp_device_upload(h._h, p0_dev, C.c_void_p(p0_h.data_ptr() + 8 * lo), 8 * nloc))
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    # Timed solves: the lattice variant's Arnoldi sweep is nothing but SpMV-class launches (one per column), so ONE event pair
    # per sweep times them without putting anything between the launches (programmatic dependent launch stays effective);
    # the explicit variant interleaves k_finalize and k_spmv, so its launches are bracketed one by one.
    h.set_profiling(1 if args.spmv_variant == 1 else 2)

    def resident_step(hh, pdev, cnt):
        check(L.kfsp_fsp_set_vector_device(hh._h, pdev, cnt))
        st = Stats()
        rc = L.kfsp_solve_resident(hh._h, args.t_final, fsp_tol, kry_tol, 0, C.byref(st))
        if rc < 0:
            raise k.KfspError(rc, "kfsp_solve_resident")
        return st

    def timed_solves(hh, pdev, cnt):
        """warm-up, then K timed resident solves: device seconds (this rank), counters, per-class kernel times"""
        for _ in range(args.warmup):
            resident_step(hh, pdev, cnt)
        barrier()
        out = dict(dev_s=0.0, nmult=0, launches=0, nstep=0, classes={})
        t0 = time.time()
        for _ in range(args.steps):
            st = resident_step(hh, pdev, cnt)
            out["dev_s"] += st.device_seconds
            out["nmult"] += st.nmult
            out["launches"] += st.kernel_launches
            out["nstep"] += st.nstep
            for name, (sec, c, b) in hh.profile().items():
                a = out["classes"].setdefault(name, [0.0, 0, 0])
                a[0] += sec
                a[1] += c
                a[2] += b
        barrier()
        out["wall"] = time.time() - t0
        return out

    def kernel_table(res, var, rows):
        """per kernel class: launches, mean ms, algorithmic GB/s and fraction of the measured HBM peak (this rank's rows)"""
        tab = {}
        for name, (sec, c, b) in res["classes"].items():
            if c == 0:
                continue
            e = {"launches": c, "avg_ms": 1e3 * sec / c, "share_of_step": sec / res["dev_s"] if res["dev_s"] > 0 else None}
            if b:
                e["algorithmic_bytes_per_state"] = b / c
                e["gbs"] = b * rows / sec / 1e9
                e["frac_of_peak"] = e["gbs"] / peak
            tab[name] = e
        return tab

    def spmv_roofline(res, var, rows):
        """the generator SpMV (all its fused variants) against the HBM roofline: bytes the timed launches must move / their time"""
        tot_b = tot_b16 = tot_s = 0.0
        cnt = 0
        for name in SPMV_CLASSES + (("sweep",) if var == 1 else ()):
            sec, c, b = res["classes"].get(name, (0.0, 0, 0))
            tot_b += b * rows
            tot_b16 += SPMV_ONLY_BYTES[var] * rows * c
            tot_s += sec
            cnt += c
        return tot_b, tot_b16, tot_s, cnt

    peak, peak_kind = measured_peak()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        h.dist_exchange_stats(reset=True)
    res = timed_solves(h, p0_dev, nloc)
    xstat = h.dist_exchange_stats() if world > 1 else None
    sampler.stop_flag = True
    sampler.join(timeout=2)
    dev_s = allmax(res["dev_s"])                          # device-timed, max over ranks
    nmult, launches, nstep = res["nmult"], res["launches"], res["nstep"]
    units = float(n) * nmult                              # the whole job: all ranks together update N states per SpMV
    value = units / dev_s
    dinfo = h.dist_info() if world > 1 else None
    if world > 1:
        # the reduction exchange fused into the tail of every reducing kernel: per rank, mean time from posting its partial to
        # holding every rank's (NVLink latency + waiting for the slowest rank); the rank that waits least IS the slowest rank,
        # so the minimum over ranks is the wire latency and the mean over ranks adds the spread of kernel durations
        xs = [None] * world
        dist.all_gather_object(xs, xstat)
        dinfo["exchange_us"] = sum(x["mean_us"] for x in xs) / world
        dinfo["exchange_us_min_over_ranks"] = min(x["mean_us"] for x in xs)
        dinfo["exchange_us_max_over_ranks"] = max(x["mean_us"] for x in xs)
        dinfo["exchange_us_worst_single"] = max(x["max_us"] for x in xs)
        dinfo["exchanges_per_rank"] = xs[0]["exchanges"]

    # plain FMATVEC (mode 0, no fused reduction) on this rank's rows, device-timed: the SURVEY 8d "generator SpMV" number
    plain = None
    if not args.scattered and world == 1:
        ybuf = C.c_void_p()
        check(L.kfsp_device_alloc(h._h, 8 * nloc, C.byref(ybuf)))
        sec = C.c_double()
        check(L.kfsp_matvec_device(h._h, p0_dev, ybuf, 3, C.byref(sec)))
        barrier()
        check(L.kfsp_matvec_device(h._h, p0_dev, ybuf, 10, C.byref(sec)))
        barrier()
        pb = SPMV_ONLY_BYTES[variant]
        plain = {"avg_ms": 1e3 * sec.value, "algorithmic_bytes_per_state": pb, "gbs": pb * nloc / sec.value / 1e9,
                 "frac_of_peak": pb * nloc / sec.value / 1e9 / peak, "launches_timed": 10,
                 "what": "kfsp_matvec_device: FMATVEC alone (KrylovSolver.f90:577-607), no fused reduction, this rank's rows"}
        check(L.kfsp_device_free(h._h, ybuf))

    # ---- end to end through the C ABI with host buffers: kfsp_solve at every GPU count -------------------------------
    # states_out is the same array as states_in, as in the reference (FSP_OUT is in/out, KrylovSolver.f90:7-36)
    def e2e_step():
        st = Stats()
        n_out = C.c_int64()
        rc = L.kfsp_solve(h._h, args.t_final, n, C.cast(states_h.data_ptr(), i32p), C.cast(p0_h.data_ptr(), f64p), fsp_tol,
                          kry_tol, 0, C.byref(n_out), C.cast(states_h.data_ptr(), i32p), C.cast(p_out.data_ptr(), f64p),
                          n, C.byref(st))
        if rc < 0:
            raise k.KfspError(rc, "kfsp_solve")
        assert n_out.value == nloc
        return st

    e2e = None
    total_mass = None
    if args.e2e_steps > 0:
        e2e_step()                                        # warm-up
        barrier()
        t0 = time.time()
        e_mult = 0
        for _ in range(args.e2e_steps):
            e_mult += e2e_step().nmult
        barrier()
        e_wall = allmax(time.time() - t0)
        # the lattice variant ships only this rank's slab of the state list; the explicit one needs the global list for its hash table
        h2d = int((nloc if variant == 1 else n) * 2 * 4 + nloc * 8)
        d2h = int(nloc * 8)
        e2e = {"value": float(n) * e_mult / e_wall, "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": 1e3 * e_wall / args.e2e_steps,
               "timed": "host wall clock around kfsp_solve with pinned host buffers (H2D of states and p0, device MATRIX_STARTER, "
                        "solve, D2H of p; the state list is in/out as in the reference and, the set being fixed, is not copied "
                        "back), barrier on both sides, max over ranks; on N GPUs every rank calls kfsp_solve on its partitioned handle"}
        mass = torch.tensor([float(p_out[:nloc].sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(mass)
        total_mass = float(mass.item())

    tot_b, tot_b16, tot_s, spmv_launches = spmv_roofline(res, variant, nloc)
    achieved = tot_b / tot_s / 1e9 if tot_s > 0 else 0.0
    traffic, traffic_src = measured_traffic(variant, nloc)
    roofline = {"bound": "hbm", "kernel": KERNEL_NAME[variant], "achieved": achieved, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": tot_b / max(spmv_launches, 1),
                "algorithmic_bytes_per_state_per_launch": tot_b / max(spmv_launches, 1) / nloc,
                "spmv_only_bytes_per_state": SPMV_ONLY_BYTES[variant],
                "achieved_spmv_only_bytes": tot_b16 / tot_s / 1e9 if tot_s > 0 else 0.0,
                "frac_spmv_only_bytes": tot_b16 / tot_s / 1e9 / peak if tot_s > 0 else 0.0,
                "timing": "CUDA events on the solver's stream inside the timed solves: " + ("one pair around each Arnoldi sweep, which in this "
                          "variant consists of the column launches of this kernel only (avg = sweep time / launches, gaps between launches included)"
                          if variant == 1 else "one pair around every launch"),
                "note": "achieved = bytes the timed SpMV launches must move (operand, result, and the operands/results of the "
                        "vector work fused into the same pass: DESIGN.md section 4, per class in `kernels`) / their CUDA-event time; "
                        "*_spmv_only_bytes counts FMATVEC's own bytes alone (SURVEY 8d) for the same launches",
                "avg_launch_ms": 1e3 * tot_s / max(spmv_launches, 1),
                "launches_timed": spmv_launches, "share_of_step": tot_s / res["dev_s"] if res["dev_s"] > 0 else None,
                "frac_of_nominal_8TBs": achieved / 8000.0, "plain_spmv": plain}
    if variant == 1:                                      # one more solve with every launch bracketed: the per-class table
        h.set_profiling(2)
        saved_steps, saved_warm = args.steps, args.warmup
        args.steps, args.warmup = 1, 0
        kernels = kernel_table(timed_solves(h, p0_dev, nloc), variant, nloc)
        args.steps, args.warmup = saved_steps, saved_warm
        h.set_profiling(0)
    else:
        kernels = kernel_table(res, variant, nloc)

    # ---- companion: the same resident solves on the EXPLICIT gather-ELL matrix (the reference's data model), so that one
    # line shows both generator-SpMV kernels against the HBM roofline; results of the two variants are bit-identical
    companion = None
    parity = None
    if variant == 1 and not args.no_companion:
        h0 = k.KrylovFspHandle(model, spmv_variant=0, **hopts)
        if world > 1:
            h0.dist_init(rank, world, new_uid())
        check(L.kfsp_fsp_init(h0._h, n, C.cast(states_h.data_ptr(), i32p)), "MATRIX_STARTER")
        lo0, hi0 = 0, n
        if world > 1:
            info0 = h0.dist_info()
            lo0, hi0 = info0["lo"], info0["hi"]
        p0_dev0 = C.c_void_p()
        check(L.kfsp_device_alloc(h0._h, 8 * (hi0 - lo0), C.byref(p0_dev0)))
        check(L.kfsp_device_upload(h0._h, p0_dev0, C.c_void_p(p0_h.data_ptr() + 8 * lo0), 8 * (hi0 - lo0)))
        h0.set_profiling(2)
        cres = timed_solves(h0, p0_dev0, hi0 - lo0)
        c_dev = allmax(cres["dev_s"])
        cb, cb16, cs_, cl = spmv_roofline(cres, 0, hi0 - lo0)
        c_tr, c_src = measured_traffic(0, hi0 - lo0)
        companion = {"spmv_variant": "explicit gather-ELL matrix", "value": float(n) * cres["nmult"] / c_dev, "unit": UNIT,
                     "ms_per_step": 1e3 * c_dev / args.steps, "same_spmv_count_as_main": bool(cres["nmult"] == nmult),
                     "roofline": {"bound": "hbm", "kernel": KERNEL_NAME[0], "achieved": cb / cs_ / 1e9 if cs_ > 0 else 0.0, "peak": peak,
                                  "unit": "GB/s", "frac": cb / cs_ / 1e9 / peak if cs_ > 0 else 0.0,
                                  "frac_spmv_only_bytes": cb16 / cs_ / 1e9 / peak if cs_ > 0 else 0.0,
                                  "avg_launch_ms": 1e3 * cs_ / max(cl, 1), "launches_timed": cl,
                                  "share_of_step": cs_ / cres["dev_s"] if cres["dev_s"] > 0 else None, "traffic": c_tr,
                                  "traffic_source": c_src},
                     "kernels": kernel_table(cres, 0, hi0 - lo0)}
        # parity at the benchmarked size: this rank's rows of the two final vectors, bit for bit.  The lattice handle's last
        # solve was the e2e one (same inputs), its result sits in p_out; the companion's is fetched now.
        if not args.no_parity and (hi0 - lo0) == nloc and lo0 == lo:
            if args.e2e_steps == 0:
                resident_step(h, p0_dev, nloc)
                check(L.kfsp_fsp_get(h._h, None, None, None, None, C.cast(p_out.data_ptr(), f64p)))
            pc = torch.empty(nloc, dtype=torch.float64).pin_memory()
            check(L.kfsp_fsp_get(h0._h, None, None, None, None, C.cast(pc.data_ptr(), f64p)))
            a, b = p_out[:nloc].numpy(), pc.numpy()
            same = bool(np.array_equal(a.view(np.int64), b.view(np.int64)))
            maxabs = allmax(float(np.abs(a - b).max()))
            allsame = allmax(0.0 if same else 1.0) == 0.0
            digests = [None] * world
            mine = sha_of(a)
            if world > 1:
                dist.all_gather_object(digests, mine)
            else:
                digests = [mine]
            parity = {"lattice_vs_explicit_bit_identical": allsame, "lattice_vs_explicit_max_abs": maxabs,
                      "vector_sha256": hashlib.sha256("".join(digests).encode()).hexdigest() if world > 1 else mine,
                      "states_compared": n, "what": "final probability vectors of the matrix-free and the explicit-matrix solves at the "
                                                    "benchmarked size, compared bit for bit on every rank's rows (sha256 over the ranks' digests)"}
            del pc
        h0.close()

    # ---- multi-GPU parity inside the bench run: a reduced rectangle solved alone on every rank and partitioned over all of
    # them (both SpMV variants, kfsp_solve and kfsp_matvec on the partitioned handles): tests/dist_check.py
    dist_parity = None
    if world > 1 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import dist_check
        ok = dist_check.run_case(rank, world, local_rank, 1200, 800, 0.03, quiet=True)
        dist_parity = {"dist_bit_identical": bool(ok), "case": "1200x800 rectangle, t=0.03: solo on each GPU vs partitioned over %d GPUs, "
                       "explicit and lattice variants, resident solve + kfsp_solve + kfsp_matvec (tests/dist_check.py)" % world}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cn, cm, t_sweep, t_mv = cpu_sample(args.cpu_bx, args.cpu_by, 10)
        cpu = {"value": cn * cm / t_sweep, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "FMATVEC + one IOP-2 Arnoldi sweep (m=10, %d SpMVs) on a %dx%d rectangle of the same workload "
                         "(%d states); oracle port of the serial Fortran reference, 1 of %d host cores"
                         % (cm, args.cpu_bx, args.cpu_by, cn, os.cpu_count()),
               "spmv_states_per_s": cn / t_mv, "spmv_gbs": SPMV_ONLY_BYTES[0] * cn / t_mv / 1e9}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 5: synthetic toggle, rectangle %dx%d = %d FSP states, expv to t_final=%g, "
                                   "KRYTOL 1e-8, Krylov dimension in [10,%d], fixed state set%s"
                                   % (bx, by, n, args.t_final, args.m_max, ", indices in scattered (Philox-permuted) order" if args.scattered else ""),
                       "states": n, "reactions": R_TOGGLE,
                       "spmv_variant": "matrix-free lattice (FMATVEC recomputed from the integer state, bit-identical to the explicit "
                                       "matrix)" if variant == 1 else "explicit gather-ELL matrix (ADJ/OFFDIAG/DIAG in HBM)" if variant == 0 else
                                       "index-only (pred + integer state in HBM, coefficients recomputed; bit-identical to the explicit matrix)",
                       "l2": "inputs (%.2f GB per vector%s) %s the 126 MB L2" % (8e-9 * n, ", %.1f GB matrix" % ((56e-9 if variant == 0 else 32e-9) * n) if variant != 1 else "",
                                                                               "exceed" if 8 * n > 126e6 else "DO NOT exceed"),
                       "parallelism": "1 GPU" if world == 1 else
                       "rows block-partitioned over %d GPUs; per SpMV the halo is gathered straight from the neighbours' HBM and per "
                       "reduction the double-double partials are exchanged inside the reducing kernel (cudaIpc peer memory over "
                       "NVLink/NVSwitch; NCCL only bootstraps, KFSP_DIST_P2P=0 selects the NCCL send/recv + all-gather path)" % world},
            "expv_wall_s_to_t_final": dev_s / args.steps, "krylov_steps_per_solve": nstep / args.steps,
            "spmv_per_solve": nmult / args.steps, "setup_s": t_setup, "host_wall_s": res["wall"],
            "roofline": roofline, "kernels": kernels, "explicit_matrix_companion": companion, "parity": parity,
            "dist_parity": dist_parity, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": sampler.summary(), "probability_mass_out": total_mass,
            "dist": dinfo,
        }
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
