"""Batched parameter sweeps: many parameter sets of one model, one CME solve each, as an outer data-parallel axis over
the GPUs of a box.  This is what `RESET_PARAMETERS` (src/model/ModelModule.f90:201-217) exists for in the
reference, where the sets would be solved one after the other on one core.  One process per GPU; parameter set `i` goes to
rank `i mod world` (the sets are independent: no data-path collective, weak scaling); every rank keeps ONE solver handle
and re-ships only the model's parameter values between solves."""
import os

import numpy as np


def my_share(n_items, rank, world):
    """Indices of the items rank `rank` of `world` solves (round robin: sets of similar cost tend to be neighbours)."""
    return list(range(rank, n_items, world))


class SweepPool:
    """`concurrency` solver handles on one GPU, each with its own CUDA stream, buffers and copy of the model, each driven by its
    own host thread (the C calls release the GIL).  The reference's own configurations are latency-bound on a GPU -- a sweep of
    a few thousand states keeps ONE of the 148 SMs busy (single-CTA sweep, Pade kernel) -- so independent solves overlap; every
    set's result is bit-identical to a sequential run (tests/test_gpu_sweep.py).  Handles live as long as the pool: device
    buffers grow during the first solves and are reused afterwards (a cudaMalloc stalls every stream of the device)."""

    def __init__(self, model, concurrency=1, handle_factory=None, blocking=None, **opt_kw):
        if handle_factory is None:
            from .host import KrylovFspHandle
            handle_factory = KrylovFspHandle
        self.factory = handle_factory
        self.opt_kw = opt_kw
        self.concurrency = max(1, int(concurrency))
        # blocking=True: host waits sleep on an event instead of spinning (kfsp_set_blocking_sync).  Measured on a 16-core host
        # (profiles/r2_sweep_throughput.txt): spinning is faster up to 32 threads (a step is two ~10 us waits), so it is off.
        self.blocking = bool(blocking)
        self.models = [model] if self.concurrency == 1 else [model.clone() for _ in range(self.concurrency)]
        self.handles = [None] * self.concurrency

    def _solve_list(self, j, param_sets, idx, x0, t, fsptol, krytol):
        out = {}
        model = self.models[j]
        for i in idx:
            model.reset_parameters(np.asarray(param_sets[i], dtype=np.float64))
            if self.handles[j] is None:
                self.handles[j] = self.factory(model, **dict(self.opt_kw))
                if self.blocking and hasattr(self.handles[j], "set_blocking_sync"):
                    self.handles[j].set_blocking_sync(True)
            else:
                self.handles[j].set_model(model)       # same shape: buffers are reused, only the model block is re-shipped
            out[i] = self.handles[j].solve(t, [x0], [1.0], fsptol, krytol)
        return out

    def run(self, param_sets, idx, x0, t, fsptol, krytol):
        """Solve the sets `idx` of `param_sets`; returns {index: result dict of KrylovFspHandle.solve}."""
        k = min(self.concurrency, max(1, len(idx)))
        if k == 1:
            return self._solve_list(0, param_sets, idx, x0, t, fsptol, krytol)
        from concurrent.futures import ThreadPoolExecutor
        out = {}
        with ThreadPoolExecutor(max_workers=k) as pool:
            futs = [pool.submit(self._solve_list, j, param_sets, idx[j::k], x0, t, fsptol, krytol) for j in range(k)]
            for f in futs:
                out.update(f.result())
        return out

    def close(self):
        for h in self.handles:
            if h is not None:
                h.close()
        self.handles = [None] * self.concurrency


def run_share(model, param_sets, x0, t, fsptol, krytol, rank=0, world=1, handle_factory=None, concurrency=1, **opt_kw):
    """Solve this rank's share.  Returns {index: result dict of KrylovFspHandle.solve}.
    `handle_factory(model, **opt_kw)` creates the solver (default: KrylovFspHandle on device LOCAL_RANK when torchrun set
    it, else `rank`: the two differ as soon as the sweep spans more than one box).
    `concurrency` > 1: that many handles solve at the same time on this rank's GPU (SweepPool)."""
    if handle_factory is None:
        opt_kw.setdefault("device", int(os.environ.get("LOCAL_RANK", rank)))
    idx = my_share(len(param_sets), rank, world)
    pool = SweepPool(model, min(max(1, int(concurrency)), max(1, len(idx))), handle_factory, **opt_kw)
    try:
        return pool.run(param_sets, idx, x0, t, fsptol, krytol)
    finally:
        pool.close()


def gather_summaries(local, n_items):
    """All ranks' per-set summaries (size, mass, steps, SpMVs, device seconds) on every rank, in set order.
    Uses torch.distributed when it is initialised (any backend), else returns the local ones."""
    summ = {i: dict(n=int(len(r["vector"])), mass=float(r["vector"].sum()), nstep=int(r["stats"]["nstep"]),
                    nmult=int(r["stats"]["nmult"]), device_seconds=float(r["stats"]["device_seconds"]), iflag=int(r["iflag"]))
            for i, r in local.items()}
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, summ)
            summ = {}
            for p in parts:
                summ.update(p)
    except ImportError:
        pass
    return [summ.get(i) for i in range(n_items)]
