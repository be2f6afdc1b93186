"""Batched parameter sweeps: many parameter sets of one model, one CME solve each, as an outer data-parallel axis over
the GPUs of a box.  This is what `RESET_PARAMETERS` (src/model/ModelModule.f90:201-217) exists for in the
reference, where the sets would be solved one after the other on one core.  One process per GPU; parameter set `i` goes to
rank `i mod world` (the sets are independent: no data-path collective, weak scaling); every rank keeps ONE solver handle
and re-ships only the model's parameter values between solves."""
import numpy as np


def my_share(n_items, rank, world):
    """Indices of the items rank `rank` of `world` solves (round robin: sets of similar cost tend to be neighbours)."""
    return list(range(rank, n_items, world))


def run_share(model, param_sets, x0, t, fsptol, krytol, rank=0, world=1, handle_factory=None, **opt_kw):
    """Solve this rank's share.  Returns {index: result dict of KrylovFspHandle.solve}.
    `handle_factory(model, **opt_kw)` creates the solver (default: KrylovFspHandle on device `rank`)."""
    if handle_factory is None:
        from .host import KrylovFspHandle
        opt_kw.setdefault("device", rank)
        handle_factory = KrylovFspHandle
    idx = my_share(len(param_sets), rank, world)
    out = {}
    h = None
    for i in idx:
        model.reset_parameters(np.asarray(param_sets[i], dtype=np.float64))
        if h is None:
            h = handle_factory(model, **opt_kw)
        else:
            h.set_model(model)                     # same shape: buffers are reused, only the model block is re-shipped
        out[i] = h.solve(t, [x0], [1.0], fsptol, krytol)
    if h is not None:
        h.close()
    return out


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
