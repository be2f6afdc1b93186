"""Worker of tests/test_sweep_host.py (gloo, world size 2, CPU): sharding and gathering of a parameter sweep with a stub
solver standing in for the device handle."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from krylovfspssa_b200 import sweep  # noqa: E402


class StubModel:
    def reset_parameters(self, p):
        self.p = np.asarray(p)


class StubHandle:
    created = 0

    def __init__(self, model, **kw):
        StubHandle.created += 1
        self.model = model

    def set_model(self, model):
        self.model = model

    def solve(self, t, states, p0, ftol, ktol):
        n = int(self.model.p[0])
        return dict(vector=np.full(n, 1.0 / n), iflag=0, stats=dict(nstep=n, nmult=10 * n, device_seconds=0.0))

    def close(self):
        pass


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    sets = [[n, 0.5] for n in range(1, 8)]
    local = sweep.run_share(StubModel(), sets, [0], 1.0, 1e-4, 1e-8, rank=rank, world=world, handle_factory=StubHandle)
    assert sorted(local) == list(range(rank, 7, world)) and StubHandle.created == 1
    allr = sweep.gather_summaries(local, len(sets))
    ok = all(r is not None and r["n"] == i + 1 and r["nmult"] == 10 * (i + 1) and abs(r["mass"] - 1) < 1e-12 for i, r in enumerate(allr))
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("SWEEP FAILED")
    if rank == 0:
        print("SWEEP HOST OK")


if __name__ == "__main__":
    main()
