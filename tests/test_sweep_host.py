"""Parameter sweeps: sharding over ranks and gathering, on the CPU (gloo, world size 2)."""
import os
import subprocess
import sys

from krylovfspssa_b200 import sweep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shares_partition_the_sets():
    for n in (0, 1, 5, 16):
        for w in (1, 2, 3, 8):
            got = sorted(i for r in range(w) for i in sweep.my_share(n, r, w))
            assert got == list(range(n))


def test_two_rank_sweep_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29521", os.path.join(ROOT, "tests", "sweep_host_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "SWEEP HOST OK" in r.stdout


def test_concurrent_share_uses_one_handle_and_one_model_copy_per_worker():
    """concurrency = K on one rank: K handles and K model copies, every set solved exactly once with ITS parameters"""
    import threading

    import numpy as np

    lock = threading.Lock()
    made = {"handles": 0, "clones": 0}

    class Model:
        def reset_parameters(self, p):
            self.p = np.asarray(p)

        def clone(self):
            with lock:
                made["clones"] += 1
            return Model()

    class Handle:
        def __init__(self, model, **kw):
            with lock:
                made["handles"] += 1
            self.model = model

        def set_model(self, model):
            assert model is self.model          # a worker never sees another worker's copy

        def solve(self, t, states, p0, ftol, ktol):
            return dict(vector=np.full(3, float(self.model.p[0])), iflag=0, stats={})

        def close(self):
            pass

    sets = [[float(i), 1.0] for i in range(11)]
    out = sweep.run_share(Model(), sets, [0], 1.0, 1e-4, 1e-8, handle_factory=Handle, concurrency=4)
    assert sorted(out) == list(range(11)) and made == {"handles": 4, "clones": 4}
    assert all(out[i]["vector"][0] == float(i) for i in range(11))
