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
