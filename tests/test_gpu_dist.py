"""Row-partitioned multi-GPU solve == single-GPU solve, bit for bit (needs >= 2 GPUs on the box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioned_solve_bit_identical():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 2 if ngpu < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_check.py")],
                       capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DIST CHECK OK" in r.stdout


@pytest.mark.parametrize("min_rows", ["0", None])
def test_adaptive_solve_on_partitioned_handles_bit_identical(min_rows):
    """expansion and pruning enabled on handles partitioned over several GPUs: the committed oracle fixtures, bit for bit.
    KFSP_REPL_MIN_ROWS=0 splits the rows of the Krylov loop over the ranks from the first step on (peer-memory halo, fused
    reduction exchange); the default keeps state sets below 4M rows whole on every rank and exchanges nothing."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 2 if ngpu < 4 else 4
    env = dict(os.environ)
    if min_rows is not None:
        env["KFSP_REPL_MIN_ROWS"] = min_rows
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                        "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "dist_adaptive_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-4000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DIST ADAPTIVE CHECK OK" in r.stdout
