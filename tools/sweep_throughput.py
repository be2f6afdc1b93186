"""Parameter-sweep throughput on ONE GPU: the reference's latency-bound configurations (toggle: a few thousand states, one SM
busy per solve) solved `concurrency` at a time on independent handles/streams (krylovfspssa_b200/sweep.py).
    python tools/sweep_throughput.py [toggle|repressilator] [nsets] [t]
Prints solves per second for concurrency 1, 2, 4, 8, 16, 32 and checks that every set's result equals the sequential one."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovfspssa_b200 as k  # noqa: E402
from krylovfspssa_b200 import sweep  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "toggle"
    nsets = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rng = np.random.default_rng(11)
    if name == "toggle":
        t = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
        sets = [[1.0, 60.0 + 40.0 * rng.random(), 1.0, 1.0, 60.0 + 40.0 * rng.random(), 1.0] for _ in range(nsets)]
        x0, ftol, ktol, cap = [0, 0], 1e-4, 1e-10, 100000
    else:
        t = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
        sets = [[60.0 + 40.0 * rng.random(), 60.0 + 40.0 * rng.random(), 60.0 + 40.0 * rng.random(), 1.0, 1.0, 1.0] for _ in range(nsets)]
        x0, ftol, ktol, cap = [22, 0, 0], 1e-4, 1e-10, 400000
    model = k.CME_MODEL().load(os.path.join(k.models_dir(), name + ".input"))
    print("host cores: %s" % os.cpu_count(), flush=True)
    ref = None
    idx = list(range(nsets))
    blocking = {"spin": False, "block": True}.get(os.environ.get("SWEEP_SYNC", ""), None)
    for conc in (1, 2, 4, 8, 16, 32):
        t0 = time.time()
        pool = sweep.SweepPool(model, conc, blocking=blocking, max_states=cap, seed=12345, device=0)
        pool.run(sets, list(range(conc)), x0, t, ftol, ktol)          # untimed: handle creation, buffer growth, module load
        t1 = time.time()
        res = pool.run(sets, idx, x0, t, ftol, ktol)
        dt = time.time() - t1
        pool.close()
        same = True
        if ref is None:
            ref = res
        else:
            same = all(np.array_equal(ref[i]["vector"], res[i]["vector"]) and np.array_equal(ref[i]["states"], res[i]["states"])
                       for i in range(nsets))
        n_avg = float(np.mean([len(res[i]["vector"]) for i in range(nsets)]))
        print("%s t=%g: %d parameter sets, concurrency %2d: %.2f s wall = %.1f solves/s (mean N = %.0f; pool set-up + first solves %.2f s) "
              "bit-identical to sequential: %s blocking=%s" % (name, t, nsets, conc, dt, nsets / dt, n_avg, t1 - t0, same, pool.blocking), flush=True)


if __name__ == "__main__":
    main()
