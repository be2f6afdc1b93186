"""Developer tool: per-phase time of the adaptive solve for BASELINE configs 1-3 (full horizons), plus the
single-CTA Pade kernel timed alone at several orders."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import krylovfspssa_b200 as k
from gpu_common import make

from krylovfspssa_b200 import examples

DRIVERS = {"driver_toggle": ("toggle", 400000), "driver_repressilator": ("repressilator", 2000000), "transcr6d": ("transcr6d", 6291469)}
RUNS = {"toggle": (1000.0, 1e-4, 1e-10, 400000), "repressilator": (10.0, 1e-4, 1e-10, 2000000), "goutsias": (300.0, 1e-6, 1e-8, 6291469)}
for name in sys.argv[1:] or list(RUNS):
    if name in DRIVERS:                       # the reference's example programs: CUSTOMPROP host callbacks
        d = examples.DRIVERS[DRIVERS[name][0]]
        t, ftol, ktol, cap, x0 = d["t"], d["fsp_tol"], d["exp_tol"], DRIVERS[name][1], d["x0"]
        h = k.KrylovFspHandle(examples.driver_model(DRIVERS[name][0]), max_states=cap, seed=12345)
    else:
        t, ftol, ktol, cap = RUNS[name]
        h, _, x0 = make(name, max_states=cap, seed=12345, spmv_variant=int(os.environ.get("KFSP_VARIANT", "0")))
    out = h.solve(t, [x0], [1.0], ftol, ktol)
    out = h.solve(t, [x0], [1.0], ftol, ktol)
    st = out["stats"]
    ph = h.phase_seconds()
    print("%s: N=%d steps=%d nmult=%d nexph=%d expand=%d drop=%d launches=%d device %.3f s wall %.3f s | %s" %
          (name, st["n_final"], st["nstep"], st["nmult"], st["nexph"], st["n_expand"], st["n_drop"], st["kernel_launches"],
           st["device_seconds"], st["wall_seconds"], " ".join("%s=%.3f" % kv for kv in ph.items())), flush=True)
    if name == "toggle":
        rng = np.random.default_rng(0)
        for n in (12, 32, 52, 77, 102):
            H = np.triu(rng.standard_normal((n, n)), -1)
            h.expm(H, 0.5)
            t0 = time.time()
            for _ in range(20):
                h.expm(H, 0.5)
            print("  expm n=%d: %.1f us per call (incl. H upload and result download)" % (n, 1e6 * (time.time() - t0) / 20))
    h.close()
