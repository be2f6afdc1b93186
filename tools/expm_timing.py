"""Developer tool: time the single-CTA Pade kernel alone (kfsp_expm) for several orders and scalings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from gpu_common import make
h, _, _ = make("birth_death", max_states=1000)
rng = np.random.default_rng(0)
for n in (32, 62, 102):
    H = np.zeros((n, n))
    for j in range(n - 2):
        H[j, j] = -abs(rng.standard_normal()) * 5; H[j + 1, j] = abs(rng.standard_normal()) * 5
        if j > 0: H[j - 1, j] = rng.standard_normal() * 5
    H[n - 1, n - 2] = 1.0
    for t in (1e-6, 0.05, 5.0):
        E, ns, hn = h.expm(H, t)
        t0 = time.time()
        for _ in range(30): h.expm(H, t)
        print("n=%3d t=%g ns=%2d : %.0f us per call (incl. ~2x80 KB transfers)" % (n, t, ns, 1e6 * (time.time() - t0) / 30), flush=True)
h.close()
