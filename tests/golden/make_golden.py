"""Regenerates tests/golden/*.npz from the CPU oracle in canonical arithmetic (reproducible=1).

The reference ships no golden vectors for the solver path and cannot be built in this image
(SURVEY.md 8c), so these fixtures pin the ORACLE (against regressions of the restatement) and give
the GPU tests fixed targets that do not depend on rebuilding anything.  Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from gpu_common_cases import CASES, GOLDEN_RUNS  # noqa: E402


def main():
    for tag, (name, t, ftol, ktol, seed) in GOLDEN_RUNS.items():
        fname, params, x0 = CASES[name]
        m = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", fname), params)
        out = oracle.solve(m, [x0], [1.0], t, ftol, ktol, seed=seed, reproducible=1)
        assert out["iflag"] == 0
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), states=out["states"].astype(np.int16),
                            vector=out["vector"], trace_i=out["trace_i"], trace_d=out["trace_d"],
                            stats=np.array([out["stats"][k] for k in ("nmult", "nexph", "nscale", "nstep", "nreject",
                                                                      "ibrkflag", "mbrkdwn", "n_expand", "n_drop")]))
        print(tag, "N =", len(out["vector"]), "steps =", out["stats"]["nstep"])


if __name__ == "__main__":
    main()
