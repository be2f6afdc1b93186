"""BASELINE configs 1-4 (and the other example drivers, CUSTOMPROP host callbacks) at their full horizons on the GPU, against SHA-256 digests of the CPU oracle's
canonical-arithmetic outputs (tests/golden/full_digests.json, made by tests/golden/make_full_digests.py):
state list, decision trace and probability vector must be bit-identical at full scale."""
import hashlib
import json
import os

import numpy as np
import pytest

from gpu_common import make

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FULL_RUNS = {
    "toggle_full": ("toggle", 1000.0, 1e-4, 1e-10, 400000),
    "repressilator_full": ("repressilator", 10.0, 1e-4, 1e-10, 2000000),
    "goutsias_full": ("goutsias", 300.0, 1e-6, 1e-8, 6291469),
}
# the reference's example programs (hard-coded CUSTOMPROP): tag -> (driver, max_states); config 4 = transcr6d
DRIVER_RUNS = {"driver_toggle_full": ("toggle", 400000), "driver_repressilator_full": ("repressilator", 2000000),
               "transcr6d_full": ("transcr6d", 6291469)}
STAT_KEYS = ("nmult", "nexph", "nscale", "nstep", "nreject", "ibrkflag", "mbrkdwn", "n_expand", "n_drop")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", sorted(FULL_RUNS) + sorted(DRIVER_RUNS))
def test_full_config_bit_identical(tag):
    path = os.path.join(HERE, "golden", "full_digests.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    if tag not in db:
        pytest.skip("no digest committed for " + tag)
    g = db[tag]
    if tag in DRIVER_RUNS:
        import krylovfspssa_b200 as k
        from krylovfspssa_b200 import examples
        name, cap = DRIVER_RUNS[tag]
        d = examples.DRIVERS[name]
        x0, t, ftol, ktol = d["x0"], d["t"], d["fsp_tol"], d["exp_tol"]
        h = k.KrylovFspHandle(examples.driver_model(name), max_states=cap, seed=12345)
    else:
        name, t, ftol, ktol, cap = FULL_RUNS[tag]
        h, _, x0 = make(name, max_states=cap, seed=12345)
    out = h.solve(t, [x0], [1.0], ftol, ktol)
    st = out["stats"]
    print("%s: N=%d steps=%d nmult=%d device %.2f s (oracle %.1f s on one CPU core)" %
          (tag, len(out["vector"]), st["nstep"], st["nmult"], st["device_seconds"], g.get("oracle_wall_s", float("nan"))))
    assert out["iflag"] == 0
    assert len(out["vector"]) == g["n"]
    assert {k: int(st[k]) for k in STAT_KEYS} == g["stats"]
    assert sha(out["trace"]["i"].astype(np.int32)) == g["trace_i_sha256"]
    assert sha(out["trace"]["d"][:, 1].astype(np.float64)) == g["t_step_sha256"]
    assert sha(out["states"].astype(np.int32)) == g["states_sha256"]          # state set and indices bit-exact
    assert abs(out["vector"].sum() - g["vector_sum"]) <= 1e-10                # north-star tolerance on the mass
    assert sha(out["vector"].astype(np.float64)) == g["vector_sha256"]        # and in fact bit-identical
    h.close()
