"""GPU parity of the whole adaptive solve (CME_SOLVE / DGEXPV_FSP) through the C ABI."""
import math
import os

import numpy as np
import pytest

import krylovfspssa_b200 as k
import oracle
from gpu_common import CASES, make, rel1

pytestmark = pytest.mark.gpu


def test_birth_death_poisson():
    h, om, x0 = make("birth_death")
    out = h.solve(2.0, [x0], [1.0], 1e-6, 1e-10)
    assert out["iflag"] == 0
    mean = 20.0 * (1 - math.exp(-2.0))
    x = out["states"][:, 0]
    pm = np.array([math.exp(-mean + xx * math.log(mean) - math.lgamma(xx + 1)) for xx in x])
    assert np.abs(out["vector"] - pm).sum() < 5e-6
    assert out["vector"].min() >= 0 and 1 - 1e-6 <= out["vector"].sum() <= 1 + 1e-12
    h.close()


@pytest.mark.parametrize("name,t,ftol,ktol", [("toggle", 5.0, 1e-4, 1e-10), ("goutsias", 10.0, 1e-6, 1e-8),
                                              ("repressilator", 0.5, 1e-4, 1e-10)])
def test_solve_matches_oracle(name, t, ftol, ktol):
    h, om, x0 = make(name)
    out = h.solve(t, [x0], [1.0], ftol, ktol)
    ref = oracle.solve(om, [x0], [1.0], t, ftol, ktol, reproducible=1)
    assert out["iflag"] == ref["iflag"] == 0
    ti, td = out["trace"]["i"], out["trace"]["d"]
    ri, rd = ref["trace_i"], ref["trace_d"]
    same = 0
    for a, b, c, d in zip(ti, ri, td, rd):
        if not (np.array_equal(a, b) and c[1] == d[1]):
            break
        same += 1
    print("trace rows identical: %d of %d (oracle %d)" % (same, len(ti), len(ri)))
    assert abs(td[-1, 0] - t) < 1e-12
    acc = ti[:, 3] & 4 == 0
    for tn, ws in zip(td[acc, 0], td[acc, 3]):
        assert 1 - ftol * tn / t - 1e-14 <= ws <= 1 + 1e-12
    assert out["vector"].min() >= 0
    # compare on the union of the two state sets
    do = {tuple(s): p for s, p in zip(ref["states"], ref["vector"])}
    dg = {tuple(s): p for s, p in zip(out["states"], out["vector"])}
    keys = set(do) | set(dg)
    err = sum(abs(do.get(q, 0.0) - dg.get(q, 0.0)) for q in keys)
    print("1-norm difference to the oracle: %.3e" % err)
    assert err < 10 * ftol
    # the parity bar: identical decision trace, identical state list, p within 1e-10 relative 1-norm
    assert same == len(ri) == len(ti)
    assert np.array_equal(out["states"], ref["states"])
    assert rel1(out["vector"], ref["vector"]) <= 1e-10
    print("bitwise identical vector:", np.array_equal(out["vector"], ref["vector"]))
    for key in ("nmult", "nexph", "nscale", "nstep", "nreject", "ibrkflag", "mbrkdwn", "n_expand", "n_drop"):
        assert out["stats"][key] == ref["stats"][key], key
    h.close()


def test_cme_solve_host_api():
    fname, params, x0 = CASES["toggle"]
    model = k.CME_MODEL().load(os.path.join(k.models_dir(), fname))
    model.reset_parameters(params)
    fsp_in, fsp = k.FINITE_STATE_PROJECTION().create(model, 100000), k.FINITE_STATE_PROJECTION().create(model, 100000)
    fsp_in.set([x0], [1.0])
    fsp.set([x0], [1.0])
    out = k.cme_solve(model, 1.0, fsp_in, fsp, 1e-4, 1e-10, verbosity=0)
    assert out["iflag"] == 0 and fsp.size == fsp.state.shape[1] == len(fsp.vector)
    j = fsp.index(fsp.state[:, 3])
    assert j == 4 and fsp.probability(fsp.state[:, 3]) == fsp.vector[3]
    assert fsp.index([9000, 9000]) == 0 and fsp.probability([9000, 9000]) == 0.0
    fsp.clear()
