"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU part: the oracle in canonical arithmetic reproduces its committed outputs exactly, and the
netlib-order restatement agrees with them to the FSP tolerance.  GPU part: the CUDA path through
the C ABI reproduces the fixtures bit for bit (state list, decision trace) and within 1e-10 relative
1-norm (probability vector) -- without the oracle being involved at run time."""
import os

import numpy as np
import pytest

import oracle
from gpu_common_cases import CASES, GOLDEN_RUNS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAT_KEYS = ("nmult", "nexph", "nscale", "nstep", "nreject", "ibrkflag", "mbrkdwn", "n_expand", "n_drop")


def load(tag):
    return np.load(os.path.join(HERE, "golden", tag + ".npz"))


@pytest.mark.parametrize("tag", ["toggle_t20", "goutsias_t30", "birth_death_t2"])
def test_oracle_reproduces_golden(tag):
    name, t, ftol, ktol, seed = GOLDEN_RUNS[tag]
    fname, params, x0 = CASES[name]
    g = load(tag)
    m = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", fname), params)
    out = oracle.solve(m, [x0], [1.0], t, ftol, ktol, seed=seed, reproducible=1)
    assert np.array_equal(out["states"], g["states"])
    assert np.array_equal(out["trace_i"], g["trace_i"])
    assert np.array_equal(out["trace_d"][:, 1], g["trace_d"][:, 1])
    assert np.array_equal(out["vector"], g["vector"])
    assert [out["stats"][k] for k in STAT_KEYS] == list(g["stats"])


@pytest.mark.parametrize("tag", ["toggle_t20", "birth_death_t2"])
def test_netlib_order_agrees_with_golden_to_fsp_tolerance(tag):
    name, t, ftol, ktol, seed = GOLDEN_RUNS[tag]
    fname, params, x0 = CASES[name]
    g = load(tag)
    m = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", fname), params)
    out = oracle.solve(m, [x0], [1.0], t, ftol, ktol, seed=seed, reproducible=0)
    a = {tuple(s): p for s, p in zip(out["states"], out["vector"])}
    b = {tuple(int(v) for v in s): p for s, p in zip(g["states"], g["vector"])}
    err = sum(abs(a.get(q, 0.0) - b.get(q, 0.0)) for q in set(a) | set(b))
    assert err < 2 * ftol


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(GOLDEN_RUNS))
def test_gpu_reproduces_golden(tag):
    from gpu_common import make, rel1
    name, t, ftol, ktol, seed = GOLDEN_RUNS[tag]
    h, _, x0 = make(name, seed=seed)
    out = h.solve(t, [x0], [1.0], ftol, ktol)
    g = load(tag)
    assert out["iflag"] == 0
    assert np.array_equal(out["states"], g["states"])                       # state set and indices bit-exact
    assert np.array_equal(out["trace"]["i"], g["trace_i"])                  # (M, N, N_after, flags, NMULT, NEXPH) per step
    assert np.array_equal(out["trace"]["d"][:, 1], g["trace_d"][:, 1])      # T_STEP per step
    assert rel1(out["vector"], g["vector"]) <= 1e-10                        # north-star tolerance
    assert [out["stats"][k] for k in STAT_KEYS] == list(g["stats"])
    h.close()
