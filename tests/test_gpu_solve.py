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


def test_repeated_solves_are_reproducible():
    h, om, x0 = make("toggle")
    a = h.solve(3.0, [x0], [1.0], 1e-4, 1e-10)
    b = h.solve(3.0, [x0], [1.0], 1e-4, 1e-10)          # same handle: SSA streams restart with the state space
    assert np.array_equal(a["states"], b["states"]) and np.array_equal(a["vector"], b["vector"])
    assert np.array_equal(a["trace"]["i"], b["trace"]["i"])
    ph = h.phase_seconds()
    assert {"sweep_pade", "combine_norms", "ssa", "drop", "onestep", "host_callbacks"} <= set(ph) and ph["sweep_pade"] > 0
    assert ph["host_callbacks"] == 0 and ph["host_propensity_evals"] == 0        # byte-code model: nothing runs on the host
    h.close()


def test_chained_solves_equal_oracle_chain():
    """CME_SOLVE output fed back as input (FSP_OUT -> next call).  The FSP criterion WSUM >= 1 - FSPTOL*t/T
    (KrylovSolver.f90:458) assumes unit initial mass, so the caller renormalises between calls; an
    un-normalised restart can never satisfy it and ends in KFSP_ERR_MOLECULE_LIMIT / overflow instead of
    looping forever like the reference would."""
    h, om, x0 = make("birth_death")
    a = h.solve(1.0, [x0], [1.0], 1e-6, 1e-10)
    p1 = a["vector"] / a["vector"].sum()
    b = h.solve(1.0, a["states"], p1, 1e-6, 1e-10)
    ra = oracle.solve(om, [x0], [1.0], 1.0, 1e-6, 1e-10, reproducible=1)
    assert np.array_equal(a["vector"], ra["vector"])
    rb = oracle.solve(om, ra["states"], ra["vector"] / ra["vector"].sum(), 1.0, 1e-6, 1e-10, reproducible=1)
    assert np.array_equal(b["states"], rb["states"]) and np.array_equal(b["vector"], rb["vector"])
    h.close()


def test_python_callable_customprop_solves_like_the_file_model():
    """MODEL%CUSTOMPROP => a Python callable (ModelModule.f90:6-12,188-189): host callbacks in batches and through
    the SSA side cache give the same bits as the byte-code model with the same arithmetic (birth-death)."""
    model = k.CME_MODEL().create(1, 2, 2)
    model.stoichiometry = [[1, -1]]
    model.reset_parameters([20.0, 1.0])
    model.set_customprop(lambda st, r, p: p[0] if r == 1 else p[1] * st[0])
    assert model.propensity([3], 2) == 3.0
    h = k.KrylovFspHandle(model, max_states=5000)
    out = h.solve(2.0, [[0]], [1.0], 1e-6, 1e-10)
    hf, om, x0 = make("birth_death", max_states=5000)
    ref = hf.solve(2.0, [x0], [1.0], 1e-6, 1e-10)
    assert out["iflag"] == ref["iflag"] == 0
    assert np.array_equal(out["states"], ref["states"]) and np.array_equal(out["vector"], ref["vector"])
    assert np.array_equal(out["trace"]["i"], ref["trace"]["i"])
    h.close(); hf.close()


def test_small_krylov_range_option():
    h, om, x0 = make("toggle", m_max=20, m_min=5)
    out = h.solve(1.0, [x0], [1.0], 1e-4, 1e-8)
    ref = oracle.solve(om, [x0], [1.0], 1.0, 1e-4, 1e-8, m_max=20, m_min=5, reproducible=1)
    assert np.array_equal(out["states"], ref["states"]) and np.array_equal(out["vector"], ref["vector"])
    assert out["trace"]["i"][:, 0].max() <= 20
    h.close()


def _birth_death(params, **kw):
    path = os.path.join(k.models_dir(), "birth_death.input")
    model = k.CME_MODEL().load(path)
    model.reset_parameters(params)
    return k.KrylovFspHandle(model, max_states=100000, seed=1, **kw), oracle.Model.load(path, params)


def test_zero_horizon_returns_the_start_up_expansion_only():
    """T = 0: the stepping loop of DGEXPV_FSP never runs (KrylovSolver.f90:199); what comes back is the five start-up
    ONESTEP_EXTENDER rounds (:173-178) around x0 with all the mass still on x0"""
    h, om = _birth_death([20.0, 1.0])
    out = h.solve(0.0, [[0]], [1.0], 1e-6, 1e-10)
    ref = oracle.solve(om, [[0]], [1.0], 0.0, 1e-6, 1e-10, seed=1, reproducible=1)
    assert out["iflag"] == ref["iflag"] == 0 and out["stats"]["nstep"] == 0
    assert np.array_equal(out["states"], ref["states"]) and np.array_equal(out["vector"], ref["vector"])
    assert out["vector"][0] == 1.0 and out["vector"][1:].sum() == 0.0
    h.close()


def test_absorbing_start_state_is_the_reference_error():
    """no reaction can fire from x0 (k = 0 at X = 0): the generator on the projection is the zero matrix and DGPADM stops with
    'NULL H IN INPUT OF DGPADM' (dgpadm.f:254); the C ABI returns KFSP_ERR_NULL_H where the reference STOPs, and the handle
    stays usable"""
    h, om = _birth_death([0.0, 1.0])
    ref = oracle.solve(om, [[0]], [1.0], 1.0, 1e-6, 1e-10, seed=1, reproducible=1)
    assert ref["iflag"] == -4
    with pytest.raises(k.KfspError) as e:
        h.solve(1.0, [[0]], [1.0], 1e-6, 1e-10)
    assert e.value.status == -4
    out = h.solve(1.0, [[5]], [1.0], 1e-6, 1e-10)                 # pure death from 5: fine, and bit-identical
    ref = oracle.solve(om, [[5]], [1.0], 1.0, 1e-6, 1e-10, seed=1, reproducible=1)
    assert out["iflag"] == ref["iflag"] == 0
    assert np.array_equal(out["states"], ref["states"]) and np.array_equal(out["vector"], ref["vector"])
    h.close()


def test_state_space_overflow_mid_solve_and_recovery():
    """'FSP SIZE EXCEEDS MEMORY LIMIT' (StateSpace.f90:388-391) raised by an expansion in the middle of a solve comes back as
    KFSP_ERR_OVERFLOW; the same handle then solves a problem that fits"""
    path = os.path.join(k.models_dir(), "toggle.input")
    params = [1.0, 100.0, 1.0, 1.0, 100.0, 1.0]
    model = k.CME_MODEL().load(path)
    model.reset_parameters(params)
    h = k.KrylovFspHandle(model, max_states=300, seed=12345)
    with pytest.raises(k.KfspError) as e:
        h.solve(20.0, [[0, 0]], [1.0], 1e-4, 1e-10)
    assert e.value.status == -10
    out = h.solve(0.05, [[0, 0]], [1.0], 1e-4, 1e-10)
    ref = oracle.solve(oracle.Model.load(path, params), [[0, 0]], [1.0], 0.05, 1e-4, 1e-10, seed=12345, reproducible=1, max_size=300)
    assert out["iflag"] == ref["iflag"] == 0
    assert np.array_equal(out["states"], ref["states"]) and np.array_equal(out["vector"], ref["vector"])
    h.close()


def test_molecule_limit_is_an_error_not_a_key_collision():
    """a count above MAXNUMBERMOLECULES aliases hash keys in the reference (HashTable: key arithmetic in base MAXNUMBERMOLECULES);
    here it is reported (documented deviation)"""
    path = os.path.join(k.models_dir(), "birth_death.input")
    model = k.CME_MODEL().load(path)
    model.reset_parameters([20.0, 1.0])
    h = k.KrylovFspHandle(model, max_states=1000, max_molecules=50)
    with pytest.raises(k.KfspError) as e:
        h.fsp_init([[51]])
    assert e.value.status in (-11, -12)
    h.fsp_init([[48]])
    with pytest.raises(k.KfspError) as e:
        for _ in range(5):
            h.onestep()
    assert e.value.status == -12
    h.close()
