"""Pins the oracle's DGEXPV_FSP / state-space restatement (oracle/kfsp_oracle.cpp)
by analytic known answers and the invariants the reference prints."""
import math
import os

import numpy as np
import pytest

import oracle


def poisson_pmf(mean, n):
    out = np.zeros(n)
    out[0] = math.exp(-mean)
    for k in range(1, n):
        out[k] = out[k - 1] * mean / k
    return out


def test_birth_death_poisson(models_dir):
    k, g, t = 20.0, 1.0, 2.0
    m = oracle.Model.load(os.path.join(models_dir, "birth_death.input"), [k, g])
    out = oracle.solve(m, [[0]], [1.0], t, 1e-6, 1e-10)
    assert out["iflag"] == 0
    x = out["states"][:, 0]
    p = out["vector"]
    pm = poisson_pmf(k / g * (1 - math.exp(-g * t)), 200)
    assert np.abs(p - pm[x]).sum() < 5e-6
    assert p.min() >= 0.0 and 1 - 1e-6 <= p.sum() <= 1 + 1e-12


def test_toggle_invariants_and_trace(models_dir):
    m = oracle.Model.load(os.path.join(models_dir, "toggle.input"), [1, 100, 1, 1, 100, 1])
    T, ftol = 5.0, 1e-4
    out = oracle.solve(m, [[0, 0]], [1.0], T, ftol, 1e-10)
    assert out["iflag"] == 0
    td, ti = out["trace_d"], out["trace_i"]
    assert abs(td[-1, 0] - T) < 1e-12
    acc = ti[:, 3] & 4 == 0                       # rows that advanced time
    for tn, ws in zip(td[acc, 0], td[acc, 3]):
        assert 1 - ftol * tn / T - 1e-15 <= ws <= 1 + 1e-12
    assert out["vector"].min() >= 0
    st = out["states"]
    assert len({tuple(s) for s in st}) == len(st)          # no duplicates
    assert st.min() >= 0


def test_matrix_invariant_adj_consistent(models_dir):
    """After any expansion ADJ(K,J) is: index of x_J+nu_K if present, 0 if legal and absent,
    -1 if it has a negative component (the invariant the device rebuild relies on)."""
    m = oracle.Model.load(os.path.join(models_dir, "toggle.input"), [1, 100, 1, 1, 100, 1])
    f = oracle.Fsp(m)
    f.set_states([[0, 0]])
    assert f.matrix_starter() == 0
    for _ in range(5):
        assert f.onestep() == 0
    assert f.size == 21                                     # 5-step reachable set
    rng = oracle.Rng(oracle.Rng.PHILOX, 7)
    f.ssa(0.5, rng)
    f.onestep()
    d = f.get()
    idx = {tuple(s): i + 1 for i, s in enumerate(d["states"])}
    for j, s in enumerate(d["states"]):
        for k in range(m.R):
            nxt = s + m.stoich[k]
            want = -1 if nxt.min() < 0 else idx.get(tuple(nxt), 0)
            assert d["adj"][j, k] == want
            assert d["offdiag"][j, k] == m.propensity(s, k + 1)
        assert d["diag"][j] == sum_in_order(d["offdiag"][j])
        assert f.index(s) == j + 1
    # column sums of the FSP generator are <= 0 (mass only leaks out)
    y = f.matvec(np.ones(f.size))
    x = np.zeros(f.size); x[3] = 1.0
    assert f.matvec(x).sum() <= 1e-12
    assert y.shape == (f.size,)


def sum_in_order(v):
    s = 0.0
    for a in v:
        s = s + a
    return s


def test_repressilator_initial_reachable_set(models_dir):
    m = oracle.Model.load(os.path.join(models_dir, "repressilator.input"), [100, 100, 100, 1, 1, 1])
    f = oracle.Fsp(m)
    f.set_states([[22, 0, 0]])
    f.matrix_starter()
    for _ in range(5):
        f.onestep()
    assert f.size == 91


def test_philox_stream_is_counter_based():
    r = oracle.Rng(oracle.Rng.PHILOX, 99)
    a = r.draw2(5, 0, 1)
    b = r.draw2(5, 1, 1)
    assert a == r.draw2(5, 0, 1) and a != b
    assert all(0.0 <= u < 1.0 for u in a + b)
    assert oracle.Rng(oracle.Rng.PHILOX, 100).draw2(5, 0, 1) != a


def test_philox_known_answer():
    # Random123 KAT for philox4x32-10: counter=key=0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    r = oracle.Rng(oracle.Rng.PHILOX, 0)
    u1, u2 = r.draw2(0, 0, 0)
    a = (0xe169c58d << 32) | 0x6627e8d5
    b = (0x9b00dbd8 << 32) | 0xbc57ac4c
    assert u1 == (a >> 11) * 2.0 ** -53 and u2 == (b >> 11) * 2.0 ** -53


def test_gfortran_stream_if_available():
    if oracle.find_libgfortran() is None:
        pytest.skip("no libgfortran in this image")
    try:
        r = oracle.Rng(oracle.Rng.GFORTRAN, 1)
    except RuntimeError:
        pytest.skip("libgfortran seeding interface not usable")
    a, b = r.draw2(1, 0, 0)
    assert 0.0 <= a < 1.0 and 0.0 <= b < 1.0 and a != b


def test_drop_states_semantics(models_dir):
    m = oracle.Model.load(os.path.join(models_dir, "birth_death.input"), [5.0, 1.0])
    f = oracle.Fsp(m)
    f.set_states([[i] for i in range(40)])
    f.matrix_starter()
    w = oracle.solve(m, [[i] for i in range(40)], np.eye(40)[0], 1.0, 1e-6, 1e-10,
                     n_init_onestep=0, enable_drop=0, enable_expand=0)["vector"]
    did, w2, tol, cnt = f.drop(w.copy(), 1e-7)
    assert did == 1 and f.size < 40
    d = f.get()
    assert np.array_equal(d["states"][:, 0], np.arange(f.size))     # stable compaction keeps order
    assert d["adj"][f.size - 1, 0] == 0                             # dropped target becomes explorable
    assert np.array_equal(w2, w[: f.size])


def cme_on_a_big_box(model, bounds, x0, t):
    """Independent solution of dp/dt = A p: the CME generator on a box far larger than the support, assembled from
    MODEL%PROPENSITY, integrated by scipy.sparse.linalg.expm_multiply (no Krylov-FSP code involved)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import expm_multiply
    S, R = model.S, model.R
    grids = np.meshgrid(*[np.arange(b) for b in bounds], indexing="ij")
    states = np.stack([g.ravel() for g in grids], axis=1).astype(np.int32)
    strides = np.array([int(np.prod(bounds[s + 1:])) for s in range(S)])
    n = len(states)
    rows, cols, vals = [], [], []
    diag = np.zeros(n)
    for k in range(R):
        a = np.array([model.propensity(s, k + 1) for s in states])
        nxt = states + model.stoich[k]
        inside = np.all((nxt >= 0) & (nxt < np.array(bounds)), axis=1)
        diag -= a
        rows.append((nxt[inside] * strides).sum(axis=1)); cols.append(np.nonzero(inside)[0]); vals.append(a[inside])
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)) + sp.diags(diag)
    p0 = np.zeros(n)
    p0[int((np.asarray(x0) * strides).sum())] = 1.0
    return states, strides, expm_multiply(A * t, p0)


@pytest.mark.parametrize("repro", [0, 1])
def test_toggle_against_an_independent_cme_solver(models_dir, repro):
    """The whole adaptive algorithm (Krylov steps + SSA/one-step expansion + pruning) against scipy on a big box: the
    FSP solution must be within FSPTOL (plus Krylov tolerance) of the CME solution in 1-norm -- the error bound the
    algorithm promises (KrylovSolver.f90:458), checked without any of its own code."""
    params, t, ftol = [1, 100, 1, 1, 100, 1], 2.0, 1e-4
    m = oracle.Model.load(os.path.join(models_dir, "toggle.input"), params)
    out = oracle.solve(m, [[0, 0]], [1.0], t, ftol, 1e-10, reproducible=repro)
    assert out["iflag"] == 0
    bounds = (int(out["states"][:, 0].max()) + 40, int(out["states"][:, 1].max()) + 40)
    states, strides, ptrue = cme_on_a_big_box(m, bounds, [0, 0], t)
    assert ptrue.sum() > 1 - 1e-9                                   # the box itself loses nothing
    idx = (out["states"] * strides).sum(axis=1)
    inside = np.zeros(len(ptrue), dtype=bool)
    inside[idx] = True
    err = np.abs(out["vector"] - ptrue[idx]).sum() + ptrue[~inside].sum()
    assert err <= 1.05 * ftol, err
    assert err > 1e-9                                               # and it IS a truncation, not the same computation


def test_repressilator_against_an_independent_cme_solver(models_dir):
    params, x0, t, ftol = [100.0, 100.0, 100.0, 1.0, 1.0, 1.0], [22, 0, 0], 0.3, 1e-4
    m = oracle.Model.load(os.path.join(models_dir, "repressilator.input"), params)
    out = oracle.solve(m, [x0], [1.0], t, ftol, 1e-10, reproducible=1)
    assert out["iflag"] == 0
    bounds = tuple(int(out["states"][:, s].max()) + 12 for s in range(3))
    states, strides, ptrue = cme_on_a_big_box(m, bounds, x0, t)
    assert ptrue.sum() > 1 - 1e-8
    idx = (out["states"] * strides).sum(axis=1)
    inside = np.zeros(len(ptrue), dtype=bool)
    inside[idx] = True
    err = np.abs(out["vector"] - ptrue[idx]).sum() + ptrue[~inside].sum()
    assert 1e-9 < err <= 1.05 * ftol, err
