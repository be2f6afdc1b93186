"""Matrix-free lattice SpMV (spmv_variant = 1, csrc/lattice.cuh) against the explicit gather-ELL path and
the CPU oracle: FMATVEC recomputed from the integer state must be BIT-IDENTICAL to FMATVEC on the stored
ADJ/OFFDIAG/DIAG (KrylovSolver.f90:577-607), for any box shape, any number of species, and reactions that
change several species at once."""
import itertools
import os

import numpy as np
import pytest

import krylovfspssa_b200 as k
import oracle

pytestmark = pytest.mark.gpu
FIXED = dict(n_init_onestep=0, enable_drop=0, enable_expand=0, m_max=30, m_min=10)


def lattice_states(bounds):
    """Natural order: first species fastest."""
    grids = np.meshgrid(*[np.arange(b, dtype=np.int32) for b in reversed(bounds)], indexing="ij")
    return np.stack([g.ravel() for g in reversed(grids)], axis=1).astype(np.int32)


def file_model(name, params):
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), name))
    m.reset_parameters(params)
    return m


def conversion_model():
    """0->X, X->Y (changes two species), Y->0, 2Y->X: stoichiometry with diagonal and +-2 moves."""
    m = k.CME_MODEL().create(2, 4, 4)
    m.stoichiometry = np.array([[1, -1, 0, 1], [0, 1, -1, -2]], dtype=np.int32)
    m.reset_parameters([30.0, 0.7, 1.3, 0.01])
    for r, e in enumerate(["p1", "p2*X1", "p3*X2", "p4*X2*(X2-1)/2"], start=1):   # default names of CREATE: X1.., p1..
        m.set_propensity(r, e)
    m.loaded = True
    return m


def pair(model, bounds, **kw):
    st = lattice_states(bounds)
    n = len(st)
    a = k.KrylovFspHandle(model, max_states=n + 64, spmv_variant=0, **FIXED, **kw)
    a.fsp_init(st)
    b = k.KrylovFspHandle(model, max_states=n + 64, spmv_variant=1, **FIXED, **kw)
    b.fsp_init_box(bounds)
    return a, b, st


CASES = [
    ("toggle_test", lambda: file_model("toggle_test.input", [5000.0, 1600.0, 1.0, 1.0]), (37, 29)),
    ("toggle_test_wide", lambda: file_model("toggle_test.input", [50.0, 16.0, 1.0, 1.0]), (700, 23)),
    ("toggle", lambda: file_model("toggle.input", [1.0, 100.0, 1.0, 1.0, 100.0, 1.0]), (300, 41)),
    ("repressilator3d", lambda: file_model("repressilator.input", [100.0, 100.0, 100.0, 1.0, 1.0, 1.0]), (13, 9, 11)),
    ("conversion", conversion_model, (45, 31)),
]


@pytest.mark.parametrize("name,mk,bounds", CASES, ids=[c[0] for c in CASES])
def test_lattice_matches_explicit(name, mk, bounds):
    model = mk()
    a, b, st = pair(model, bounds)
    n = len(st)
    assert a.size == b.size == n
    ga, gb = a.get(), b.get()
    for key in ("states", "adj", "offdiag", "diag"):
        assert np.array_equal(ga[key], gb[key]), key            # the recomputed column form IS the stored one
    rng = np.random.default_rng(5)
    for rep in range(3):
        x = rng.standard_normal(n) if rep else np.abs(rng.standard_normal(n))
        assert np.array_equal(a.matvec(x), b.matvec(x))         # FMATVEC bit-identical
    v = np.abs(rng.standard_normal(n))
    Ha, ava, brka, _ = a.arnoldi(v, 12)
    Hb, avb, brkb, _ = b.arnoldi(v, 12)
    assert np.array_equal(Ha, Hb) and ava == avb and brka == brkb
    q = np.array([st[0], st[n // 2], st[-1], [bounds[0], 0] + [0] * (len(bounds) - 2)], dtype=np.int32)
    assert np.array_equal(a.index(q), b.index(q))
    a.close(); b.close()


def test_lattice_fmatvec_matches_oracle():
    model = file_model("toggle_test.input", [5000.0, 1600.0, 1.0, 1.0])
    bounds = (61, 47)
    st = lattice_states(bounds)
    b = k.KrylovFspHandle(model, max_states=len(st) + 64, spmv_variant=1, **FIXED)
    b.fsp_init_box(bounds)
    om = oracle.Model.load(os.path.join(k.models_dir(), "toggle_test.input"), [5000.0, 1600.0, 1.0, 1.0])
    of = oracle.Fsp(om, reproducible=1)
    of.set_states(st); of.matrix_starter()
    x = np.random.default_rng(2).standard_normal(len(st))
    assert np.array_equal(b.matvec(x), of.matvec(x))
    b.close()


@pytest.mark.parametrize("bounds,t", [((301, 257), 0.05), ((1200, 800), 0.03)])
def test_lattice_solve_bit_identical(bounds, t):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    model = file_model("toggle_test.input", bench.PARAMS)
    states, p0 = bench.synthetic(*bounds)
    assert np.array_equal(states, lattice_states(bounds))
    a, b, _ = pair(model, bounds)
    a.set_vector(p0); b.set_vector(p0)
    rca, sa = a.solve_resident(t, 1e-6, 1e-8)
    rcb, sb = b.solve_resident(t, 1e-6, 1e-8)
    assert rca == rcb == 0
    ta, tb = a.trace(), b.trace()
    assert np.array_equal(ta["i"], tb["i"]) and np.array_equal(ta["d"], tb["d"])
    assert np.array_equal(a.get(matrix=False)["vector"], b.get(matrix=False)["vector"])
    for key in ("nmult", "nexph", "nscale", "nstep", "nreject"):
        assert sa[key] == sb[key]
    # the reference-facing call: explicit state list in, detected as a lattice on the device
    c = k.KrylovFspHandle(model, max_states=len(p0) + 64, spmv_variant=1, **FIXED)
    out = c.solve(t, states, p0, 1e-6, 1e-8)
    assert np.array_equal(out["states"], states)
    assert np.array_equal(out["vector"], b.get(matrix=False)["vector"])
    a.close(); b.close(); c.close()


def test_lattice_rejects_what_it_cannot_do():
    model = file_model("toggle_test.input", [5000.0, 1600.0, 1.0, 1.0])
    h = k.KrylovFspHandle(model, max_states=10000, spmv_variant=1, **FIXED)
    st = lattice_states((20, 10))
    with pytest.raises(k.KfspError):                      # not a full box in natural order
        h.fsp_init(st[::-1].copy())
    with pytest.raises(k.KfspError):
        h.fsp_init(st[:-3])
    p0 = np.zeros(len(st)); p0[0] = 1.0
    with pytest.raises(k.KfspError):                      # kfsp_solve verifies the list beside the solve: same verdict
        h.solve(0.01, st[::-1].copy(), p0, 1e-6, 1e-8)
    h.fsp_init(st)                                        # a lattice given as a list is accepted
    with pytest.raises(k.KfspError):
        h.onestep()                                       # fixed state set
    h.close()
    g = file_model("goutsias.input", [0.1] * 10)          # c5*DNA*D reads two species
    h2 = k.KrylovFspHandle(g, max_states=100000, spmv_variant=1, **FIXED)
    with pytest.raises(k.KfspError):
        h2.fsp_init_box((3, 3, 3, 3, 3, 3))
    h2.close()
    with pytest.raises(k.KfspError):                      # adaptivity must be off
        h3 = k.KrylovFspHandle(model, max_states=10000, spmv_variant=1)
        h3.fsp_init_box((20, 10))


def role_model(order, mask):
    """2-species model of four one-molecule reactions in reaction order `order` (0: X+,X-,Y+,Y-; 1: X+,Y+,X-,Y-) whose
    k-th propensity reads Y if bit k of `mask` is set, else X: every instantiation of the stencil kernel k_spmv_bd2."""
    steps = {0: [(1, 0), (-1, 0), (0, 1), (0, -1)], 1: [(1, 0), (0, 1), (-1, 0), (0, -1)]}[order]
    m = k.CME_MODEL().create(2, 4, 4)
    m.stoichiometry = np.array(steps, dtype=np.int32).T
    m.reset_parameters([3.0, 0.7, 2.5, 1.1])
    for r in range(4):
        v = "X2" if (mask >> r) & 1 else "X1"
        # production terms are Hill-like (transcendental: host-tabulated), degradation terms linear
        expr = "p%d/(1.0+%s^1.5)" % (r + 1, v) if steps[r][0] + steps[r][1] > 0 else "p%d*%s" % (r + 1, v)
        m.set_propensity(r + 1, expr)
    m.loaded = True
    return m


@pytest.mark.parametrize("order", [0, 1])
def test_stencil_kernel_every_table_role(order):
    bounds = (67, 35)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(bounds[0] * bounds[1])
    v = np.abs(rng.standard_normal(bounds[0] * bounds[1]))
    for mask in range(16):
        a, b, st = pair(role_model(order, mask), bounds)
        assert np.array_equal(a.matvec(x), b.matvec(x)), (order, mask)           # plain SpMV
        Ha, ava, _, _ = a.arnoldi(v, 10)
        Hb, avb, _, _ = b.arnoldi(v, 10)
        assert np.array_equal(Ha, Hb) and ava == avb, (order, mask)              # dot-fused and norm-fused variants
        a.close(); b.close()


@pytest.mark.parametrize("zc", [1, 5, 9, 17, 64, 256])
def test_stencil_kernel_chunk_lengths(zc, monkeypatch):
    """The stencil kernel walks a z-chunk in three pieces (head at the box boundary, a hot loop without bounds checks, a
    general tail).  At test sizes the automatic chunking gives chunks shorter than the hot loop needs, so the chunk length
    is forced (KFSP_BD2_ZC) to run every piece: plain, dot-fused, finalising (FIN) and norm variants against the explicit path."""
    monkeypatch.setenv("KFSP_BD2_ZC", str(zc))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    bounds = (301, 257)
    model = file_model("toggle_test.input", bench.PARAMS)
    states, p0 = bench.synthetic(*bounds)
    a, b, _ = pair(model, bounds)
    rng = np.random.default_rng(zc)
    x = rng.standard_normal(len(p0))
    assert np.array_equal(a.matvec(x), b.matvec(x))
    v = np.abs(rng.standard_normal(len(p0)))
    Ha, ava, brka, _ = a.arnoldi(v, 12)
    Hb, avb, brkb, _ = b.arnoldi(v, 12)
    assert np.array_equal(Ha, Hb) and ava == avb and brka == brkb
    a.set_vector(p0); b.set_vector(p0)
    rca, sa = a.solve_resident(0.05, 1e-6, 1e-8)
    rcb, sb = b.solve_resident(0.05, 1e-6, 1e-8)
    assert rca == rcb == 0
    assert np.array_equal(a.trace()["i"], b.trace()["i"])
    assert np.array_equal(a.get(matrix=False)["vector"], b.get(matrix=False)["vector"])
    a.close(); b.close()
