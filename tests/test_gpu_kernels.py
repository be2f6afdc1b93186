"""GPU parity of the numerical kernels against the oracle."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import expm

import oracle
from gpu_common import make, rel1

pytestmark = pytest.mark.gpu


def grown(name, steps):
    h, om, x0 = make(name)
    of = oracle.Fsp(om, reproducible=1)          # canonical arithmetic: the device must match bit for bit
    h.fsp_init([x0]); of.set_states([x0]); of.matrix_starter()
    for _ in range(steps):
        h.onestep(); of.onestep()
    return h, of


@pytest.mark.parametrize("name,steps", [("toggle", 30), ("repressilator", 12), ("goutsias", 8), ("birth_death", 40)])
def test_fmatvec(name, steps):
    h, of = grown(name, steps)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(of.size)
    y, yo = h.matvec(x), of.matvec(x)
    assert np.array_equal(y, yo)                                       # canonical order: bit-exact
    lib = oracle.lib()
    lib.ko_fsp_set_reproducible(of.h, 0)                               # the reference's scatter order: rounding level
    assert rel1(y, of.matvec(x)) < 1e-14
    lib.ko_fsp_set_reproducible(of.h, 1)
    # column sums <= 0: e^T A x for x >= 0 must not be positive
    assert h.matvec(np.abs(x)).sum() <= 1e-9
    h.close()


@pytest.mark.parametrize("name,steps,m", [("toggle", 40, 10), ("toggle", 40, 30), ("goutsias", 8, 20)])
def test_arnoldi_hessenberg(name, steps, m):
    h, of = grown(name, steps)
    n = of.size
    v = np.zeros(n); v[0] = 1.0; v[1:5] = 0.3
    H, av, brk, _ = h.arnoldi(v, m)
    ref = oracle.arnoldi_sweep(of, v, m)                               # canonical sweep on the un-normalised basis
    assert brk == 0 and ref["brk"] == 0
    assert np.array_equal(H, ref["H"])                                 # every Hessenberg entry bit-exact
    assert av == ref["avnorm"]
    # against the reference's own operation order (netlib BLAS order, DSCAL as a pass): rounding level
    lib = oracle.lib()
    lib.ko_fsp_set_reproducible(of.h, 0)
    net = oracle.arnoldi_sweep(of, v, m)
    lib.ko_fsp_set_reproducible(of.h, 1)
    assert np.abs(H - net["H"]).max() <= 1e-12 * np.abs(net["H"]).max()
    # IOP-2: H is tridiagonal apart from the unit entry
    assert np.count_nonzero(np.triu(H[:m, :m], 2)) == 0
    h.close()


def test_happy_breakdown():
    h, om, _ = make("birth_death")
    st = [[i] for i in range(6)]
    h.fsp_init(st)
    v = np.zeros(6); v[0] = 1.0
    H, av, brk, _ = h.arnoldi(v, 10)
    assert 1 <= brk <= 6
    h.close()


@pytest.mark.parametrize("n", [3, 12, 13, 32, 50, 77, 101, 102])
@pytest.mark.parametrize("t", [1e-3, 0.3, 4.0])
def test_expm_single_cta(n, t):
    h, _, _ = make("birth_death", max_states=1000)
    rng = np.random.default_rng(n)
    H = np.zeros((n, n))
    for j in range(max(n - 2, 1)):
        H[j, j] = -abs(rng.standard_normal()) * 5
        if j + 1 < n:
            H[j + 1, j] = abs(rng.standard_normal()) * 5
        if j > 0:
            H[j - 1, j] = rng.standard_normal() * 5
    if n >= 2:
        H[n - 1, n - 2] = 1.0
    E, ns, hn = h.expm(H, t)
    Er, nsr, hnr = oracle.dgpadm(H, t, reproducible=1)
    assert ns == nsr and hn == hnr
    assert np.array_equal(E, Er)                                       # canonical operation order: bit-exact
    Eo, nso, hno = oracle.dgpadm(H, t)                                 # netlib-order restatement of dgpadm.f
    assert ns == nso and abs(hn - hno) <= 1e-15 * hno
    assert np.abs(E - Eo).max() <= 1e-12 * max(1.0, np.abs(Eo).max())
    assert np.abs(E - expm(t * H)).max() <= 1e-11 * max(1.0, np.abs(Eo).max())
    h.close()


def test_expm_dense_and_leading_block():
    h, _, _ = make("birth_death", max_states=1000)
    rng = np.random.default_rng(2)
    H = rng.standard_normal((40, 40))
    E, _, _ = h.expm(H, 0.2)
    assert np.abs(E - expm(0.2 * H)).max() < 1e-11
    E2, _, _ = h.expm(H, 0.2, m=39)
    assert np.abs(E2 - expm(0.2 * H[:39, :39])).max() < 1e-11
    import krylovfspssa_b200 as k
    with pytest.raises(k.KfspError) as e:
        h.expm(np.zeros((5, 5)), 1.0)
    assert e.value.status == -4
    h.close()


@pytest.mark.parametrize("n,mx", [(1, 3), (1000, 12), (70001, 31), (5000, 102)])
def test_combine(n, mx):
    h, _, _ = make("birth_death", max_states=1000)
    rng = np.random.default_rng(n + mx)
    V = rng.standard_normal((n, mx))
    e = rng.standard_normal(mx)
    w, ws = h.combine(V, e, 0.7)
    wo = np.maximum(V @ (0.7 * e), 0.0)
    assert np.abs(w - wo).max() <= 1e-13 * max(1.0, np.abs(wo).max())
    assert abs(ws - wo.sum()) <= 1e-12 * max(1.0, wo.sum())
    assert w.min() >= 0.0
    # SURVEY a6/a7 against the oracle, bit for bit, with the per-column scales of the un-normalised basis
    cs = 1.0 / np.abs(rng.standard_normal(mx) * 3.0 + 0.1)
    for scale in (None, cs):
        w, ws, ssq = h.combine(V, e, 0.7, colscale=scale, with_ssq=True)
        wr, wsr, ssqr = oracle.combine_reproducible(V, e, 0.7, colscale=scale)
        assert np.array_equal(w, wr)
        assert ws == wsr and ssq == ssqr
    h.close()
