"""Index-only generator SpMV (spmv_variant = 2, csrc/krylov.cuh: k_spmv_idx): a_k(x - nu_k) is recomputed from the row's
integer state through the factored propensity tables instead of being streamed (FMATVEC, KrylovSolver.f90:577-607;
OFFDIAG(K,J) = a_K of state J, StateSpace.f90:207-212).  Everything must stay bit-identical: against the oracle, against
the explicit matrix, on grown, dropped and adaptively expanded state sets, and at the full horizons of BASELINE configs 1-3."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from gpu_common import GOLDEN_RUNS, make

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def grown(name, steps, **opt):
    h, om, x0 = make(name, **opt)
    of = oracle.Fsp(om, reproducible=1)
    h.fsp_init([x0]); of.set_states([x0]); of.matrix_starter()
    for _ in range(steps):
        h.onestep(); of.onestep()
    return h, of


@pytest.mark.parametrize("name,steps", [("toggle", 30), ("repressilator", 12), ("goutsias", 8), ("birth_death", 40), ("toggle_test", 25)])
def test_fmatvec_index_only(name, steps):
    h, of = grown(name, steps, spmv_variant=2)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(of.size)
    assert np.array_equal(h.matvec(x), of.matvec(x))             # bit-exact against the oracle's canonical order
    h.close()


@pytest.mark.parametrize("name,steps,m", [("toggle", 40, 10), ("goutsias", 8, 20), ("repressilator", 14, 30)])
def test_arnoldi_index_only(name, steps, m):
    h, of = grown(name, steps, spmv_variant=2)
    n = of.size
    v = np.zeros(n); v[0] = 1.0; v[1:5] = 0.3
    H, av, brk, _ = h.arnoldi(v, m)
    ref = oracle.arnoldi_sweep(of, v, m)
    assert brk == 0 and ref["brk"] == 0
    assert np.array_equal(H, ref["H"])
    assert av == ref["avnorm"]
    h.close()


@pytest.mark.parametrize("small_sweep", ["1", "0"])
@pytest.mark.parametrize("tag", sorted(GOLDEN_RUNS))
def test_solve_index_only_matches_golden(tag, small_sweep, monkeypatch):
    """adaptive solves (expansion by SSA + one-step, drops): state list, decision trace and vector equal the committed oracle
    fixtures bit for bit, through the single-CTA sweep and through the multi-launch sweep"""
    monkeypatch.setenv("KFSP_SMALL_SWEEP", small_sweep)
    name, t, ftol, ktol, seed = GOLDEN_RUNS[tag]
    g = np.load(os.path.join(HERE, "golden", tag + ".npz"))
    h, _, x0 = make(name, seed=seed, spmv_variant=2)
    out = h.solve(t, [x0], [1.0], ftol, ktol)
    assert out["iflag"] == 0
    assert np.array_equal(out["states"], g["states"])
    assert np.array_equal(out["vector"], g["vector"])
    h.close()


def test_index_only_rejects_what_it_cannot_factor():
    """a propensity that divides by a sub-expression of two species has no factored form: KFSP_ERR_UNSUPPORTED at set_model,
    never a silent switch of variant; CUSTOMPROP host callbacks likewise"""
    import krylovfspssa_b200 as k
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle.input"))
    m.reset_parameters([1.0, 100.0, 1.0, 1.0, 100.0, 1.0])
    m.set_propensity(1, "kx/(1.0 + X*Y)")
    with pytest.raises(k.KfspError):
        k.KrylovFspHandle(m, max_states=1000, spmv_variant=2)
    m.set_propensity(1, "kx*X*Y + 0.5*Y")                  # sums and products of several species: fine
    h = k.KrylovFspHandle(m, max_states=1000, spmv_variant=2)
    h0 = k.KrylovFspHandle(m, max_states=1000, spmv_variant=0)
    for hh in (h, h0):
        hh.fsp_init([[1, 2]])
        for _ in range(6):
            hh.onestep()
    x = np.random.default_rng(3).standard_normal(h.size)
    assert np.array_equal(h.matvec(x), h0.matvec(x))
    h.close(); h0.close()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag,case", [("toggle_full", ("toggle", 1000.0, 1e-4, 1e-10, 400000)),
                                      ("repressilator_full", ("repressilator", 10.0, 1e-4, 1e-10, 2000000)),
                                      ("goutsias_full", ("goutsias", 300.0, 1e-6, 1e-8, 6291469))])
def test_full_config_index_only(tag, case):
    db = json.load(open(os.path.join(HERE, "golden", "full_digests.json")))
    if tag not in db:
        pytest.skip("no digest committed for " + tag)
    g = db[tag]
    name, t, ftol, ktol, cap = case
    h, _, x0 = make(name, max_states=cap, seed=12345, spmv_variant=2)
    out = h.solve(t, [x0], [1.0], ftol, ktol)
    st = out["stats"]
    print("%s (index-only): N=%d steps=%d nmult=%d device %.2f s" % (tag, len(out["vector"]), st["nstep"], st["nmult"], st["device_seconds"]))
    assert out["iflag"] == 0 and len(out["vector"]) == g["n"]
    assert sha(out["states"].astype(np.int32)) == g["states_sha256"]
    assert sha(out["vector"].astype(np.float64)) == g["vector_sha256"]
    h.close()
