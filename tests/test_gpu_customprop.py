"""CUSTOMPROP on the device path (BASELINE config 4: examples/transcr6d.f90, and the other two example
drivers): the propensity is an opaque HOST function, evaluated in batches for MATRIX_STARTER /
ONESTEP_EXTENDER and served to the device SSA walks from a side cache filled in rounds.  Decision
trace, state list and probability vector must be bit-identical to the CPU oracle, which uses its own
restatement of the same Fortran functions."""
import numpy as np
import pytest

import krylovfspssa_b200 as k
import oracle
from krylovfspssa_b200 import examples

pytestmark = pytest.mark.gpu

# shortened horizons (the oracle must finish in seconds); full horizons: test_gpu_full_configs.py
SHORT_T = {"toggle": 10.0, "repressilator": 2.0, "transcr6d": 100.0}


def oracle_model(name):
    d = examples.DRIVERS[name]
    om = oracle.Model(d["S"], d["R"], d["P"], d["stoich"], d["params"])
    om.set_custom(d["oracle_kind"])
    return om


# probe "1": the callback's structure is found by probing -- single-species reactions are tabulated, bilinear mass action
# (transcr6d, reactions 5 and 7) becomes byte code -- and the model is served by the device;
# probe "0": every callback stays a host function
PROBE_CASES = [(n, p) for n in sorted(examples.DRIVERS) for p in ("1", "0")]


@pytest.mark.parametrize("name,probe", PROBE_CASES)
def test_fsp_routines_with_host_callback(name, probe, monkeypatch):
    monkeypatch.setenv("KFSP_CUSTOM_PROBE", probe)
    d = examples.DRIVERS[name]
    h = k.KrylovFspHandle(examples.driver_model(name), max_states=400000, seed=777)
    f = oracle.Fsp(oracle_model(name), max_size=400000, reproducible=1)
    h.fsp_init([d["x0"]])
    f.set_states([d["x0"]])
    f.matrix_starter()
    rng = oracle.Rng(1, 777)
    for it in range(4):
        h.onestep()
        f.onestep()
    for it in range(3):
        ts = 0.05 * (it + 1) if name != "transcr6d" else 2.0 * (it + 1)
        h.ssa(ts)
        f.ssa(ts, rng)
        h.onestep()
        f.onestep()
        a, b = h.get(), f.get()
        assert h.size == f.size and h.size > 50
        for key in ("states", "adj", "offdiag", "diag"):
            assert np.array_equal(a[key], b[key]), (name, it, key)
    h.close()


@pytest.mark.parametrize("name,probe", PROBE_CASES)
def test_solve_with_host_callback_bit_identical(name, probe, monkeypatch):
    monkeypatch.setenv("KFSP_CUSTOM_PROBE", probe)
    d = examples.DRIVERS[name]
    t = SHORT_T[name]
    h = k.KrylovFspHandle(examples.driver_model(name), max_states=2000000, seed=12345)
    on_device = probe == "1"
    info = h.model_info()
    n_tab = d["R"] - (2 if name == "transcr6d" else 0)
    assert (info["n_tabulated"], info["n_host_evaluated"]) == ((n_tab, 0) if on_device else (0, d["R"]))
    out = h.solve(t, [d["x0"]], [1.0], d["fsp_tol"], d["exp_tol"])
    assert (h.phase_seconds()["host_propensity_evals"] == 0) == on_device
    ref = oracle.solve(oracle_model(name), [d["x0"]], [1.0], t, d["fsp_tol"], d["exp_tol"], seed=12345,
                       max_size=2000000, reproducible=1)
    assert out["iflag"] == 0 and ref["iflag"] == 0
    assert np.array_equal(out["trace"]["i"], ref["trace_i"])                 # decision trace
    assert np.array_equal(out["trace"]["d"][:, 1], ref["trace_d"][:, 1])     # step sizes
    assert np.array_equal(out["states"], ref["states"])                      # state set and indices bit-exact
    err = np.abs(out["vector"] - ref["vector"]).sum() / np.abs(ref["vector"]).sum()
    assert err <= 1e-10, err                                                 # north-star tolerance
    assert np.array_equal(out["vector"], ref["vector"])                      # in fact bit-identical
    assert out["stats"]["n_expand"] == ref["stats"]["n_expand"]
    h.close()


def test_cme_solve_driver_program():
    """examples/toggle.f90 as a program: CREATE, CUSTOMPROP =>, FSP%CREATE, CME_SOLVE."""
    d = examples.DRIVERS["toggle"]
    model = examples.driver_model("toggle")
    fsp_in, fsp = k.FINITE_STATE_PROJECTION(), k.FINITE_STATE_PROJECTION()
    fsp_in.create(model, 400000)
    fsp.create(model, 400000)
    fsp_in.set([d["x0"]], [1.0])
    fsp.set([d["x0"]], [1.0])
    k.CME_SOLVE(model, 5.0, fsp_in, fsp, d["fsp_tol"], d["exp_tol"], verbosity=0)
    w = fsp.vector[:fsp.size]
    assert (w >= 0).all() and 1.0 - d["fsp_tol"] <= w.sum() <= 1.0 + 1e-12
    top = int(np.argmax(w))
    assert fsp.index(fsp.state[:, top]) == top + 1                           # FSP%INDEX is 1-based
    assert fsp.probability(fsp.state[:, top]) == w[top]
    assert fsp.index([9999, 9999]) == 0


def test_multi_species_transcendental_program_is_evaluated_by_the_host(tmp_path):
    """a byte-code propensity with a transcendental operation on TWO species (Hill term times a linear one) cannot be tabulated;
    the CUDA math library may differ from the host libm by ulps, which could flip an SSA pick or a pruning decision, so the
    library evaluates it on the host through the CUSTOMPROP machinery (kfsp_model_info says so) and the adaptive solve stays
    bit-identical to the oracle"""
    import os
    import numpy as np
    import krylovfspssa_b200 as k
    import oracle
    src = open(os.path.join(k.models_dir(), "toggle.input")).read().splitlines()
    i = src.index("propensities")
    src[i + 1] = "(bx + kx/(2.0 + 0.2*Y^2.5))*(1.0 + 0.001*X)"
    path = tmp_path / "toggle_hill2.input"
    path.write_text("\n".join(src) + "\n")
    params = [1.0, 100.0, 1.0, 1.0, 100.0, 1.0]
    model = k.CME_MODEL().load(str(path))
    model.reset_parameters(params)
    h = k.KrylovFspHandle(model, max_states=200000, seed=7)
    info = h.model_info()
    assert info["n_host_evaluated"] == 1 and info["n_device_libm"] == 0
    out = h.solve(3.0, [[0, 0]], [1.0], 1e-4, 1e-10)
    ref = oracle.solve(oracle.Model.load(str(path), params), [[0, 0]], [1.0], 3.0, 1e-4, 1e-10, reproducible=1, seed=7)
    assert out["iflag"] == 0
    assert np.array_equal(out["states"], ref["states"])
    assert np.array_equal(out["vector"], ref["vector"])
    h.close()
    # the same program cannot take the index-only variant together with host evaluation: refused, not switched
    with pytest.raises(k.KfspError):
        k.KrylovFspHandle(model, max_states=1000, spmv_variant=2)


EXAMPLE_TOGGLE_AS_STRINGS = ["bx + kx/(1.0 + Y^1.5)", "dx*X", "by + ky/(1.0 + X^3.5)", "dy*Y"]     # examples/toggle.f90:60-74


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_customprop_on_every_spmv_variant(variant, tmp_path):
    """examples/toggle.f90's CUSTOMPROP model on a fixed 300x200 lattice through the explicit, the matrix-free lattice and the
    index-only SpMV: the callback is probed, tabulated and verified once, after which the handle is an ordinary tabulated
    model.  Same bits as the same propensities given as parser strings on the explicit path."""
    import os
    d = examples.DRIVERS["toggle"]
    src = open(os.path.join(k.models_dir(), "toggle.input")).read().splitlines()
    i = src.index("propensities")
    src[i + 1:i + 5] = EXAMPLE_TOGGLE_AS_STRINGS
    path = tmp_path / "toggle_example.input"
    path.write_text("\n".join(src) + "\n")
    parsed = k.CME_MODEL().load(str(path))
    parsed.reset_parameters(d["params"])
    bx, by = 300, 200
    yy, xx = np.meshgrid(np.arange(by, dtype=np.int32), np.arange(bx, dtype=np.int32), indexing="ij")
    states = np.stack([xx.ravel(), yy.ravel()], axis=1)
    p0 = np.zeros(bx * by)
    p0[(by // 3) * bx + bx // 4] = 1.0
    opts = dict(max_states=bx * by + 64, m_max=30, m_min=10, n_init_onestep=0, enable_drop=0, enable_expand=0)
    ref_h = k.KrylovFspHandle(parsed, spmv_variant=0, **opts)
    ref = ref_h.solve(0.5, states, p0, 1e-6, 1e-8)
    h = k.KrylovFspHandle(examples.driver_model("toggle"), spmv_variant=variant, **opts)
    assert h.model_info()["n_tabulated"] == 4
    out = h.solve(0.5, states, p0, 1e-6, 1e-8)
    assert out["iflag"] == 0 and ref["iflag"] == 0
    assert np.array_equal(out["trace"]["i"], ref["trace"]["i"])
    assert np.array_equal(out["vector"], ref["vector"])
    assert h.phase_seconds()["host_propensity_evals"] == 0
    h.close()
    ref_h.close()


def test_transcr6d_callback_on_the_index_only_variant():
    """examples/transcr6d.f90's CUSTOMPROP (two species in reactions 5 and 7, bilinear) through the index-only SpMV: possible
    because the probed model has a factored form; bit-identical to the oracle that calls its own restatement of the function"""
    d = examples.DRIVERS["transcr6d"]
    h = k.KrylovFspHandle(examples.driver_model("transcr6d"), max_states=2000000, seed=12345, spmv_variant=2)
    assert h.model_info()["factored"] == 1
    out = h.solve(60.0, [d["x0"]], [1.0], d["fsp_tol"], d["exp_tol"])
    ref = oracle.solve(oracle_model("transcr6d"), [d["x0"]], [1.0], 60.0, d["fsp_tol"], d["exp_tol"], seed=12345,
                       max_size=2000000, reproducible=1)
    assert out["iflag"] == 0 and ref["iflag"] == 0
    assert np.array_equal(out["states"], ref["states"]) and np.array_equal(out["vector"], ref["vector"])
    assert h.phase_seconds()["host_propensity_evals"] == 0
    h.close()


def test_unrecognised_callback_keeps_the_explicit_path():
    """a callback the probe cannot reproduce (Hill-type coupling of two species) stays a host function: explicit matrix only, the
    lattice and index-only variants refuse it loudly"""
    def make():
        m = k.CME_MODEL().create(2, 4, 2)
        m.stoichiometry = np.array([[1, -1, 0, 0], [0, 0, 1, -1]], dtype=np.int32)
        m.reset_parameters([30.0, 1.0])
        m.set_customprop(lambda st, r, p: (p[0] / (1.0 + 0.1 * st[1]) * (1.0 + 0.01 * st[0]), p[1] * st[0],
                                           p[0] / (1.0 + 0.1 * st[0]), p[1] * st[1])[r - 1])
        return m
    assert make().custom_structure(max_molecules=2000)[1] == 0
    for variant in (1, 2):
        with pytest.raises(k.KfspError):
            k.KrylovFspHandle(make(), spmv_variant=variant, max_states=1000, max_molecules=2000)
    h = k.KrylovFspHandle(make(), max_states=20000, max_molecules=2000, seed=5)
    assert h.model_info()["n_host_evaluated"] == 4
    out = h.solve(0.5, [[0, 0]], [1.0], 1e-4, 1e-8)
    assert out["iflag"] == 0 and abs(out["vector"].sum() - 1.0) < 1e-3 and h.phase_seconds()["host_propensity_evals"] > 0
    h.close()
