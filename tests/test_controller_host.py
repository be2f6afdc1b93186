"""Host logic without a GPU: the product's DGEXPV_FSP controller template
(krylovfspssa_b200/csrc/controller.h) driven by oracle numerics must take exactly the
decisions of the oracle's own restatement of KrylovSolver.f90:127-573 -- same (T_STEP, M, N)
per step, same expansion / drop / breakdown events, same counters."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from krylovfspssa_b200._lib import Stats, TraceRow

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def cc():
    so = os.path.join(HERE, "host", "libcontroller_check.so")
    srcs = [os.path.join(HERE, "host", "controller_check.cpp"), os.path.join(ROOT, "oracle", "kfsp_oracle.cpp")]
    deps = srcs + [os.path.join(ROOT, "krylovfspssa_b200", "csrc", "controller.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off",
                               "-o", so] + srcs + ["-ldl"])
    L = C.CDLL(so)
    L.cc_run.restype = C.c_int
    L.cc_run.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_long, C.c_double, C.c_double, C.c_double, C.c_uint64,
                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(TraceRow), C.c_long,
                         C.POINTER(C.c_long), C.POINTER(Stats), C.POINTER(C.c_double), C.c_long]
    # the check library carries its own copy of the oracle: bind the ko_* entry points we need from it
    for name in ("ko_model_create", "ko_model_set_program", "ko_fsp_create", "ko_fsp_set_states", "ko_fsp_get", "ko_fsp_size"):
        src = getattr(oracle.lib(), name)
        fn = getattr(L, name)
        fn.restype, fn.argtypes = src.restype, src.argtypes
    return L


def run_controller(L, model_path, params, x0, t, ftol, ktol, seed=12345, m_max=100, m_min=10, n_init=5, drop=1, expand=1):
    im = oracle.model_input.load(model_path)
    st = np.ascontiguousarray(np.array(im.stoichiometry, dtype=np.int32))
    par = np.ascontiguousarray(params, dtype=np.float64)
    mh = L.ko_model_create(im.nspecies, im.nreactions, im.nparameters, st.ctypes.data_as(C.POINTER(C.c_int32)),
                           par.ctypes.data_as(C.POINTER(C.c_double)))
    for r, pr in enumerate(im.programs):
        code = np.asarray(pr.code, dtype=np.int32)
        imm = np.asarray(pr.immed if pr.immed else [0.0])
        L.ko_model_set_program(mh, r, code.ctypes.data_as(C.POINTER(C.c_int32)), len(pr.code),
                               imm.ctypes.data_as(C.POINTER(C.c_double)), len(pr.immed))
    fsp = L.ko_fsp_create(mh, 6291469, 10000)
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.int32).reshape(-1, im.nspecies))
    L.ko_fsp_set_states(fsp, x0.ctypes.data_as(C.POINTER(C.c_int32)), x0.shape[0])
    p0 = np.zeros(x0.shape[0]); p0[0] = 1.0
    cap = 100000
    rows = (TraceRow * cap)()
    nrows = C.c_long()
    stats = Stats()
    w = np.zeros(2000000)
    rc = L.cc_run(fsp, p0.ctypes.data_as(C.POINTER(C.c_double)), len(p0), t, ftol, ktol, seed, im.nreactions,
                  m_max, m_min, n_init, drop, expand, rows, cap, C.byref(nrows), C.byref(stats),
                  w.ctypes.data_as(C.POINTER(C.c_double)), len(w))
    n = L.ko_fsp_size(fsp)
    states = np.zeros((n, im.nspecies), dtype=np.int32)
    L.ko_fsp_get(fsp, states.ctypes.data_as(C.POINTER(C.c_int32)), None, None, None, None)
    ti = np.array([[r.m, r.n_step, r.n_after, r.flags, r.nmult, r.nexph] for r in rows[:nrows.value]])
    td = np.array([[r.t_now, r.t_step, r.t_new, r.wsum, r.err_loc, r.beta] for r in rows[:nrows.value]])
    return rc, ti, td, states, w[:n].copy(), stats


CASES = [
    ("toggle.input", [1, 100, 1, 1, 100, 1], [0, 0], 20.0, 1e-4, 1e-10),
    ("birth_death.input", [20.0, 1.0], [0], 2.0, 1e-6, 1e-10),
    ("goutsias.input", [0.043, 0.0007, 0.0715, 0.0039, 0.0199264663575241, 0.4791, 0.000199264663575241,
                        0.8765e-11, 0.0830269431563506104, 0.5], [2, 6, 0, 2, 0, 0], 10.0, 1e-6, 1e-8),
    ("repressilator.input", [100, 100, 100, 1, 1, 1], [22, 0, 0], 0.5, 1e-4, 1e-10),
]


@pytest.mark.parametrize("name,params,x0,t,ftol,ktol", CASES)
def test_controller_matches_oracle_trace(cc, models_dir, name, params, x0, t, ftol, ktol):
    path = os.path.join(models_dir, name)
    rc, ti, td, states, w, stats = run_controller(cc, path, params, x0, t, ftol, ktol)
    m = oracle.Model.load(path, params)
    ref = oracle.solve(m, [x0], [1.0], t, ftol, ktol)
    assert rc == ref["iflag"] == 0
    assert ti.shape == ref["trace_i"].shape
    assert np.array_equal(ti, ref["trace_i"])                       # M, N, N_after, flags, NMULT, NEXPH per step
    assert np.array_equal(td[:, 1], ref["trace_d"][:, 1])           # T_STEP is rounded to 2 digits: exact
    assert np.array_equal(td[:, 2], ref["trace_d"][:, 2])           # T_NEW
    assert np.allclose(td[:, 0], ref["trace_d"][:, 0], rtol=0, atol=1e-12)
    assert np.array_equal(states, ref["states"])                    # state set and order identical
    assert np.abs(w - ref["vector"]).sum() <= 1e-10 * np.abs(ref["vector"]).sum()
    s = ref["stats"]
    for k in ("nmult", "nexph", "nscale", "nstep", "nreject", "ibrkflag", "mbrkdwn"):
        assert getattr(stats, k) == s[k], k
    assert stats.n_expand == s["n_expand"] and stats.n_drop == s["n_drop"]


def test_controller_fixed_state_set(cc, models_dir):
    # adaptivity off: no start-up expansion, no SSA/one-step, no drop (the synthetic benchmark mode)
    path = os.path.join(models_dir, "birth_death.input")
    x0 = [[i] for i in range(60)]
    rc, ti, td, states, w, stats = run_controller(cc, path, [20.0, 1.0], x0, 1.0, 1e-6, 1e-10, n_init=0, drop=0, expand=0)
    assert rc == 0 and states.shape[0] == 60 and stats.n_expand == 0 and stats.n_drop == 0
    m = oracle.Model.load(path, [20.0, 1.0])
    ref = oracle.solve(m, x0, np.eye(60)[0], 1.0, 1e-6, 1e-10, n_init_onestep=0, enable_drop=0, enable_expand=0)
    assert np.array_equal(ti, ref["trace_i"])
    assert np.abs(w - ref["vector"]).sum() <= 1e-12
