"""ORACLE package -- TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/kfsp_oracle.cpp (the CPU restatement of the reference's
DGEXPV_FSP hot path) plus the pure-Python restatements of the reference's parser and
`.input` reader.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this package; the product (krylovfspssa_b200/) never does.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

from . import fparser, model_input  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CUSTOM_NONE, CUSTOM_GOUTSIAS, CUSTOM_REPRESSILATOR, CUSTOM_TOGGLE, CUSTOM_PARSER_TEST = 0, 1, 2, 3, 4

PROP_CALLBACK = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double), C.c_void_p)


class Stats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("nmult", "nexph", "nscale", "nstep", "nreject", "ibrkflag", "mbrkdwn", "iflag")] + \
               [(n, C.c_double) for n in
                ("step_min", "step_max", "x_error", "s_error", "tbrkdwn", "t_now", "hump", "beta_ratio")] + \
               [("n_expand", C.c_int64), ("n_drop", C.c_int64), ("wall_seconds", C.c_double), ("setup_seconds", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libkfsp_oracle.so")
    src = os.path.join(_HERE, "kfsp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libkfsp_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp, i32p, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    sig = {
        "ko_model_create": (vp, [C.c_int, C.c_int, C.c_int, i32p, dp]),
        "ko_model_free": (None, [vp]),
        "ko_model_set_params": (None, [vp, dp]),
        "ko_model_set_program": (None, [vp, C.c_int, i32p, C.c_int, dp, C.c_int]),
        "ko_model_set_custom": (None, [vp, C.c_int]),
        "ko_model_set_callback": (None, [vp, PROP_CALLBACK, vp]),
        "ko_model_propensity": (C.c_double, [vp, i32p, C.c_int]),
        "ko_fsp_create": (vp, [vp, C.c_long, C.c_int]),
        "ko_fsp_free": (None, [vp]),
        "ko_fsp_set_states": (None, [vp, i32p, C.c_long]),
        "ko_fsp_matrix_starter": (C.c_int, [vp]),
        "ko_fsp_onestep": (C.c_int, [vp]),
        "ko_fsp_size": (C.c_long, [vp]),
        "ko_fsp_get": (None, [vp, i32p, i32p, dp, dp, dp]),
        "ko_fsp_set_vector": (None, [vp, dp, C.c_long]),
        "ko_fsp_index": (C.c_int, [vp, i32p]),
        "ko_fsp_matvec": (None, [vp, dp, dp]),
        "ko_rng_create": (vp, [C.c_int, C.c_uint64, C.c_char_p]),
        "ko_rng_free": (None, [vp]),
        "ko_rng_draw2": (None, [vp, C.c_uint32, C.c_uint32, C.c_uint32, dp, dp]),
        "ko_fsp_ssa": (None, [vp, C.c_double, vp]),
        "ko_fsp_drop": (C.c_int, [vp, dp, C.c_double, dp, C.POINTER(C.c_long)]),
        "ko_dgpadm": (C.c_int, [C.c_int, C.c_int, C.c_double, dp, C.c_int, dp, i32p, dp]),
        "ko_solver_create": (vp, []),
        "ko_solver_free": (None, [vp]),
        "ko_solver_set_options": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "ko_solver_set_reproducible": (None, [vp, C.c_int]),
        "ko_fsp_set_reproducible": (None, [vp, C.c_int]),
        "ko_dgpadm_reproducible": (C.c_int, [C.c_int, C.c_double, dp, C.c_int, dp, i32p, dp]),
        "ko_dot_reproducible": (C.c_double, [C.c_long, dp, dp]),
        "ko_solve": (C.c_int, [vp, vp, C.c_double, dp, C.c_long, C.c_double, C.c_double, C.c_int, vp]),
        "ko_trace_len": (C.c_long, [vp]),
        "ko_trace_get": (None, [vp, dp, i32p]),
        "ko_stats_get": (None, [vp, C.POINTER(Stats)]),
        "ko_arnoldi_sweep": (C.c_double, [vp, dp, C.c_int, dp, dp, i32p]),
        "ko_time_matvec": (C.c_double, [vp, dp, dp, C.c_int]),
        "ko_last_sweep": (C.c_int, [dp, C.c_int, dp]),
        "ko_combine_reproducible": (C.c_double, [C.c_long, C.c_int, C.c_double, dp, C.c_long, dp, dp, dp, dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def find_libgfortran():
    try:
        import scipy
    except ImportError:
        return None
    d = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs")
    c = sorted(glob.glob(os.path.join(d, "libgfortran*.so*")))
    return c[0] if c else None


class Model:
    """Oracle-side CME_MODEL (ModelModule.f90:14-42)."""

    def __init__(self, nspecies, nreactions, nparameters, stoichiometry, params=None):
        self.S, self.R, self.P = nspecies, nreactions, nparameters
        st = np.ascontiguousarray(np.asarray(stoichiometry, dtype=np.int32).reshape(nreactions, nspecies))
        self.stoich = st                      # [reaction, species] == Fortran (S,R) column-major
        p = np.zeros(max(nparameters, 1)) if params is None else np.asarray(params, dtype=np.float64)
        self.params = np.ascontiguousarray(p, dtype=np.float64)
        self.h = lib().ko_model_create(nspecies, nreactions, nparameters, _i32(st), _f64(self.params))
        self.programs = None
        self._cb = None

    @classmethod
    def load(cls, path, params=None):
        im = model_input.load(path)
        m = cls(im.nspecies, im.nreactions, im.nparameters, im.stoichiometry, params)
        m.species_names = im.species_names
        m.parameter_names = im.parameter_names
        m.set_programs(im.programs)
        return m

    def set_programs(self, programs):
        self.programs = programs
        for r, pr in enumerate(programs):
            code = np.asarray(pr.code, dtype=np.int32)
            imm = np.asarray(pr.immed if pr.immed else [0.0], dtype=np.float64)
            lib().ko_model_set_program(self.h, r, _i32(code), len(pr.code), _f64(imm), len(pr.immed))

    def set_custom(self, kind):
        lib().ko_model_set_custom(self.h, kind)

    def set_callback(self, fn):
        def tramp(state_p, reaction, params_p, _ctx):
            st = np.ctypeslib.as_array(state_p, shape=(self.S,))
            pr = np.ctypeslib.as_array(params_p, shape=(max(self.P, 1),))
            return float(fn(st, reaction, pr))
        self._cb = PROP_CALLBACK(tramp)
        lib().ko_model_set_callback(self.h, self._cb, None)

    def reset_parameters(self, p):
        self.params = np.ascontiguousarray(p, dtype=np.float64)
        lib().ko_model_set_params(self.h, _f64(self.params))

    def propensity(self, state, reaction):
        st = np.ascontiguousarray(state, dtype=np.int32)
        return lib().ko_model_propensity(self.h, _i32(st), reaction)

    def __del__(self):
        try:
            lib().ko_model_free(self.h)
        except Exception:
            pass


class Fsp:
    """Oracle-side FINITE_STATE_PROJECTION (StateSpace.f90:19-45)."""
    NMAX = 6291469

    def __init__(self, model, max_size=NMAX, maxmol=10000, reproducible=0):
        self.model = model
        self.h = lib().ko_fsp_create(model.h, max_size, maxmol)
        if reproducible:
            lib().ko_fsp_set_reproducible(self.h, 1)

    def set_states(self, states):
        st = np.ascontiguousarray(np.asarray(states, dtype=np.int32).reshape(-1, self.model.S))
        lib().ko_fsp_set_states(self.h, _i32(st), st.shape[0])

    def matrix_starter(self):
        return lib().ko_fsp_matrix_starter(self.h)

    def onestep(self):
        return lib().ko_fsp_onestep(self.h)

    def ssa(self, timestep, rng):
        lib().ko_fsp_ssa(self.h, timestep, rng.h)

    def drop(self, w, dsum):
        w = np.ascontiguousarray(w, dtype=np.float64)
        tol = C.c_double(0)
        cnt = C.c_long(0)
        did = lib().ko_fsp_drop(self.h, _f64(w), dsum, C.byref(tol), C.byref(cnt))
        return did, w[: self.size].copy(), tol.value, cnt.value

    @property
    def size(self):
        return lib().ko_fsp_size(self.h)

    def get(self):
        n, S, R = self.size, self.model.S, self.model.R
        states = np.zeros((n, S), dtype=np.int32)
        adj = np.zeros((n, R), dtype=np.int32)
        off = np.zeros((n, R))
        diag = np.zeros(n)
        vec = np.zeros(n)
        lib().ko_fsp_get(self.h, _i32(states), _i32(adj), _f64(off), _f64(diag), _f64(vec))
        return dict(states=states, adj=adj, offdiag=off, diag=diag, vector=vec)

    def index(self, state):
        st = np.ascontiguousarray(state, dtype=np.int32)
        return lib().ko_fsp_index(self.h, _i32(st))

    def matvec(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.size)
        lib().ko_fsp_matvec(self.h, _f64(x), _f64(y))
        return y

    def __del__(self):
        try:
            lib().ko_fsp_free(self.h)
        except Exception:
            pass


class Rng:
    GFORTRAN, PHILOX = 0, 1

    def __init__(self, mode=1, seed=12345):
        path = None
        if mode == 0:
            path = find_libgfortran()
            if path is None:
                raise RuntimeError("libgfortran not found")
        self.h = lib().ko_rng_create(mode, seed, path.encode() if path else None)
        if not self.h:
            raise RuntimeError("cannot create rng (mode %d)" % mode)

    def draw2(self, j0, jump, call_no):
        a, b = C.c_double(), C.c_double()
        lib().ko_rng_draw2(self.h, j0, jump, call_no, C.byref(a), C.byref(b))
        return a.value, b.value

    def __del__(self):
        try:
            lib().ko_rng_free(self.h)
        except Exception:
            pass


def dot_reproducible(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().ko_dot_reproducible(len(x), _f64(x), _f64(y))


def arnoldi_sweep(fsp, v, m):
    """One IOP-2 sweep (KrylovSolver.f90:236-263) from v on the oracle's matrix.  Canonical mode (the Fsp was made with
    reproducible=1): returns H ((m+2)^2, Fortran order, unit entry set), the UN-NORMALISED basis (n x (m+2), Fortran
    order), its column scales, AVNORM and the happy-breakdown column."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    n = len(v)
    work = np.zeros((n, m + 2), order="F")
    H = np.zeros((m + 2, m + 2), order="F")
    nm = C.c_int32()
    lib().ko_arnoldi_sweep(fsp.h, _f64(v), m, _f64(work), _f64(H), C.byref(nm))
    cs = np.ones(m + 2)
    av = C.c_double(0)
    brk = lib().ko_last_sweep(_f64(cs), m + 2, C.byref(av))
    H[m + 1, m] = 1.0
    return dict(H=H, basis=work, colscale=cs, avnorm=av.value, brk=brk, nmult=nm.value)


def combine_reproducible(V, e, beta, colscale=None):
    """W = max(beta * sum_j e_j (cs_j V_j), 0) in canonical arithmetic; returns (w, ||w||_1, sum w^2)."""
    V = np.asfortranarray(V, dtype=np.float64)
    n, mx = V.shape
    e = np.ascontiguousarray(e, dtype=np.float64)
    cs = np.ones(mx) if colscale is None else np.ascontiguousarray(colscale, dtype=np.float64)
    w = np.zeros(n)
    ssq = C.c_double(0)
    ws = lib().ko_combine_reproducible(n, mx, beta, _f64(V), n, _f64(e), _f64(cs), _f64(w), C.byref(ssq))
    return w, ws, ssq.value


def dgpadm(H, t, ideg=6, m=None, reproducible=0):
    """exp(t*H[:m,:m]) by the reference's Pade routine; returns (E, ns, hnorm).
    reproducible=1 selects the canonical operation order the device kernel is held to."""
    H = np.asfortranarray(H, dtype=np.float64)
    ldh = H.shape[0]
    m = ldh if m is None else m
    out = np.zeros((m, m), order="F")
    ns = C.c_int32(0)
    hn = C.c_double(0)
    if reproducible:
        rc = lib().ko_dgpadm_reproducible(m, t, _f64(H), ldh, _f64(out), C.byref(ns), C.byref(hn))
    else:
        rc = lib().ko_dgpadm(ideg, m, t, _f64(H), ldh, _f64(out), C.byref(ns), C.byref(hn))
    if rc:
        raise RuntimeError("dgpadm iflag=%d" % rc)
    return out, ns.value, hn.value


def solve(model, states0, p0, t, fsptol, krytol, seed=12345, rng_mode=1, max_size=Fsp.NMAX,
          m_max=100, m_min=10, n_init_onestep=5, enable_drop=1, enable_expand=1, itrace=0, reproducible=0):
    """CME_SOLVE / DGEXPV_FSP (KrylovSolver.f90:7-36, 40-573) on the oracle.
    reproducible=1: canonical arithmetic (double-double reductions, fixed-order fma chains)."""
    L = lib()
    fsp = Fsp(model, max_size)
    fsp.set_states(states0)
    s = L.ko_solver_create()
    L.ko_solver_set_options(s, m_max, m_min, n_init_onestep, enable_drop, enable_expand)
    L.ko_solver_set_reproducible(s, 1 if reproducible else 0)
    rng = Rng(rng_mode, seed)
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    rc = L.ko_solve(s, fsp.h, t, _f64(p0), len(p0), fsptol, krytol, itrace, rng.h)
    n = L.ko_trace_len(s)
    td = np.zeros((max(n, 1), 6))
    ti = np.zeros((max(n, 1), 6), dtype=np.int32)
    L.ko_trace_get(s, _f64(td), _i32(ti))
    st = Stats()
    L.ko_stats_get(s, C.byref(st))
    L.ko_solver_free(s)
    out = fsp.get()
    out.update(iflag=rc, trace_d=td[:n], trace_i=ti[:n],
               stats={k: getattr(st, k) for k, _ in Stats._fields_}, fsp=fsp)
    return out
