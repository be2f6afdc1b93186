"""Host-side mirror of the reference interface for the hot path (names, argument meaning
and error behaviour follow the Fortran; STOP becomes KfspError)."""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Options, Stats, TraceRow, check, lib

NMAX = 6291469                    # src/state_space/StateSpace.f90:10
MAXNUMBERMOLECULES = 10000        # src/state_space/StateSpace.f90:11


def models_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_options(**kw):
    o = Options()
    check(lib().kfsp_default_options(C.byref(o)))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError("unknown option %r" % k)
        setattr(o, k, v)
    return o


class CME_MODEL:
    """TYPE CME_MODEL (ModelModule.f90:14-42): CREATE, LOAD, RESET_PARAMETERS, PROPENSITY, CUSTOMPROP."""

    def __init__(self):
        self._h = C.c_void_p()
        self.loaded = False
        self._cb = None
        self.customprop = None

    # CREATE(THIS, N_SPECIES, N_REACTIONS, N_PARAMETERS)  ModelModule.f90:46-57
    def create(self, n_species, n_reactions, n_parameters):
        self._free()
        check(lib().kfsp_model_create(n_species, n_reactions, n_parameters, C.byref(self._h)), "CME_MODEL%CREATE")
        return self

    # LOAD(THIS, FILENAME)  ModelModule.f90:59-161
    def load(self, filename="model.input"):
        self._free()
        check(lib().kfsp_model_load(os.fsencode(filename), C.byref(self._h)), "CME_MODEL%LOAD")
        self.loaded = True
        return self

    def _dims(self):
        s, r, p = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().kfsp_model_dims(self._h, C.byref(s), C.byref(r), C.byref(p)))
        return s.value, r.value, p.value

    nspecies = property(lambda self: self._dims()[0])
    nreactions = property(lambda self: self._dims()[1])
    nparameters = property(lambda self: self._dims()[2])

    @property
    def stoichiometry(self):
        """STOICHIOMETRY(NSPECIES, NREACTIONS): returned as an (S, R) array, one column per reaction."""
        s, r, _ = self._dims()
        buf = np.zeros((r, s), dtype=np.int32)
        check(lib().kfsp_model_get_stoichiometry(self._h, _i32(buf)))
        return buf.T.copy()

    @stoichiometry.setter
    def stoichiometry(self, value):
        s, r, _ = self._dims()
        v = np.asarray(value, dtype=np.int32)
        if v.shape != (s, r):
            raise ValueError("stoichiometry must have shape (nspecies, nreactions)")
        buf = np.ascontiguousarray(v.T)
        check(lib().kfsp_model_set_stoichiometry(self._h, _i32(buf)))

    def _names(self, fn, n):
        out = []
        for i in range(n):
            b = C.create_string_buffer(64)
            check(fn(self._h, i, b, 64))
            out.append(b.value.decode())
        return out

    species_names = property(lambda self: self._names(lib().kfsp_model_species_name, self._dims()[0]))
    parameter_names = property(lambda self: self._names(lib().kfsp_model_parameter_name, self._dims()[2]))

    # RESET_PARAMETERS(THIS, PVAL)  ModelModule.f90:201-217
    def reset_parameters(self, pval):
        p = np.ascontiguousarray(pval, dtype=np.float64)
        check(lib().kfsp_model_reset_parameters(self._h, _f64(p), len(p)), "CME_MODEL%RESET_PARAMETERS")
        self.parameter_val = p.copy()
        return self

    def clone(self):
        """An independent copy (sizes, stoichiometry, byte code or CUSTOMPROP, current parameter values): concurrent solves of
        a parameter sweep each RESET_PARAMETERS their own copy."""
        s, r, p = self._dims()
        m = CME_MODEL().create(s, r, p)
        m.stoichiometry = self.stoichiometry
        if self.customprop is not None:
            m.set_customprop(self.customprop)
        else:
            for k in range(1, r + 1):
                code, imm = self.bytecode(k)
                c = np.asarray(code, dtype=np.int32)
                v = np.asarray(imm if imm else [0.0], dtype=np.float64)
                check(lib().kfsp_model_set_propensity_bytecode(m._h, k, _i32(c), len(code), _f64(v), len(imm)))
        if getattr(self, "parameter_val", None) is not None:
            m.reset_parameters(self.parameter_val)
        m.loaded = self.loaded
        return m

    def set_propensity(self, reaction, expr):
        """Compile one propensity string as LOAD does (ModelModule.f90:152-155)."""
        check(lib().kfsp_model_set_propensity_string(self._h, reaction, expr.encode()), "EQUATIONPARSER")

    def bytecode(self, reaction):
        code = np.zeros(1024, dtype=np.int32)
        imm = np.zeros(256)
        nc, ni = C.c_int32(1024), C.c_int32(256)
        check(lib().kfsp_model_get_propensity_bytecode(self._h, reaction, _i32(code), C.byref(nc), _f64(imm), C.byref(ni)))
        return code[:nc.value].tolist(), imm[:ni.value].tolist()

    def set_customprop(self, fn):
        """MODEL%CUSTOMPROP => fn(state, reaction, parameters)  (ModelModule.f90:6-12, 31).
        `fn` is a Python callable, or a compiled host function given as a ctypes function pointer."""
        if isinstance(fn, C._CFuncPtr):
            self._cb = C.cast(fn, _lib.PROPENSITY_FN)
            self.customprop = fn
            check(lib().kfsp_model_set_custom_propensity(self._h, self._cb, None))
            return
        s = self._dims()[0]
        p = max(self._dims()[2], 1)

        def tramp(state_p, reaction, params_p, _ctx):
            return float(fn(np.ctypeslib.as_array(state_p, shape=(s,)), reaction,
                            np.ctypeslib.as_array(params_p, shape=(p,))))
        self._cb = _lib.PROPENSITY_FN(tramp)
        self.customprop = fn
        check(lib().kfsp_model_set_custom_propensity(self._h, self._cb, None))

    def custom_structure(self, max_molecules=10000):
        """Structure of the CUSTOMPROP callback found by probing (kfsp_model_custom_structure; host only): (species, kind) with
        species[k] the 0-based species reaction k+1 reads (0: constant, -2: several) and kind 1 (all single-species: device
        tables), 2 (single-species + bilinear mass action: device tables and byte code) or 0 (host callbacks)."""
        sp = np.zeros(self._dims()[1], dtype=np.int32)
        kind = C.c_int32()
        check(lib().kfsp_model_custom_structure(self._h, int(max_molecules), _i32(sp), C.byref(kind)))
        return sp.tolist(), int(kind.value)

    # PROPENSITY(THIS, STATE, REACTION)  ModelModule.f90:163-199
    def propensity(self, state, reaction):
        st = np.ascontiguousarray(state, dtype=np.int32)
        out = C.c_double()
        check(lib().kfsp_model_propensity(self._h, _i32(st), reaction, C.byref(out)), "CME_MODEL%PROPENSITY")
        return out.value

    def propensity_factored(self, state, reaction):
        """(value, nterms, nops): the propensity evaluated through the factored form of the index-only SpMV (spmv_variant = 2)."""
        st = np.ascontiguousarray(state, dtype=np.int32)
        out, nt, no = C.c_double(), C.c_int32(), C.c_int32()
        check(lib().kfsp_model_propensity_factored(self._h, _i32(st), reaction, C.byref(out), C.byref(nt), C.byref(no)), "factor_program")
        return out.value, nt.value, no.value

    def _free(self):
        if self._h:
            lib().kfsp_model_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass


class KrylovFspHandle:
    """Owner of one device solver (kfsp_handle): streams, device state space, workspaces."""

    def __init__(self, model, options=None, **opt_kw):
        self.model = model
        self.options = options if options is not None else default_options(**opt_kw)
        self._h = C.c_void_p()
        check(lib().kfsp_create(C.byref(self.options), C.byref(self._h)), "kfsp_create")
        check(lib().kfsp_set_model(self._h, model._h), "kfsp_set_model")
        self.S, self.R, _ = model._dims()

    def set_model(self, model=None):
        """Ship MODEL to the device again, e.g. after RESET_PARAMETERS (ModelModule.f90:201-217): the device state space
        and workspaces of the handle are kept when the model's shape is unchanged."""
        if model is not None:
            self.model = model
        check(lib().kfsp_set_model(self._h, self.model._h), "kfsp_set_model")
        self.S, self.R, _ = self.model._dims()

    def close(self):
        if self._h:
            lib().kfsp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU (one process per GPU, rows block-partitioned) ---------------------
    @staticmethod
    def dist_unique_id():
        buf = (C.c_uint8 * _lib.NCCL_ID_BYTES)()
        check(lib().kfsp_dist_unique_id(buf), "kfsp_dist_unique_id")
        return bytes(buf)

    def dist_init(self, rank, nranks, unique_id):
        buf = (C.c_uint8 * _lib.NCCL_ID_BYTES).from_buffer_copy(unique_id)
        check(lib().kfsp_dist_init(self._h, rank, nranks, buf), "kfsp_dist_init")

    def dist_info(self):
        v = [C.c_int64() for _ in range(6)]
        check(lib().kfsp_dist_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("lo", "hi", "n_halo", "n_send", "halo_bytes", "reductions"), [x.value for x in v]))

    def model_info(self):
        """how the propensities of the current model are evaluated: tabulated / by the host / by the CUDA math library / factored"""
        v = [C.c_int32() for _ in range(4)]
        check(lib().kfsp_model_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("n_tabulated", "n_host_evaluated", "n_device_libm", "factored"), [x.value for x in v]))

    def dist_exchange_stats(self, reset=False):
        """fused reduction exchanges of the peer-memory path: count, mean and max microseconds from posting to holding all partials"""
        n, mean, mx = C.c_int64(), C.c_double(), C.c_double()
        check(lib().kfsp_dist_exchange_stats(self._h, C.byref(n), C.byref(mean), C.byref(mx), 1 if reset else 0))
        return {"exchanges": n.value, "mean_us": mean.value, "max_us": mx.value}

    # ---- state space ------------------------------------------------------------------
    def fsp_init(self, states):
        st = np.ascontiguousarray(np.asarray(states, dtype=np.int32).reshape(-1, self.S))
        check(lib().kfsp_fsp_init(self._h, st.shape[0], _i32(st)), "MATRIX_STARTER")

    def fsp_init_box(self, bounds):
        """spmv_variant = 1: the projection is the lattice [0,bounds[0]) x ... in natural order (no state list)."""
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        check(lib().kfsp_fsp_init_box(self._h, _i32(b)), "kfsp_fsp_init_box")

    def onestep(self):
        check(lib().kfsp_fsp_onestep(self._h), "ONESTEP_EXTENDER")

    def ssa(self, timestep):
        check(lib().kfsp_fsp_ssa(self._h, timestep), "SSA_EXTENDER")

    def drop(self, dsum):
        d, tol, cnt = C.c_int32(), C.c_double(), C.c_int64()
        check(lib().kfsp_fsp_drop(self._h, dsum, C.byref(d), C.byref(tol), C.byref(cnt)), "DROP_STATES")
        return d.value, tol.value, cnt.value

    @property
    def size(self):
        n = C.c_int64()
        check(lib().kfsp_fsp_size(self._h, C.byref(n)))
        return n.value

    def set_vector(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        check(lib().kfsp_fsp_set_vector(self._h, _f64(v), len(v)))

    def get(self, matrix=True):
        n = self.size
        states = np.zeros((n, self.S), dtype=np.int32)
        vec = np.zeros(n)
        if matrix:
            adj = np.zeros((n, self.R), dtype=np.int32)
            off = np.zeros((n, self.R))
            diag = np.zeros(n)
            check(lib().kfsp_fsp_get(self._h, _i32(states), _i32(adj), _f64(off), _f64(diag), _f64(vec)))
            return dict(states=states, adj=adj, offdiag=off, diag=diag, vector=vec)
        check(lib().kfsp_fsp_get(self._h, _i32(states), None, None, None, _f64(vec)))
        return dict(states=states, vector=vec)

    def index(self, states):
        st = np.ascontiguousarray(np.asarray(states, dtype=np.int32).reshape(-1, self.S))
        out = np.zeros(st.shape[0], dtype=np.int32)
        check(lib().kfsp_fsp_index(self._h, st.shape[0], _i32(st), _i32(out)))
        return out

    def probability(self, states):
        st = np.ascontiguousarray(np.asarray(states, dtype=np.int32).reshape(-1, self.S))
        out = np.zeros(st.shape[0])
        check(lib().kfsp_fsp_probability(self._h, st.shape[0], _i32(st), _f64(out)))
        return out

    # ---- kernels ----------------------------------------------------------------------
    def matvec(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.size)
        check(lib().kfsp_matvec(self._h, _f64(x), _f64(y)), "FMATVEC")
        return y

    def arnoldi(self, v, m):
        v = np.ascontiguousarray(v, dtype=np.float64)
        H = np.zeros((m + 2, m + 2), order="F")
        av, brk, sec = C.c_double(), C.c_int32(), C.c_double()
        check(lib().kfsp_arnoldi(self._h, _f64(v), m, _f64(H), C.byref(av), C.byref(brk), C.byref(sec)), "ARNOLDI")
        return H, av.value, brk.value, sec.value

    def expm(self, H, t, m=None):
        H = np.asfortranarray(H, dtype=np.float64)
        ldh = H.shape[0]
        m = ldh if m is None else m
        out = np.zeros((m, m), order="F")
        ns, hn = C.c_int32(), C.c_double()
        check(lib().kfsp_expm(self._h, m, t, _f64(H), ldh, _f64(out), C.byref(ns), C.byref(hn)), "DGPADM")
        return out, ns.value, hn.value

    def combine(self, V, e, beta, colscale=None, with_ssq=False):
        V = np.asfortranarray(V, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.float64)
        n, mx = V.shape
        w = np.zeros(n)
        ws, ssq = C.c_double(), C.c_double()
        cs = None if colscale is None else _f64(np.ascontiguousarray(colscale, dtype=np.float64))
        check(lib().kfsp_combine(self._h, n, mx, beta, _f64(V), _f64(e), cs, _f64(w), C.byref(ws), C.byref(ssq)))
        return (w, ws.value, ssq.value) if with_ssq else (w, ws.value)

    # ---- solve ------------------------------------------------------------------------
    def solve(self, t, states_in, p_in, fsp_tol, kry_tol, verbosity=0, max_out=None):
        st = np.ascontiguousarray(np.asarray(states_in, dtype=np.int32).reshape(-1, self.S))
        p = np.ascontiguousarray(p_in, dtype=np.float64)
        if len(p) < st.shape[0]:
            p = np.concatenate([p, np.zeros(st.shape[0] - len(p))])
        cap = int(max_out if max_out is not None else self.options.max_states)
        states_out = np.zeros((cap, self.S), dtype=np.int32)
        p_out = np.zeros(cap)
        n_out = C.c_int64()
        stats = Stats()
        rc = lib().kfsp_solve(self._h, t, st.shape[0], _i32(st), _f64(p), fsp_tol, kry_tol, verbosity,
                              C.byref(n_out), _i32(states_out), _f64(p_out), cap, C.byref(stats))
        if rc < 0:
            raise _lib.KfspError(rc, "DGEXPV_FSP")
        n = n_out.value
        return dict(iflag=rc, states=states_out[:n].copy(), vector=p_out[:n].copy(), stats=stats.as_dict(),
                    trace=self.trace())

    def solve_resident(self, t, fsp_tol, kry_tol, verbosity=0):
        stats = Stats()
        rc = lib().kfsp_solve_resident(self._h, t, fsp_tol, kry_tol, verbosity, C.byref(stats))
        if rc < 0:
            raise _lib.KfspError(rc, "DGEXPV_FSP")
        return rc, stats.as_dict()

    def trace(self):
        n = C.c_int64()
        check(lib().kfsp_trace_length(self._h, C.byref(n)))
        rows = (TraceRow * max(n.value, 1))()
        check(lib().kfsp_trace_get(self._h, rows, n.value))
        d = np.array([[r.t_now, r.t_step, r.t_new, r.wsum, r.err_loc, r.beta] for r in rows[:n.value]]).reshape(-1, 6)
        i = np.array([[r.m, r.n_step, r.n_after, r.flags, r.nmult, r.nexph] for r in rows[:n.value]], dtype=np.int32).reshape(-1, 6)
        return dict(d=d, i=i)

    PROF_CLASSES = ("spmv_plain", "spmv_dot", "spmv_nrm", "spmv_fin_dot", "spmv_fin_nrm", "axpy_dot", "axpy_nrm", "combine",
                    "scale_copy", "expm", "sweep")

    def profile(self):
        """Device seconds, launches and algorithmic bytes per state (summed over the launches) of the last solve by
        kernel class (set_profiling(True) before the solve)."""
        sec = (C.c_double * 12)()
        cnt = (C.c_int64 * 12)()
        bps = (C.c_int64 * 12)()
        check(lib().kfsp_profile_get(self._h, sec, cnt, bps))
        return {name: (sec[i], cnt[i], bps[i]) for i, name in enumerate(self.PROF_CLASSES)}

    def set_profiling(self, level=2):
        """0 off, 1 one event pair per Arnoldi sweep, 2 one pair per launch"""
        check(lib().kfsp_set_profiling(self._h, int(level)))

    def set_blocking_sync(self, on=True):
        """host waits of this handle sleep on a blocking event instead of spinning (many concurrent handles on few cores)"""
        check(lib().kfsp_set_blocking_sync(self._h, 1 if on else 0))

    def phase_seconds(self):
        buf = (C.c_double * 8)()
        check(lib().kfsp_phase_seconds(self._h, buf))
        return dict(zip(("sweep_pade", "combine_norms", "ssa", "drop", "onestep", "host_callbacks", "ssa_cache_rounds",
                         "host_propensity_evals"), list(buf)))

    @property
    def launches(self):
        n = C.c_int64()
        check(lib().kfsp_launch_count(self._h, C.byref(n)))
        return n.value


class FINITE_STATE_PROJECTION:
    """TYPE FINITE_STATE_PROJECTION (StateSpace.f90:19-45) as the host sees it: SIZE, STATE, VECTOR,
    CREATE, CLEAR, PROBABILITY, INDEX.  After CME_SOLVE, PROBABILITY/INDEX are served from the device
    table of the solve that filled this object."""

    def __init__(self):
        self.max_size = NMAX
        self.size = 0
        self.state = None            # (NSPECIES, SIZE) like the Fortran STATE(:, 1:SIZE)
        self.vector = None
        self._model = None
        self._handle = None

    # CREATE(FSP, MODEL, MAX_SIZE_CUSTOM)  StateSpace.f90:51-83
    def create(self, model, max_size_custom=None):
        self._model = model
        if max_size_custom is not None:
            self.max_size = int(max_size_custom)
        self.size = 0
        self.state = np.zeros((model.nspecies, 0), dtype=np.int32)
        self.vector = np.zeros(0)
        return self

    # CLEAR(FSP)  StateSpace.f90:85-93
    def clear(self):
        self.state = None
        self.vector = None
        self.size = 0
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def set(self, states, vector=None):
        st = np.asarray(states, dtype=np.int32).reshape(-1, self._model.nspecies)
        self.state = st.T.copy()
        self.size = st.shape[0]
        self.vector = np.zeros(self.size) if vector is None else np.asarray(vector, dtype=np.float64).copy()

    # PROBABILITY(FSP, X)  StateSpace.f90:96-114
    def probability(self, x):
        if self._handle is None:
            raise RuntimeError("FSP holds no solved state space")
        return float(self._handle.probability([x])[0])

    # INDEX(FSP, X)  StateSpace.f90:116-134
    def index(self, x):
        if self._handle is None:
            raise RuntimeError("FSP holds no solved state space")
        return int(self._handle.index([x])[0])


def CME_SOLVE(model, t, fsp_in, fsp_out, fsptol, exp_tol, verbosity=0, options=None, **opt_kw):
    """CME_SOLVE(MODEL, T, FSP_IN, FSP_OUT, FSPTOL, EXP_TOL, VERBOSITY)  KrylovSolver.f90:7-36.

    FSP_IN supplies the initial probability vector, FSP_OUT the initial states; on return FSP_OUT
    holds the final states and vector.  Returns the solver statistics the reference discards."""
    if options is None:
        opt_kw.setdefault("max_states", fsp_out.max_size)
        options = default_options(**opt_kw)
    h = KrylovFspHandle(model, options)
    n_in = fsp_out.size
    states = fsp_out.state.T[:n_in]
    p = np.zeros(n_in)
    k = min(n_in, len(fsp_in.vector))
    p[:k] = fsp_in.vector[:k]
    out = h.solve(t, states, p, fsptol, exp_tol, verbosity, max_out=options.max_states)
    fsp_out.state = out["states"].T.copy()
    fsp_out.vector = out["vector"]
    fsp_out.size = out["states"].shape[0]
    if fsp_out._handle is not None:
        fsp_out._handle.close()
    fsp_out._handle = h
    out["handle"] = h
    return out


cme_solve = CME_SOLVE
