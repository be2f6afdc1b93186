"""ctypes binding of libkfsp.so (include/kfsp.h).  There is no fallback: if the CUDA
library is missing the import of any solver entry point raises."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KFSP_LIB", os.path.join(_HERE, "libkfsp.so"))     # KFSP_LIB: an alternative build (A/B measurements)
NCCL_ID_BYTES = 128


class Options(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("m_max", "m_min", "ideg", "n_init_onestep", "fsp_reject_limit", "mxstep", "mxreject",
                 "enable_drop", "enable_expand", "max_molecules", "device", "spmv_variant")] + \
               [("max_states", C.c_int64)] + \
               [(n, C.c_double) for n in ("delta", "gamma", "break_tol", "drop_tol0", "drop_deriv_tol", "drop_fraction")] + \
               [("seed", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("nmult", "nexph", "nscale", "nstep", "nreject", "ibrkflag", "mbrkdwn", "iflag")] + \
               [(n, C.c_double) for n in
                ("step_min", "step_max", "x_error", "s_error", "tbrkdwn", "t_now", "hump", "beta_ratio")] + \
               [(n, C.c_int64) for n in ("n_expand", "n_drop", "n_final", "n_max", "kernel_launches")] + \
               [(n, C.c_double) for n in ("device_seconds", "spmv_seconds", "wall_seconds")] + \
               [("spmv_launches", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TraceRow(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("t_now", "t_step", "t_new", "wsum", "err_loc", "beta")] + \
               [(n, C.c_int32) for n in ("m", "n_step", "n_after", "flags", "nmult", "nexph")]


PROPENSITY_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double), C.c_void_p)

_vp, _i32p, _i64p, _dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/kfsp.h declaration by declaration
SIGNATURES = {
    "kfsp_version": (C.c_char_p, []),
    "kfsp_status_string": (C.c_char_p, [C.c_int]),
    "kfsp_default_options": (C.c_int, [C.POINTER(Options)]),
    "kfsp_model_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "kfsp_model_load": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "kfsp_model_free": (C.c_int, [_vp]),
    "kfsp_model_dims": (C.c_int, [_vp, _i32p, _i32p, _i32p]),
    "kfsp_model_get_stoichiometry": (C.c_int, [_vp, _i32p]),
    "kfsp_model_set_stoichiometry": (C.c_int, [_vp, _i32p]),
    "kfsp_model_species_name": (C.c_int, [_vp, C.c_int32, C.c_char_p, C.c_int32]),
    "kfsp_model_parameter_name": (C.c_int, [_vp, C.c_int32, C.c_char_p, C.c_int32]),
    "kfsp_model_reset_parameters": (C.c_int, [_vp, _dp, C.c_int32]),
    "kfsp_model_set_propensity_string": (C.c_int, [_vp, C.c_int32, C.c_char_p]),
    "kfsp_model_set_propensity_bytecode": (C.c_int, [_vp, C.c_int32, _i32p, C.c_int32, _dp, C.c_int32]),
    "kfsp_model_get_propensity_bytecode": (C.c_int, [_vp, C.c_int32, _i32p, _i32p, _dp, _i32p]),
    "kfsp_model_set_custom_propensity": (C.c_int, [_vp, PROPENSITY_FN, _vp]),
    "kfsp_model_propensity": (C.c_int, [_vp, _i32p, C.c_int32, _dp]),
    "kfsp_model_propensity_factored": (C.c_int, [_vp, _i32p, C.c_int32, _dp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "kfsp_model_custom_structure": (C.c_int, [_vp, C.c_int32, _i32p, C.POINTER(C.c_int32)]),
    "kfsp_create": (C.c_int, [C.POINTER(Options), C.POINTER(_vp)]),
    "kfsp_destroy": (C.c_int, [_vp]),
    "kfsp_set_model": (C.c_int, [_vp, _vp]),
    "kfsp_solve": (C.c_int, [_vp, C.c_double, C.c_int64, _i32p, _dp, C.c_double, C.c_double, C.c_int32,
                             _i64p, _i32p, _dp, C.c_int64, C.POINTER(Stats)]),
    "kfsp_solve_resident": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_int32, C.POINTER(Stats)]),
    "kfsp_trace_length": (C.c_int, [_vp, _i64p]),
    "kfsp_trace_get": (C.c_int, [_vp, C.POINTER(TraceRow), C.c_int64]),
    "kfsp_fsp_init": (C.c_int, [_vp, C.c_int64, _i32p]),
    "kfsp_fsp_init_box": (C.c_int, [_vp, _i32p]),
    "kfsp_fsp_onestep": (C.c_int, [_vp]),
    "kfsp_fsp_ssa": (C.c_int, [_vp, C.c_double]),
    "kfsp_fsp_drop": (C.c_int, [_vp, C.c_double, _i32p, _dp, _i64p]),
    "kfsp_fsp_size": (C.c_int, [_vp, _i64p]),
    "kfsp_fsp_set_vector": (C.c_int, [_vp, _dp, C.c_int64]),
    "kfsp_fsp_get": (C.c_int, [_vp, _i32p, _i32p, _dp, _dp, _dp]),
    "kfsp_fsp_index": (C.c_int, [_vp, C.c_int64, _i32p, _i32p]),
    "kfsp_fsp_probability": (C.c_int, [_vp, C.c_int64, _i32p, _dp]),
    "kfsp_matvec": (C.c_int, [_vp, _dp, _dp]),
    "kfsp_matvec_device": (C.c_int, [_vp, _vp, _vp, C.c_int32, _dp]),
    "kfsp_arnoldi": (C.c_int, [_vp, _dp, C.c_int32, _dp, _dp, _i32p, _dp]),
    "kfsp_expm": (C.c_int, [_vp, C.c_int32, C.c_double, _dp, C.c_int32, _dp, _i32p, _dp]),
    "kfsp_combine": (C.c_int, [_vp, C.c_int64, C.c_int32, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp]),
    "kfsp_dist_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "kfsp_dist_init": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "kfsp_dist_partition": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, _i64p, _i64p]),
    "kfsp_dist_owner": (C.c_int, [C.c_int64, C.c_int32, C.c_int64, _i32p]),
    "kfsp_lattice_partition": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p]),
    "kfsp_lattice_kernel": (C.c_int, [C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _i32p]),
    "kfsp_model_info": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p]),
    "kfsp_repl_partition": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int64, _i64p, _i64p, _i32p]),
    "kfsp_dist_exchange_stats": (C.c_int, [_vp, _i64p, _dp, _dp, C.c_int32]),
    "kfsp_dist_info": (C.c_int, [_vp, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p]),
    "kfsp_device_alloc": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp)]),
    "kfsp_device_free": (C.c_int, [_vp, _vp]),
    "kfsp_device_upload": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "kfsp_device_download": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "kfsp_device_vector": (C.c_int, [_vp, C.POINTER(_vp)]),
    "kfsp_flush_l2": (C.c_int, [_vp]),
    "kfsp_set_profiling": (C.c_int, [_vp, C.c_int32]),
    "kfsp_set_blocking_sync": (C.c_int, [_vp, C.c_int32]),
    "kfsp_fsp_set_vector_device": (C.c_int, [_vp, _vp, C.c_int64]),
    "kfsp_launch_count": (C.c_int, [_vp, _i64p]),
    "kfsp_profile_get": (C.c_int, [_vp, _dp, _i64p, _i64p]),
    "kfsp_spmv_launch_counts": (C.c_int, [_vp, _i64p]),
    "kfsp_phase_seconds": (C.c_int, [_vp, _dp]),
}

_LIB = None


class KfspError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = status
        msg = lib().kfsp_status_string(status).decode()
        super().__init__("%s: %s (status %d)" % (where or "libkfsp", msg, status))


def build(extra=()):
    """Compile libkfsp.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ)
    try:
        import nvidia.nccl as _n  # noqa: F401
        base = os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0]
        if os.path.exists(os.path.join(base, "include", "nccl.h")):
            env.setdefault("KFSP_NCCL_INC", os.path.join(base, "include"))
            env.setdefault("KFSP_NCCL_LIB", os.path.join(base, "lib"))
    except Exception:
        pass
    subprocess.check_call([os.path.join(_HERE, "csrc", "build.sh"), *extra], env=env)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("libkfsp.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "krylovfspssa_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def check(status, where=""):
    if status != 0:
        raise KfspError(status, where)
    return status
