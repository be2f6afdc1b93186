"""krylovfspssa_b200 -- B200-native Krylov-FSP-SSA time stepping.

Python mirror of the reference's Fortran host API for the hot path
(voduchuy/KrylovFspSsa):

    CME_MODEL                 src/model/ModelModule.f90:14-42
    FINITE_STATE_PROJECTION   src/state_space/StateSpace.f90:19-45
    CME_SOLVE                 src/fsp/KrylovSolver.f90:7-36

Everything numerical happens in libkfsp.so (hand-written CUDA for sm_100a) behind the C ABI
of include/kfsp.h; this package only marshals arrays.  There is no CPU fallback.
"""
from .host import (CME_MODEL, FINITE_STATE_PROJECTION, CME_SOLVE, cme_solve, KrylovFspHandle, NMAX,
                   MAXNUMBERMOLECULES, default_options, models_dir)
from ._lib import KfspError, Options, Stats, build, lib, LIB_PATH

__all__ = ["CME_MODEL", "FINITE_STATE_PROJECTION", "CME_SOLVE", "cme_solve", "KrylovFspHandle", "NMAX",
           "MAXNUMBERMOLECULES", "default_options", "models_dir", "KfspError", "Options", "Stats", "build", "lib",
           "LIB_PATH"]
