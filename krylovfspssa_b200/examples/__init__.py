"""The reference's three example drivers (examples/toggle.f90, examples/repressilator.f90,
examples/transcr6d.f90) on the device path.  Like the Fortran programs they CREATE the model by hand,
fill STOICHIOMETRY, point CUSTOMPROP at a hard-coded host function and call CME_SOLVE; the host
functions are compiled C (driver_props.c) so a solve is not throttled by Python callbacks."""
import ctypes as C
import os
import subprocess

import numpy as np

from .. import _lib
from ..host import CME_MODEL

_HERE = os.path.dirname(os.path.abspath(__file__))
PROPS_PATH = os.path.join(_HERE, "libkfsp_examples.so")
_PROPS = None

# name -> (S, R, P, stoichiometry as the RESHAPE source (species fastest), C symbol, parameters, x0, t, FSPTOL, KRYTOL)
GOUTSIAS_STOICH = np.zeros((10, 6), dtype=np.int32)     # examples/transcr6d.f90:91-133, [reaction, species]
for _k, _row in enumerate([{0: 1}, {0: -1}, {2: 1}, {2: -1}, {3: -1, 1: -1, 4: 1}, {3: 1, 1: 1, 4: -1},
                           {4: -1, 1: -1, 5: 1}, {4: 1, 1: 1, 5: -1}, {0: -2, 1: 1}, {0: 2, 1: -1}]):
    for _s, _v in _row.items():
        GOUTSIAS_STOICH[_k, _s] = _v

DRIVERS = {
    # examples/toggle.f90:14,23-26,42
    "toggle": dict(S=2, R=4, P=6, stoich=np.array([1, 0, -1, 0, 0, 1, 0, -1], dtype=np.int32).reshape(4, 2),
                   symbol="kfsp_example_toggle_propensity", params=[1.0, 100.0, 1.0, 1.0, 100.0, 1.0],
                   x0=[0, 0], t=100.0, fsp_tol=1e-4, exp_tol=1e-8, oracle_kind=3),
    # examples/repressilator.f90:14,23-26,37
    "repressilator": dict(S=3, R=6, P=3,
                          stoich=np.array([1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1], dtype=np.int32).reshape(6, 3),
                          symbol="kfsp_example_repressilator_propensity", params=[100.0, 25.0, 1.0],
                          x0=[22, 0, 0], t=10.0, fsp_tol=1e-4, exp_tol=1e-14, oracle_kind=2),
    # examples/transcr6d.f90:16,23-32,50
    "transcr6d": dict(S=6, R=10, P=10, stoich=GOUTSIAS_STOICH, symbol="kfsp_example_goutsias_propensity",
                      params=[0.043, 0.0007, 0.0715, 0.0039, 0.0199264663575241, 0.4791, 0.000199264663575241,
                              0.8765 * 1.0e-11, 0.0830269431563506104, 0.5],
                      x0=[2, 6, 0, 2, 0, 0], t=300.0, fsp_tol=1e-6, exp_tol=1e-8, oracle_kind=1),
}


def build_props(force=False):
    src = os.path.join(_HERE, "driver_props.c")
    if force or not os.path.exists(PROPS_PATH) or os.path.getmtime(PROPS_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", PROPS_PATH, src, "-lm"])
    return PROPS_PATH


def props():
    global _PROPS
    if _PROPS is None:
        _PROPS = C.CDLL(build_props())
    return _PROPS


def customprop(name):
    """ctypes function pointer of one driver's CUSTOMPROP."""
    fn = getattr(props(), DRIVERS[name]["symbol"])
    return C.cast(fn, _lib.PROPENSITY_FN)


def driver_model(name):
    """CALL MODEL%CREATE(S,R,P); MODEL%STOICHIOMETRY = ...; MODEL%CUSTOMPROP => ...; RESET_PARAMETERS(...)."""
    d = DRIVERS[name]
    m = CME_MODEL().create(d["S"], d["R"], d["P"])
    m.stoichiometry = np.ascontiguousarray(d["stoich"].T)
    m.set_customprop(customprop(name))
    m.reset_parameters(d["params"])
    m.loaded = True
    return m
