import os

import numpy as np

import krylovfspssa_b200 as k
import oracle

from gpu_common_cases import CASES, GOLDEN_RUNS, GOUTSIAS  # noqa: F401


def make(name, max_states=400000, **opt):
    """(device handle, oracle model, oracle fsp) for one of the shipped models."""
    fname, params, x0 = CASES[name]
    path = os.path.join(k.models_dir(), fname)
    model = k.CME_MODEL().load(path)
    model.reset_parameters(params)
    h = k.KrylovFspHandle(model, max_states=max_states, **opt)
    om = oracle.Model.load(path, params)
    return h, om, x0


def rel1(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).sum() / max(np.abs(np.asarray(b)).sum(), 1e-300)
