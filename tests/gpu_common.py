import os

import numpy as np

import krylovfspssa_b200 as k
import oracle

GOUTSIAS = [0.043, 0.0007, 0.0715, 0.0039, 0.0199264663575241, 0.4791, 0.000199264663575241,
            0.8765e-11, 0.0830269431563506104, 0.5]
CASES = {
    "toggle": ("toggle.input", [1.0, 100.0, 1.0, 1.0, 100.0, 1.0], [0, 0]),
    "repressilator": ("repressilator.input", [100.0, 100.0, 100.0, 1.0, 1.0, 1.0], [22, 0, 0]),
    "goutsias": ("goutsias.input", GOUTSIAS, [2, 6, 0, 2, 0, 0]),
    "birth_death": ("birth_death.input", [20.0, 1.0], [0]),
    "toggle_test": ("toggle_test.input", [5000.0, 1600.0, 1.0, 1.0], [0, 0]),
}


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
