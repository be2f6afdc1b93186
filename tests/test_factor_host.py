"""Host logic of the index-only SpMV (spmv_variant = 2): propensity programs factored into single-species terms
(csrc/model_host.cpp: factor_program) must reproduce MODEL%PROPENSITY (ModelModule.f90:163-199) bit for bit -- the device
combines the tabulated terms with the same IEEE operations.  CPU only: the C-ABI library, no compute on a GPU."""
import os

import numpy as np
import pytest

import krylovfspssa_b200 as k
from gpu_common_cases import CASES


@pytest.mark.parametrize("name", sorted(CASES))
def test_factored_form_equals_propensity(name):
    fname, params, x0 = CASES[name]
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), fname))
    m.reset_parameters(params)
    S, R = m.nspecies, m.nreactions
    rng = np.random.default_rng(11)
    states = np.vstack([np.zeros((1, S), dtype=np.int32), rng.integers(0, 400, size=(300, S)).astype(np.int32),
                        rng.integers(0, 10001, size=(50, S)).astype(np.int32)])
    for r in range(1, R + 1):
        for st in states:
            v, nt, no = m.propensity_factored(st, r)
            assert v == m.propensity(st, r), (name, r, st)
            assert 1 <= nt <= 4 and no <= 8                      # what the device structure holds (common.cuh: FacModel)


def test_factored_shapes():
    """Goutsias: eight single-species propensities (one table each) and two bimolecular ones (product of two tables)"""
    fname, params, _ = CASES["goutsias"]
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), fname))
    m.reset_parameters(params)
    st = np.array([3, 4, 5, 2, 1, 1], dtype=np.int32)
    shapes = [m.propensity_factored(st, r)[1:] for r in range(1, 11)]
    assert shapes == [(1, 1), (1, 1), (1, 1), (1, 1), (2, 3), (1, 1), (2, 3), (1, 1), (1, 1), (1, 1)]


def test_unfactorable_is_refused():
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle.input"))
    m.reset_parameters([1.0, 100.0, 1.0, 1.0, 100.0, 1.0])
    st = np.array([3, 4], dtype=np.int32)
    m.set_propensity(1, "kx/(1.0 + X*Y)")                         # a division across two species
    with pytest.raises(k.KfspError):
        m.propensity_factored(st, 1)
    m.set_propensity(1, "(X*Y)**2")                               # pow across two species
    with pytest.raises(k.KfspError):
        m.propensity_factored(st, 1)
    m.set_propensity(1, "kx*X*Y + 0.5*Y")                         # sums and products: the general postfix form
    v, nt, no = m.propensity_factored(st, 1)
    assert v == m.propensity(st, 1) and nt == 3 and no == 5
    m.set_propensity(1, "kx*X*Y + 0.5*Y - X*(Y - 1.0)")           # five terms: more than the device structure holds
    with pytest.raises(k.KfspError):
        m.propensity_factored(st, 1)
    m.set_propensity(1, "X*(1.0/Y)")                              # fparser: a division by zero makes the WHOLE value 0 ...
    assert m.propensity(np.array([3, 0], dtype=np.int32), 1) == 0.0
    with pytest.raises(k.KfspError):                              # ... which a product of terms cannot express: refused
        m.propensity_factored(np.array([3, 0], dtype=np.int32), 1)
