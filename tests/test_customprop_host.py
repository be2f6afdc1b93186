"""CUSTOMPROP host callbacks (ModelModule.f90:6-12,31,188-190) without a GPU: the compiled driver
functions of krylovfspssa_b200/examples (examples/*.f90 restated in C) against the oracle's own
restatements, through MODEL%PROPENSITY on both sides."""
import itertools

import numpy as np
import pytest

import krylovfspssa_b200 as k
import oracle
from krylovfspssa_b200 import examples


def grid(name):
    if name == "toggle":
        return itertools.product(range(0, 60, 3), range(0, 60, 7))
    if name == "repressilator":
        return itertools.product(range(0, 40, 5), range(0, 40, 7), range(0, 40, 9))
    return itertools.product((0, 1, 2, 9), (0, 3, 6), (0, 1, 4), (0, 1, 2), (0, 1, 2), (0, 1, 2))


@pytest.mark.parametrize("name", sorted(examples.DRIVERS))
def test_driver_callbacks_match_oracle(name):
    d = examples.DRIVERS[name]
    m = examples.driver_model(name)
    assert (m.nspecies, m.nreactions, m.nparameters) == (d["S"], d["R"], d["P"])
    om = oracle.Model(d["S"], d["R"], d["P"], d["stoich"], d["params"])
    om.set_custom(d["oracle_kind"])
    acc = 0.0
    for st in grid(name):
        for r in range(1, d["R"] + 1):
            a, b = m.propensity(list(st), r), om.propensity(list(st), r)
            assert a == b, (st, r, a, b)
            acc += abs(a - b)
    assert acc == 0.0           # the 'ACCUMULATED ERROR' of test/TestModelParser.f90:45,77


def test_python_callable_customprop():
    m = k.CME_MODEL().create(1, 2, 2)
    m.stoichiometry = np.array([[1, -1]], dtype=np.int32)
    m.reset_parameters([20.0, 1.0])
    m.set_customprop(lambda state, reaction, p: p[0] if reaction == 1 else p[1] * state[0])
    assert m.propensity([7], 1) == 20.0
    assert m.propensity([7], 2) == 7.0


def test_goutsias_stoichiometry_matches_input_file():
    import os
    fm = k.CME_MODEL().load(os.path.join(k.models_dir(), "goutsias.input"))
    assert np.array_equal(fm.stoichiometry, examples.driver_model("transcr6d").stoichiometry)


def test_custom_structure_probe_of_the_example_drivers():
    """kfsp_model_custom_structure (host only): which species each reaction of an opaque callback reads.  toggle.f90 and
    repressilator.f90 read one species per reaction (kind 1: device tables); transcr6d.f90 reads two in reactions 5 and 7, as
    bilinear mass action (c * x_a) * x_b (kind 2: tables + three-operation byte code)."""
    sp, kind = examples.driver_model("toggle").custom_structure()
    assert kind == 1 and sp == [1, 0, 0, 1]
    sp, kind = examples.driver_model("repressilator").custom_structure()
    assert kind == 1 and sp == [1, 0, 2, 1, 0, 2]
    sp, kind = examples.driver_model("transcr6d").custom_structure()
    assert kind == 2 and sp == [2, 0, 4, 2, -2, 4, -2, 5, 0, 1]


def test_custom_structure_probe_rejects_what_it_cannot_reproduce_bit_for_bit():
    m = k.CME_MODEL().create(2, 2, 2)
    m.stoichiometry = np.array([[1, -1], [0, 0]], dtype=np.int32)
    m.reset_parameters([20.0, 1.0])
    # a second species that matters only far from the axes: invisible to the one-species-at-a-time probes, caught by the
    # bit-for-bit verification on random states
    m.set_customprop(lambda st, r, p: p[0] + (1.0 if (st[0] > 1000 and st[1] > 1000) else 0.0) if r == 1 else p[1] * st[0])
    assert m.custom_structure(max_molecules=2000)[1] == 0
    # two species, but not (c * x_a) * x_b: a Hill-type coupling, and mass action evaluated as c * (x_a * x_b) with a coefficient
    # for which the two roundings differ
    m.set_customprop(lambda st, r, p: p[0] * st[0] / (1.0 + st[1]) if r == 1 else p[1] * st[0])
    assert m.custom_structure(max_molecules=2000)[1] == 0
    m.set_customprop(lambda st, r, p: 0.1 * (float(st[0]) * float(st[1])) if r == 1 else p[1] * st[0])
    sp, kind = m.custom_structure(max_molecules=2000)
    assert kind == 0 and sp[0] == -2
    # the same product the way a compiler evaluates c*x*y, either operand first: recognised
    m.set_customprop(lambda st, r, p: (0.1 * float(st[1])) * float(st[0]) if r == 1 else p[1] * st[0])
    sp, kind = m.custom_structure(max_molecules=2000)
    assert kind == 2 and sp == [-2, 0]
    m.set_customprop(lambda st, r, p: p[0] if r == 1 else p[1] * st[0])
    sp, kind = m.custom_structure(max_molecules=2000)
    assert kind == 1 and sp == [0, 0]
