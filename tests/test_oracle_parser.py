"""Pins the oracle's parser / model reader (oracle/fparser.py, oracle/model_input.py).

The only result-bearing check the reference has for this code is the accumulated
|P1-P2| of test/TestModelParser.f90:31-45 (parsed propensity vs the closed-form PROP
of :80-102 on the 50x50x4 grid, expected 0); it is reproduced here exactly."""
import math
import os

import numpy as np
import pytest

import oracle
from oracle import fparser, model_input


def test_prop_grid_known_answer(models_dir):
    # test/TestModelParser.f90:13-45
    m = oracle.Model.load(os.path.join(models_dir, "toggle_test.input"), [5000.0, 1600.0, 1.0, 1.0])
    ref = oracle.Model(2, 4, 0, m.stoich)
    ref.set_custom(oracle.CUSTOM_PARSER_TEST)
    err = 0.0
    for i in range(1, 51):
        for j in range(1, 51):
            for r in range(1, 5):
                err += abs(ref.propensity([i, j], r) - m.propensity([i, j], r))
    assert err == 0.0


def test_prop_grid_python_evaluator(models_dir):
    im = model_input.load(os.path.join(models_dir, "toggle_test.input"))
    par = [5000.0, 1600.0, 1.0, 1.0]
    for i in (0, 1, 7, 50):
        for j in (0, 2, 49):
            v = [float(i), float(j)] + par
            assert im.programs[0].evaluate(v) == 5000.0 / (1.0 + math.pow(j, 2.5))
            assert im.programs[1].evaluate(v) == 1600.0 / (1.0 + math.pow(i, 1.5))
            assert im.programs[2].evaluate(v) == 1.0 * i
            assert im.programs[3].evaluate(v) == 1.0 * j


def test_operator_split_order():
    # FortranParser.f90:679-706: operators searched + - * / ^, scanning right to left
    V = ["a", "b", "c"]
    A, B, Cc = fparser.VarBegin, fparser.VarBegin + 1, fparser.VarBegin + 2
    assert fparser.Program("a*b/c", V).code == [A, B, Cc, fparser.cDiv, fparser.cMul]      # a*(b/c)
    assert fparser.Program("a/b*c", V).code == [A, B, fparser.cDiv, Cc, fparser.cMul]      # (a/b)*c
    assert fparser.Program("a-b-c", V).code == [A, B, fparser.cSub, Cc, fparser.cSub]      # (a-b)-c
    assert fparser.Program("a+b-c", V).code == [A, B, Cc, fparser.cSub, fparser.cAdd]      # a+(b-c)
    assert fparser.Program("a^b^c", V).code == [A, B, fparser.cPow, Cc, fparser.cPow]      # (a^b)^c
    assert fparser.Program("a**2", V).code == [A, fparser.cImmed, fparser.cPow]
    assert fparser.Program("-a*b", V).code == [A, B, fparser.cMul, fparser.cNeg]
    assert fparser.Program("exp(-a)+sqrt(b)", V).code == [A, fparser.cNeg, fparser.cExp, B, fparser.cSqrt, fparser.cAdd]
    p = fparser.Program("1.5d0*a + 2.0e-1", V)
    assert p.immed == [1.5, 0.2]
    assert p.evaluate([2.0, 0, 0]) == 3.2


def test_division_by_zero_yields_zero():
    # FortranParser.f90:217-224
    p = fparser.Program("a/b", ["a", "b"])
    assert p.evaluate([1.0, 0.0]) == 0.0


def test_unknown_variable_rejected():
    with pytest.raises(fparser.ParseError):
        fparser.Program("a*zz", ["a", "b"])


def test_stoichiometry_of_shipped_models(models_dir):
    g = model_input.load(os.path.join(models_dir, "goutsias.input"))
    # examples/transcr6d.f90:92-136 (hard-coded stoichiometry of the same network)
    M, D, RNA, DNA, DNAD, DNA2D = range(6)
    want = np.zeros((10, 6), dtype=int)
    want[0, M] = 1; want[1, M] = -1; want[2, RNA] = 1; want[3, RNA] = -1
    want[4, DNA] = -1; want[4, D] = -1; want[4, DNAD] = 1
    want[5, DNA] = 1; want[5, D] = 1; want[5, DNAD] = -1
    want[6, DNAD] = -1; want[6, D] = -1; want[6, DNA2D] = 1
    want[7, DNAD] = 1; want[7, D] = 1; want[7, DNA2D] = -1
    want[8, M] = -2; want[8, D] = 1; want[9, M] = 2; want[9, D] = -1
    assert np.array_equal(np.array(g.stoichiometry), want)
    t = model_input.load(os.path.join(models_dir, "toggle.input"))
    assert t.stoichiometry == [[1, 0], [-1, 0], [0, 1], [0, -1]]
    r = model_input.load(os.path.join(models_dir, "repressilator.input"))
    assert r.stoichiometry == [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]]


def test_reverse_arrow_and_coefficients():
    assert model_input.stoich_input(2, "2A + B -> 3B", ["A", "B"]) == [-2, 2]
    assert model_input.stoich_input(2, "A <- B", ["A", "B"]) == [1, -1]
    with pytest.raises(ValueError):
        model_input.stoich_input(2, "A + B", ["A", "B"])


def test_incomplete_model_rejected(tmp_path):
    # models/ge5d_model.input declares 14 reactions and lists 10
    p = tmp_path / "bad.input"
    p.write_text("nspecies\n1\nnreactions\n3\nnparameters\n1\nspecies\nX\nparameters\nk\nreactions\n0 -> X\n")
    with pytest.raises(ValueError):
        model_input.load(str(p))


def test_goutsias_parsed_vs_hardcoded(models_dir):
    # parsed c9*M*(M-1)/2.0d0 vs the integer M*(M-1)/2 of examples/transcr6d.f90:85
    tp = [0.043, 0.0007, 0.0715, 0.0039, 0.0199264663575241, 0.4791, 0.000199264663575241,
          0.8765e-11, 0.0830269431563506104, 0.5]
    a = oracle.Model.load(os.path.join(models_dir, "goutsias.input"), tp)
    b = oracle.Model(6, 10, 10, a.stoich, tp)
    b.set_custom(oracle.CUSTOM_GOUTSIAS)
    rng = np.random.default_rng(1)
    for _ in range(200):
        st = rng.integers(0, 40, size=6)
        for r in range(1, 11):
            x, y = a.propensity(st, r), b.propensity(st, r)
            assert abs(x - y) <= 4e-16 * max(abs(x), abs(y))
