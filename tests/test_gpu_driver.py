"""The file-driven driver end to end on the GPU (generic TestSolverFromFile + result writer)."""
import os

import numpy as np
import pytest

import krylovfspssa_b200 as k
import oracle
from krylovfspssa_b200 import driver

pytestmark = pytest.mark.gpu


def test_driver_writes_what_the_oracle_computes(tmp_path):
    path = os.path.join(k.models_dir(), "toggle.input")
    out_file = str(tmp_path / "toggle.npz")
    driver.main([path, "--params", "1,100,1,1,100,1", "--x0", "0,0", "--t", "5", "--fsptol", "1e-4", "--krytol", "1e-10",
                 "--max-states", "400000", "--verbosity", "0", "--out", out_file])
    r = driver.read_result(out_file)
    ref = oracle.solve(oracle.Model.load(path, [1, 100, 1, 1, 100, 1]), [[0, 0]], [1.0], 5.0, 1e-4, 1e-10, reproducible=1)
    assert np.array_equal(r["states"], ref["states"]) and np.array_equal(r["vector"], ref["vector"])
    assert np.array_equal(r["trace"]["i"], ref["trace_i"])
    for name, m in r["marginals"].items():
        assert abs(m.sum() - ref["vector"].sum()) < 1e-12
    assert r["stats"]["nmult"] == ref["stats"]["nmult"] and r["meta"]["t"] == 5.0
