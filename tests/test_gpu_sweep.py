"""Parameter sweep on one GPU: one handle, RESET_PARAMETERS + kfsp_set_model between solves, every set bit-identical to
the oracle solved with that set."""
import os

import numpy as np
import pytest

import krylovfspssa_b200 as k
import oracle
from krylovfspssa_b200 import sweep

pytestmark = pytest.mark.gpu


def test_sweep_matches_oracle_per_set():
    path = os.path.join(k.models_dir(), "toggle.input")
    sets = [[1.0, 100.0, 1.0, 1.0, 100.0, 1.0], [2.0, 60.0, 1.0, 1.0, 80.0, 1.5], [1.0, 100.0, 1.0, 1.0, 100.0, 1.0]]
    model = k.CME_MODEL().load(path)
    res = sweep.run_share(model, sets, [0, 0], 2.0, 1e-4, 1e-10, max_states=400000, seed=12345, device=0)
    assert sorted(res) == [0, 1, 2]
    for i, ps in enumerate(sets):
        ref = oracle.solve(oracle.Model.load(path, ps), [[0, 0]], [1.0], 2.0, 1e-4, 1e-10, seed=12345, reproducible=1)
        assert np.array_equal(res[i]["states"], ref["states"]) and np.array_equal(res[i]["vector"], ref["vector"]), i
    assert np.array_equal(res[0]["vector"], res[2]["vector"])            # the handle carries nothing over between sets
    summ = sweep.gather_summaries(res, 3)
    assert [s["n"] for s in summ] == [len(res[i]["vector"]) for i in range(3)]


def test_concurrent_sweep_is_bit_identical_to_the_sequential_one():
    """several handles (own stream, buffers, model copy) driven by one host thread each on ONE GPU: same bits per set"""
    path = os.path.join(k.models_dir(), "toggle.input")
    rng = np.random.default_rng(3)
    sets = [[1.0, 60.0 + 40.0 * rng.random(), 1.0, 1.0, 60.0 + 40.0 * rng.random(), 1.0] for _ in range(12)]
    model = k.CME_MODEL().load(path)
    seq = sweep.run_share(model, sets, [0, 0], 2.0, 1e-4, 1e-10, max_states=200000, seed=12345, device=0)
    par = sweep.run_share(model, sets, [0, 0], 2.0, 1e-4, 1e-10, max_states=200000, seed=12345, device=0, concurrency=4)
    assert sorted(par) == sorted(seq) == list(range(12))
    for i in range(12):
        assert np.array_equal(seq[i]["states"], par[i]["states"]) and np.array_equal(seq[i]["vector"], par[i]["vector"]), i
        assert np.array_equal(seq[i]["trace"]["i"], par[i]["trace"]["i"])


def test_blocking_host_waits_change_nothing_but_the_wait():
    """kfsp_set_blocking_sync: the host thread sleeps on a blocking event instead of spinning in the driver; same bits"""
    path = os.path.join(k.models_dir(), "toggle.input")
    model = k.CME_MODEL().load(path)
    model.reset_parameters([1.0, 100.0, 1.0, 1.0, 100.0, 1.0])
    a = k.KrylovFspHandle(model, max_states=200000, seed=12345)
    ra = a.solve(2.0, [[0, 0]], [1.0], 1e-4, 1e-10)
    a.close()
    b = k.KrylovFspHandle(model, max_states=200000, seed=12345)
    b.set_blocking_sync(True)
    rb = b.solve(2.0, [[0, 0]], [1.0], 1e-4, 1e-10)
    b.close()
    assert np.array_equal(ra["states"], rb["states"]) and np.array_equal(ra["vector"], rb["vector"])
