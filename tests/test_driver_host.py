"""Result writer / argument handling of the file-driven driver (no GPU)."""
import numpy as np

from krylovfspssa_b200 import driver


def fake_out():
    states = np.array([[0, 0], [1, 0], [0, 2], [1, 2], [3, 1]], dtype=np.int32)
    p = np.array([0.1, 0.2, 0.3, 0.15, 0.25])
    return dict(states=states, vector=p, iflag=0, stats=dict(nstep=3, nmult=40, n_expand=1, n_drop=0, device_seconds=0.5),
                trace=dict(d=np.zeros((3, 6)), i=np.zeros((3, 6), dtype=np.int32)))


def test_marginals():
    o = fake_out()
    mx, my = driver.marginals(o["states"], o["vector"])
    assert np.allclose(mx, [0.4, 0.35, 0.0, 0.25]) and np.allclose(my, [0.3, 0.25, 0.45])
    assert abs(mx.sum() - o["vector"].sum()) < 1e-15 and abs(my.sum() - o["vector"].sum()) < 1e-15


def test_result_file_round_trip(tmp_path):
    o = fake_out()
    path = str(tmp_path / "r.npz")
    driver.write_result(path, o, ["X", "Y"], dict(model="m.input", params=[1.0, 2.0], x0=[0, 0], t=1.0))
    r = driver.read_result(path)
    assert np.array_equal(r["states"], o["states"]) and np.array_equal(r["vector"], o["vector"])
    assert r["stats"]["nmult"] == 40 and r["meta"]["species"] == ["X", "Y"] and r["meta"]["iflag"] == 0
    assert np.allclose(r["marginals"]["Y"], [0.3, 0.25, 0.45])


def test_arguments():
    a = driver.parse_args(["toggle.input", "--params", "1,100,1,1,100,1", "--x0", "0,0", "--t", "1000", "--fsptol", "1e-4"])
    assert a.params == [1, 100, 1, 1, 100, 1] and a.x0 == [0, 0] and a.t == 1000.0 and a.krytol == 1e-8 and a.seed == 12345
