"""GPU parity of the state-space routines against the oracle:
state lists, indices and ADJ bit-exact; OFFDIAG/DIAG bit-exact where the propensity has no
transcendental operation or is served from a host-built table."""
import numpy as np
import pytest

import oracle
from gpu_common import make

pytestmark = pytest.mark.gpu


def assert_same_fsp(h, of, exact_values=True):
    d, o = h.get(), of.get()
    assert d["states"].shape == o["states"].shape
    assert np.array_equal(d["states"], o["states"])
    assert np.array_equal(d["adj"], o["adj"])
    if exact_values:
        assert np.array_equal(d["offdiag"], o["offdiag"])
        assert np.array_equal(d["diag"], o["diag"])
    else:
        assert np.allclose(d["offdiag"], o["offdiag"], rtol=1e-15, atol=0)
        assert np.allclose(d["diag"], o["diag"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("name", ["toggle", "repressilator", "goutsias", "birth_death"])
def test_matrix_starter_and_onestep(name):
    h, om, x0 = make(name)
    of = oracle.Fsp(om)
    h.fsp_init([x0])
    of.set_states([x0])
    assert of.matrix_starter() == 0
    assert_same_fsp(h, of)
    for _ in range(5):
        h.onestep()
        assert of.onestep() == 0
        assert h.size == of.size
    assert_same_fsp(h, of)
    h.close()


def test_matrix_starter_many_states_given_order():
    h, om, _ = make("toggle")
    rng = np.random.default_rng(5)
    pts = rng.permutation(60 * 60)[:900]
    states = np.stack([pts % 60, pts // 60], axis=1).astype(np.int32)
    of = oracle.Fsp(om)
    h.fsp_init(states)
    of.set_states(states)
    of.matrix_starter()
    assert_same_fsp(h, of)
    idx = h.index(states)
    assert np.array_equal(idx, np.arange(1, len(states) + 1))
    assert h.index([[9999, 9999]])[0] == 0
    h.close()


@pytest.mark.parametrize("name,ts", [("toggle", 0.05), ("toggle", 0.5), ("repressilator", 0.1), ("goutsias", 5.0)])
def test_ssa_extender_order_exact(name, ts):
    h, om, x0 = make(name, seed=777)
    of = oracle.Fsp(om)
    rng = oracle.Rng(oracle.Rng.PHILOX, 777)
    h.fsp_init([x0])
    of.set_states([x0])
    of.matrix_starter()
    for _ in range(5):
        h.onestep(); of.onestep()
    for rep in range(3):
        h.ssa(ts)
        of.ssa(ts, rng)
        assert h.size == of.size, (rep, h.size, of.size)
        h.onestep(); of.onestep()
        assert h.size == of.size
    assert_same_fsp(h, of)
    h.close()


def test_drop_states():
    h, om, x0 = make("toggle")
    of = oracle.Fsp(om)
    h.fsp_init([x0]); of.set_states([x0]); of.matrix_starter()
    for _ in range(12):
        h.onestep(); of.onestep()
    n = of.size
    rng = np.random.default_rng(0)
    w = np.exp(-0.6 * np.arange(n)) * rng.uniform(0.5, 1.0, n)
    w /= w.sum()
    h.set_vector(w)
    dropped, tol, cnt = h.drop(1e-7)
    did, w2, otol, ocnt = of.drop(w.copy(), 1e-7)
    assert (dropped, cnt) == (did, ocnt) and tol == otol
    assert dropped == 1
    assert_same_fsp(h, of)
    assert np.array_equal(h.get()["vector"], w2)
    # after compaction the dropped neighbours are explorable again
    h.onestep(); of.onestep()
    assert_same_fsp(h, of)
    h.close()


def test_drop_not_triggered_below_ten_percent():
    h, om, x0 = make("birth_death")
    of = oracle.Fsp(om)
    st = [[i] for i in range(50)]
    h.fsp_init(st); of.set_states(st); of.matrix_starter()
    w = np.full(50, 1.0 / 50); w[-2:] = 1e-12
    h.set_vector(w)
    dropped, tol, cnt = h.drop(1e-9)
    did, _, otol, ocnt = of.drop(w.copy(), 1e-9)
    assert dropped == did == 0 and cnt == ocnt and tol == otol
    assert h.size == 50
    h.close()


def test_error_codes():
    import krylovfspssa_b200 as k
    h, om, _ = make("toggle")
    with pytest.raises(k.KfspError) as e:
        h.fsp_init([[0, 0], [0, 0]])
    assert e.value.status == -11
    with pytest.raises(k.KfspError) as e:
        h.fsp_init([[-1, 0]])
    assert e.value.status == -11
    h.close()
    h, om, _ = make("toggle", max_states=30)
    h.fsp_init([[0, 0]])
    with pytest.raises(k.KfspError) as e:
        for _ in range(10):
            h.onestep()
    assert e.value.status == -10
    h.close()
