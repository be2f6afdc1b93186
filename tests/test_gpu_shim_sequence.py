"""The reference-side binding's call sequence from a COMPILED host (tests/host/shim_sequence.c, the C twin of
krylovfspssa_b200/fortran/kfsp_c_binding.f90: CME_SOLVE of KrylovSolver.f90:7-36 over the C ABI): column-major stoichiometry,
byte code handed over per reaction or a CUSTOMPROP trampoline with a module-level model pointer, FSP_OUT%STATE passed as both
states_in and states_out with capacity MAX_SIZE.  Result = the committed oracle fixture, bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import krylovfspssa_b200 as k
from gpu_common_cases import CASES, GOLDEN_RUNS
from shim_build import build_shim_sequence

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run(exe, mode, tag, out):
    name, t, ftol, ktol, seed = GOLDEN_RUNS[tag]
    fname, params, x0 = CASES[name]
    cmd = [exe, mode, os.path.join(k.models_dir(), fname), repr(t), repr(ftol), repr(ktol), "400000", str(seed), str(out)]
    cmd += [str(v) for v in x0] + [repr(v) for v in params]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    n = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
    iflag, nstep = np.frombuffer(raw[8:16], dtype=np.int32)
    S = len(x0)
    states = np.frombuffer(raw[16:16 + 4 * n * S], dtype=np.int32).reshape(n, S)
    w = np.frombuffer(raw[16 + 4 * n * S:], dtype=np.float64)
    assert len(w) == n and iflag == 0 and nstep > 0
    return states, w


@pytest.mark.parametrize("tag", ["toggle_t20", "goutsias_t30"])
def test_compiled_host_bytecode_path_equals_oracle_fixture(tag, tmp_path):
    exe = build_shim_sequence(tmp_path)
    states, w = run(exe, "bytecode", tag, tmp_path / "o.bin")
    g = np.load(os.path.join(HERE, "golden", tag + ".npz"))
    assert np.array_equal(states, g["states"])
    assert np.array_equal(w, g["vector"])


def test_compiled_host_customprop_trampoline_equals_oracle_fixture(tmp_path):
    """MODEL%CUSTOMPROP => f through the trampoline (ctx = NULL, model found through the module-level pointer), f written with the
    expressions of the byte-code toggle: the same fixture again"""
    exe = build_shim_sequence(tmp_path)
    states, w = run(exe, "custom", "toggle_t20", tmp_path / "o.bin")
    g = np.load(os.path.join(HERE, "golden", "toggle_t20.npz"))
    assert np.array_equal(states, g["states"])
    assert np.array_equal(w, g["vector"])
