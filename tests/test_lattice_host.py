"""Host logic of the matrix-free lattice variant without a GPU: slab partition, kernel dispatch for the reference's
models, and (gloo, world size 2) the halo plan + an independent restatement of the row formula against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import krylovfspssa_b200 as k
from krylovfspssa_b200._lib import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slabs_tile_the_slowest_species():
    L = lib()
    for nz in (1, 2, 7, 29, 1031, 10000):
        for p in (1, 2, 3, 4, 8):
            if nz < p:
                a, b = C.c_int32(), C.c_int32()
                assert L.kfsp_lattice_partition(nz, p, 0, C.byref(a), C.byref(b)) != 0
                continue
            prev = 0
            for r in range(p):
                a, b = C.c_int32(), C.c_int32()
                assert L.kfsp_lattice_partition(nz, p, r, C.byref(a), C.byref(b)) == 0
                assert a.value == prev and b.value - a.value in (nz // p, nz // p + 1) and b.value > a.value
                prev = b.value
            assert prev == nz


def kernel_of(model_file, table_species):
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), model_file))
    st = np.ascontiguousarray(m.stoichiometry.T, dtype=np.int32)          # [reaction, species] = species fastest
    ts = np.ascontiguousarray(table_species, dtype=np.int32)
    kind, mask = C.c_int32(), C.c_int32()
    rc = lib().kfsp_lattice_kernel(m.nspecies, m.nreactions, st.ctypes.data_as(C.POINTER(C.c_int32)),
                                   ts.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(kind), C.byref(mask))
    return rc, kind.value, mask.value


def test_kernel_dispatch_for_the_reference_models():
    # toggle_model.input: 0->X, X->0, 0->Y, Y->0 ; propensities read Y, X, X, Y
    assert kernel_of("toggle.input", [1, 0, 0, 1]) == (0, 1, 0b1001)
    # toggle_test_model.input (config 5): 0->X, 0->Y, X->0, Y->0 ; propensities read Y, X, X, Y
    assert kernel_of("toggle_test.input", [1, 0, 0, 1]) == (0, 2, 0b1001)
    # three species: the generic lattice kernel
    assert kernel_of("repressilator.input", [1, 2, 0, 0, 1, 2])[:2] == (0, 0)
    # a propensity that names a species outside the model is refused
    assert kernel_of("toggle.input", [1, 0, 0, 2])[0] != 0


def test_two_rank_lattice_halo_and_row_formula_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "lattice_host_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "LATTICE HOST OK" in r.stdout
