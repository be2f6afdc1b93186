"""The C-ABI library loads without a GPU and exports every symbol include/kfsp.h declares;
host-side entry points (model reader, propensity compiler) work on the CPU; device entry points
fail loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import krylovfspssa_b200 as k
import oracle
from krylovfspssa_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "kfsp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kfsp_[a-z0-9_]+)\s*\(", txt)) - {"kfsp_propensity_fn"})


def test_every_declared_symbol_is_exported_and_bound():
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(L, n), "libkfsp.so does not export " + n
        assert n in _lib.SIGNATURES, "ctypes binding misses " + n
    assert set(_lib.SIGNATURES) <= set(names)


def test_fortran_shim_binds_only_exported_symbols():
    """Every BIND(C, NAME='...') of the reference-side binding names a symbol the header declares and the library exports,
    and its derived types list the fields of the C structs in order (the shim cannot be compiled here: no Fortran compiler)."""
    src = open(os.path.join(ROOT, "krylovfspssa_b200", "fortran", "kfsp_c_binding.f90")).read()
    bound = sorted(set(re.findall(r"NAME\s*=\s*'(kfsp_[a-z0-9_]+)'", src)))
    assert len(bound) >= 8 and "kfsp_solve" in bound and "kfsp_model_set_custom_propensity" in bound
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    for n in bound:
        assert n in names and hasattr(L, n), n

    def fortran_fields(type_name):
        body = re.search(r"TYPE, BIND\(C\) :: %s(.*?)END TYPE" % type_name, src, flags=re.S).group(1)
        out = []
        for line in body.splitlines():
            if "::" in line:
                out += [f.strip().lower() for f in line.split("::")[1].split(",")]
        return out
    assert fortran_fields("KFSP_OPTIONS") == [n for n, _ in _lib.Options._fields_]
    assert fortran_fields("KFSP_STATS") == [n for n, _ in _lib.Stats._fields_]


def test_struct_layouts_match_the_header():
    # sizes the Fortran shim (krylovfspssa_b200/fortran/kfsp_c_binding.f90) relies on
    assert C.sizeof(_lib.Options) == 12 * 4 + 8 + 6 * 8 + 8
    assert C.sizeof(_lib.Stats) == 8 * 4 + 8 * 8 + 5 * 8 + 3 * 8 + 8
    assert C.sizeof(_lib.TraceRow) == 6 * 8 + 6 * 4
    o = k.default_options()
    assert (o.m_max, o.m_min, o.ideg, o.n_init_onestep, o.fsp_reject_limit, o.max_molecules) == (100, 10, 6, 5, 5, 10000)
    assert o.max_states == 6291469 and (o.delta, o.gamma, o.break_tol, o.drop_tol0, o.drop_fraction) == (1.2, 0.9, 1e-7, 1e-8, 0.1)


@pytest.mark.parametrize("name", ["toggle.input", "toggle_test.input", "repressilator.input", "goutsias.input", "birth_death.input"])
def test_model_reader_matches_oracle_reader(name):
    path = os.path.join(k.models_dir(), name)
    m = k.CME_MODEL().load(path)
    im = oracle.model_input.load(path)
    assert (m.nspecies, m.nreactions, m.nparameters) == (im.nspecies, im.nreactions, im.nparameters)
    assert m.species_names == im.species_names and m.parameter_names == im.parameter_names
    assert np.array_equal(m.stoichiometry.T, np.array(im.stoichiometry))
    for r, pr in enumerate(im.programs):
        code, imm = m.bytecode(r + 1)
        assert code == pr.code and imm == pr.immed          # same byte code as the fparser restatement


def test_propensity_compiler_against_python_restatement():
    rng = np.random.default_rng(0)
    m = k.CME_MODEL().create(2, 1, 2)
    exprs = ["p1*X1", "p1/(1.0+X2^2.5)", "-X1*p2+3.0e0", "p1*X1*(X1-1)/2.0d0", "exp(-p1)+sqrt(X2)", "X1**2-X2**2",
             "p1+p2/(2.0+0.2*X2^2)", "abs(X1-X2)*p1/p2/2", "-(X1+1)^2", "p1-p2-X1+X2", "log10(X1+10)*log(p2+2)"]
    for e in exprs:
        m.set_propensity(1, e)
        pr = oracle.fparser.Program(e, ["X1", "X2", "p1", "p2"])
        code, imm = m.bytecode(1)
        assert code == pr.code and imm == pr.immed, e
        for _ in range(5):
            st = rng.integers(0, 50, 2)
            par = rng.uniform(0.1, 5.0, 2)
            m.reset_parameters(par)
            assert m.propensity(st, 1) == pr.evaluate([float(st[0]), float(st[1]), par[0], par[1]]), e
    with pytest.raises(k.KfspError) as err:
        m.set_propensity(1, "p1*(X1")
    assert err.value.status == -24
    with pytest.raises(k.KfspError):
        m.set_propensity(1, "p1*Z")


def test_prop_grid_known_answer_product_reader():
    # test/TestModelParser.f90:31-45 through the product's own reader/compiler
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle_test.input"))
    m.reset_parameters([5000.0, 1600.0, 1.0, 1.0])
    err = 0.0
    for i in range(1, 51):
        for j in range(1, 51):
            want = [5000.0 / (1.0 + float(j) ** 2.5), 1600.0 / (1.0 + float(i) ** 1.5), 1.0 * i, 1.0 * j]
            for r in range(4):
                err += abs(want[r] - m.propensity([i, j], r + 1))
    assert err == 0.0


def test_bad_model_files(tmp_path):
    with pytest.raises(k.KfspError) as e:
        k.CME_MODEL().load(str(tmp_path / "missing.input"))
    assert e.value.status == -25
    p = tmp_path / "bad.input"
    p.write_text("nspecies\n1\nnreactions\n3\nnparameters\n1\nspecies\nX\nparameters\nk\nreactions\n0 -> X\n")
    with pytest.raises(k.KfspError) as e:
        k.CME_MODEL().load(str(p))
    assert e.value.status == -24


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle.input"))
    with pytest.raises(k.KfspError) as e:
        k.KrylovFspHandle(m)
    assert e.value.status == -20          # KFSP_ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "krylovfspssa_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90", ".sh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "kfsp_oracle" not in txt.replace("oracle/kfsp_oracle.cpp", ""), f


def test_shim_call_sequence_compiles_as_c_and_fails_loudly_without_a_gpu(tmp_path):
    """tests/host/shim_sequence.c: the statements of the Fortran shim's CME_SOLVE (krylovfspssa_b200/fortran/kfsp_c_binding.f90,
    which no compiler here can build) from a compiled C99 host -- include/kfsp.h must be plain C and every entry point the shim
    binds must link.  Without a GPU the program stops at kfsp_create with KFSP_ERR_NO_DEVICE (the GPU run: test_gpu_shim_sequence.py)."""
    import subprocess
    import torch
    from shim_build import build_shim_sequence
    exe = build_shim_sequence(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by test_gpu_shim_sequence.py")
    r = subprocess.run([exe, "bytecode", os.path.join(k.models_dir(), "toggle.input"), "5", "1e-4", "1e-8", "100000", "12345",
                        str(tmp_path / "o.bin"), "0", "0", "1", "100", "1", "1", "100", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CUDA device" in r.stderr
