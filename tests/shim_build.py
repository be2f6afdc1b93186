"""Builds tests/host/shim_sequence.c (the C twin of the Fortran shim's CME_SOLVE body) against include/kfsp.h and libkfsp.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim_sequence(outdir):
    exe = os.path.join(str(outdir), "shim_sequence")
    libdir = os.path.join(ROOT, "krylovfspssa_b200")
    # strict C99: the header has to be plain C for a cgo / ISO_C_BINDING / ctypes host
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(ROOT, "tests", "host", "shim_sequence.c"), "-L", libdir, "-l:libkfsp.so",
                           "-Wl,-rpath," + libdir, "-lm"])
    return exe
