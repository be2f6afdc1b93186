"""A/B timing of the lattice SpMV variants (KFSP_BOX_TUNE) and the explicit kernel on one GPU:
python tools/lattice_tune.py [bx by]  ->  ms per launch and GB/s on the 16 B/state (lattice) and 72 B/state (explicit) accounting."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(variant, bx, by):
    import numpy as np
    import krylovfspssa_b200 as k
    from krylovfspssa_b200._lib import lib, check
    import bench
    model = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle_test.input"))
    model.reset_parameters(bench.PARAMS)
    n = bx * by
    h = k.KrylovFspHandle(model, max_states=n + 64, spmv_variant=variant, n_init_onestep=0, enable_drop=0, enable_expand=0, m_max=12)
    if variant == 1:
        h.fsp_init_box([bx, by])
    else:
        states, _ = bench.synthetic(bx, by)
        h.fsp_init(states)
    L = lib()
    px, py = C.c_void_p(), C.c_void_p()
    check(L.kfsp_device_alloc(h._h, 8 * n, C.byref(px)))
    check(L.kfsp_device_alloc(h._h, 8 * n, C.byref(py)))
    x = np.random.default_rng(0).random(n)
    check(L.kfsp_device_upload(h._h, px, x.ctypes.data_as(C.c_void_p), 8 * n))
    sec = C.c_double()
    check(L.kfsp_matvec_device(h._h, px, py, 5, C.byref(sec)))
    check(L.kfsp_matvec_device(h._h, px, py, 20, C.byref(sec)))
    by_state = 16 if variant == 1 else 72
    print("variant %d tune %s: %.4f ms/launch  %.0f GB/s on %d B/state  (%.3e states/s)" %
          (variant, os.environ.get("KFSP_BOX_TUNE", "-"), 1e3 * sec.value, by_state * n / sec.value / 1e9, by_state, n / sec.value), flush=True)
    # fused-dot variant inside a sweep: time one Arnoldi sweep of 10 columns
    v = np.abs(np.random.default_rng(1).standard_normal(n))
    H, av, brk, s = h.arnoldi(v, 10)
    H, av, brk, s = h.arnoldi(v, 10)
    print("    arnoldi sweep m=10: %.2f ms" % (1e3 * s), flush=True)
    h.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        bx, by = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 10000)
        for variant, tune in [(1, "0"), (1, "10")]:
            env = dict(os.environ, KFSP_BOX_TUNE=tune)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(variant), str(bx), str(by)], env=env)
