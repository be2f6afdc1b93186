"""Developer tool: run the synthetic config-5 solve for several t_final and print the decision trace."""
import sys, os, time, numpy as np, ctypes as C
sys.path.insert(0, os.getcwd())
import bench, krylovfspssa_b200 as k
from krylovfspssa_b200._lib import Stats, check, lib
L = lib()
bx = by = int(sys.argv[1])
states, p0 = bench.synthetic(bx, by)
n = len(p0)
model = k.CME_MODEL().load(os.path.join(k.models_dir(), "toggle_test.input")); model.reset_parameters(bench.PARAMS)
for mmax in (30, 60):
  h = k.KrylovFspHandle(model, max_states=n + 64, m_max=mmax, m_min=10, n_init_onestep=0, enable_drop=0, enable_expand=0)
  h.fsp_init(states)
  for tf in [float(x) for x in sys.argv[2:]]:
    h.set_vector(p0)
    t0 = time.time(); rc, st = h.solve_resident(tf, 1e-6, 1e-8); dt = time.time() - t0
    tr = h.trace()
    print("m_max", mmax, "t_final", tf, "rc", rc, "steps", st["nstep"], "nmult", st["nmult"], "nexph", st["nexph"], "nreject", st["nreject"], "dev_s %.4f" % st["device_seconds"], "m[:12]", tr["i"][:12, 0].tolist(), "t_step[:12]", [float("%.3g" % v) for v in tr["d"][:12, 1]], "wsum", tr["d"][-1, 3], flush=True)
  h.close()
