"""Full-horizon runs of BASELINE configs 1-4 (and the other two example drivers) on the CPU oracle (canonical arithmetic): too large to
commit as arrays, so tests/golden/full_digests.json keeps sizes, counters and SHA-256 digests of the
state list, the decision trace and the probability vector.  The GPU test (tests/test_gpu_full_configs.py)
must reproduce the digests, i.e. be bit-identical at full scale.
Usage: python tests/golden/make_full_digests.py [tag ...]   (goutsias_full takes ~1 h of CPU)
Tags are independent: to use several cores run one process per tag with KFSP_DIGEST_OUT=<part file> each and merge
the part files with `python tests/golden/make_full_digests.py --merge part1.json part2.json ...`."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from gpu_common_cases import CASES  # noqa: E402

FULL_RUNS = {
    # tag: (case, t, FSPTOL, KRYTOL)   test/TestSolverFromFile.f90:12,35 ; SURVEY 8d ; examples/transcr6d.f90:16,50
    "toggle_full": ("toggle", 1000.0, 1e-4, 1e-10),
    "repressilator_full": ("repressilator", 10.0, 1e-4, 1e-10),
    "goutsias_full": ("goutsias", 300.0, 1e-6, 1e-8),
}
# the reference's example programs with their hard-coded CUSTOMPROP (config 4 = transcr6d): tag -> driver name
DRIVER_RUNS = {"driver_toggle_full": "toggle", "driver_repressilator_full": "repressilator", "transcr6d_full": "transcr6d"}
STAT_KEYS = ("nmult", "nexph", "nscale", "nstep", "nreject", "ibrkflag", "mbrkdwn", "n_expand", "n_drop")


def digest(out_states, out_vector, trace_i, trace_d, stats):
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return {"n": int(len(out_vector)), "states_sha256": sha(out_states.astype(np.int32)), "vector_sha256": sha(out_vector.astype(np.float64)),
            "trace_i_sha256": sha(trace_i.astype(np.int32)), "t_step_sha256": sha(trace_d[:, 1].astype(np.float64)),
            "vector_sum": float(out_vector.sum()), "vector_head": [float(v) for v in out_vector[:8]],
            "rows": int(len(trace_i)), "stats": {k: int(stats[k]) for k in STAT_KEYS}}


def main():
    path = os.path.join(HERE, "full_digests.json")
    if len(sys.argv) > 1 and sys.argv[1] == "--merge":
        db = json.load(open(path)) if os.path.exists(path) else {}
        for part in sys.argv[2:]:
            db.update(json.load(open(part)))
        json.dump(db, open(path, "w"), indent=1, sort_keys=True)
        print("merged", len(sys.argv) - 2, "part files ->", sorted(db))
        return
    path = os.environ.get("KFSP_DIGEST_OUT", path)
    db = json.load(open(path)) if os.path.exists(path) else {}
    tags = sys.argv[1:] or list(FULL_RUNS) + list(DRIVER_RUNS)
    for tag in tags:
        if tag in DRIVER_RUNS:
            from krylovfspssa_b200.examples import DRIVERS
            dr = DRIVERS[DRIVER_RUNS[tag]]
            m = oracle.Model(dr["S"], dr["R"], dr["P"], dr["stoich"], dr["params"])
            m.set_custom(dr["oracle_kind"])
            x0, t, ftol, ktol = dr["x0"], dr["t"], dr["fsp_tol"], dr["exp_tol"]
        else:
            name, t, ftol, ktol = FULL_RUNS[tag]
            fname, params, x0 = CASES[name]
            m = oracle.Model.load(os.path.join(ROOT, "krylovfspssa_b200", "models", fname), params)
        t0 = time.time()
        out = oracle.solve(m, [x0], [1.0], t, ftol, ktol, seed=12345, reproducible=1)
        assert out["iflag"] == 0
        d = digest(out["states"], out["vector"], out["trace_i"], out["trace_d"], out["stats"])
        d["oracle_wall_s"] = time.time() - t0
        db[tag] = d
        json.dump(db, open(path, "w"), indent=1, sort_keys=True)
        print(tag, d["n"], d["rows"], "%.1f s" % d["oracle_wall_s"], flush=True)


if __name__ == "__main__":
    main()
