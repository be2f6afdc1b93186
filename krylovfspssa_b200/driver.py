"""File-driven driver with a result writer: the generic form of the reference's `test/TestSolverFromFile.f90`
(which hard-codes the model file, parameter values, start state, horizon and tolerances, `:12-35`, and prints only the
elapsed time; the solver statistics of `KrylovSolver.f90:554-573` are discarded by the reference).

    python -m krylovfspssa_b200.driver models/toggle.input --params 1,100,1,1,100,1 --x0 0,0 --t 1000 \
        --fsptol 1e-4 --krytol 1e-10 --out toggle_t1000.npz

The result file holds the final state list (index order), the probability vector, the marginal distribution of every
species, the decision trace (one row per pass of the time-step loop) and the statistics, so two runs -- or a run and
the CPU oracle -- can be compared array by array."""
import argparse
import json
import sys
import time

import numpy as np


def marginals(states, p):
    """Marginal distribution of every species: list of arrays m_s[c] = sum of p over states with count c of species s."""
    states = np.asarray(states)
    p = np.asarray(p, dtype=np.float64)
    out = []
    for s in range(states.shape[1]):
        m = np.zeros(int(states[:, s].max()) + 1 if len(p) else 0)
        np.add.at(m, states[:, s], p)
        out.append(m)
    return out


def write_result(path, out, species_names, meta):
    """states, vector, marginals, trace and stats of one solve -> compressed .npz (json for the scalars)."""
    marg = marginals(out["states"], out["vector"])
    arrays = {"states": out["states"].astype(np.int32), "vector": out["vector"].astype(np.float64),
              "trace_d": out["trace"]["d"], "trace_i": out["trace"]["i"]}
    for name, m in zip(species_names, marg):
        arrays["marginal_" + name] = m
    stats = {k: (float(v) if isinstance(v, float) else int(v)) for k, v in out["stats"].items()}
    arrays["stats_json"] = np.array(json.dumps(stats))
    arrays["meta_json"] = np.array(json.dumps(dict(meta, species=list(species_names), iflag=int(out["iflag"]))))
    np.savez_compressed(path, **arrays)
    return marg


def read_result(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta_json"]))
    return dict(states=z["states"], vector=z["vector"], trace=dict(d=z["trace_d"], i=z["trace_i"]),
                stats=json.loads(str(z["stats_json"])), meta=meta,
                marginals={n: z["marginal_" + n] for n in meta["species"]})


def _floats(text):
    return [float(v) for v in text.replace(";", ",").split(",") if v.strip()]


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="krylovfspssa_b200.driver", description=__doc__.split("\n\n")[0])
    ap.add_argument("model", help="`.input` model file (the reference's format, src/model/ModelModule.f90:59-161)")
    ap.add_argument("--params", required=True, type=_floats, help="parameter values in file order (RESET_PARAMETERS)")
    ap.add_argument("--x0", required=True, type=lambda t: [int(v) for v in _floats(t)], help="initial state")
    ap.add_argument("--t", required=True, type=float, help="final time")
    ap.add_argument("--fsptol", type=float, default=1e-4)
    ap.add_argument("--krytol", type=float, default=1e-8)
    ap.add_argument("--max-states", type=int, default=6291469)
    ap.add_argument("--seed", type=int, default=12345, help="seed of the per-trajectory SSA streams")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--verbosity", type=int, default=1)
    ap.add_argument("--out", default=None, help="result file (.npz)")
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(sys.argv[1:] if argv is None else argv)
    import krylovfspssa_b200 as k
    model = k.CME_MODEL().load(args.model)
    if len(args.params) != model.nparameters or len(args.x0) != model.nspecies:
        raise SystemExit("the model has %d parameters and %d species" % (model.nparameters, model.nspecies))
    model.reset_parameters(args.params)
    h = k.KrylovFspHandle(model, max_states=args.max_states, seed=args.seed, device=args.device)
    t0 = time.time()
    out = h.solve(args.t, [args.x0], [1.0], args.fsptol, args.krytol, verbosity=args.verbosity)
    wall = time.time() - t0
    st = out["stats"]
    names = model.species_names
    print("Elapsed time :%10.2f" % wall)                                       # test/TestSolverFromFile.f90:37-38
    print("states %d  steps %d  SpMVs %d  expansions %d  drops %d  mass %.12f  iflag %d" %
          (len(out["vector"]), st["nstep"], st["nmult"], st["n_expand"], st["n_drop"], out["vector"].sum(), out["iflag"]))
    marg = marginals(out["states"], out["vector"])
    for name, m in zip(names, marg):
        mean = float((np.arange(len(m)) * m).sum())
        print("  %-12s mean %.6g  support 0..%d" % (name, mean, len(m) - 1))
    if args.out:
        write_result(args.out, out, names, dict(model=args.model, params=args.params, x0=args.x0, t=args.t, fsptol=args.fsptol,
                                                  krytol=args.krytol, seed=args.seed, wall_seconds=wall))
        print("wrote", args.out)
    h.close()
    return out


if __name__ == "__main__":
    main()
