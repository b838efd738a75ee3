"""Setting sweep for the streaming solver: 64 config-2 scenarios + accuracy on two golden fixtures + one 100-agent scenario."""
import json, os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))
import torch
from path_planning.scenarios.position_generator import generate_positions, generate_positions_large
from path_planning.solvers.stream import StreamSolver

def c2(B):
    st, go = [], []
    for b in range(B):
        random.seed(10_000 + b); p0, pf = generate_positions(25, 0.8); st.append(p0); go.append(pf)
    return np.stack(st), np.stack(go)

combos = [json.loads(a) for a in sys.argv[1:]] or [{}]
S, G = c2(64)
random.seed(10_000); L0, L1, Lspace = generate_positions_large(100, 0.8, time_horizon=20.0)
gold = [np.load(os.path.join(ROOT, "tests/golden", n)) for n in ("n10_s1.npz", "n25_s3.npz")]
for kw in combos:
    s = StreamSolver(25, 10.0, 0.2, 0.8, [0, 0, 20, 20], n_scenarios=64, **kw)
    traj, recs = s.solve(S, G); ms = s.last_device_ms; s.close()
    it = np.array([r["admm_iterations"] for r in recs])
    line = dict(set=kw, c2_ms=round(ms, 1), admm_mean=int(it.mean()), admm_max=int(it.max()), unsolved=sum(r["qp_unsolved"] for r in recs),
                all_solved=sum(r["qp_unsolved"] == 0 for r in recs), minsep_pass=sum(r["min_separation"] >= 0.79 for r in recs),
                scp_mean=round(float(np.mean([r["scp_iterations"] for r in recs])), 2))
    errs = []
    for g in gold:
        N = int(g["N"])
        s = StreamSolver(N, float(g["T"]), float(g["h"]), float(g["R"]), list(g["space"]), n_scenarios=1, **kw)
        t, r = s.solve(g["p0"][None], g["pf"][None]); s.close()
        errs.append((float("%.1e" % (np.linalg.norm(t["positions"][0] - g["positions"]) / np.linalg.norm(g["positions"]))),
                     r[0]["admm_iterations"], r[0]["scp_iterations"] == int(g["iterations"]), r[0]["qp_unsolved"]))
    line["golden_err_iters_sameit_unsolved"] = errs
    s = StreamSolver(100, 20.0, 0.2, 0.8, Lspace, n_scenarios=1, **kw)
    t, r = s.solve(L0[None], L1[None]); r = r[0]
    line["n100"] = dict(ms=round(s.last_device_ms, 1), scp=r["scp_iterations"], conv=r["converged"], admm=r["admm_iterations"], unsolved=r["qp_unsolved"],
                        infeas=r["qp_infeasible"], minsep=round(r["min_separation"], 4), copies=r["max_copies"], rebuilds=r["rebuilds"])
    s.close()
    print(json.dumps(line), flush=True)
