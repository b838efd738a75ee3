"""Per-rank step time of the weak-scaling bench on ONE GPU: the scenario sets of ranks 0..7 solved one after another,
with the slowest scenarios of each set (SM cycles -> ms at 1.965 GHz)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))
import torch
import bench
from path_planning.solvers.batch import BatchSolver
w = bench.WORKLOADS["c2"]
B, N = w["scenarios"], w["n_agents"]
kw = {}
ranks = range(8)
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "ranks":
        ranks = [int(x) for x in v.split(",")]
    else:
        kw[k] = float(v) if "." in v else int(v)
s = BatchSolver(N, w["time_horizon"], bench.COMMON["time_step"], bench.COMMON["min_distance"], w["space"], **kw)
for rank in ranks:
    p0, pf, _ = bench.make_scenarios(w, rank * B, B)
    d0, d1 = torch.from_numpy(p0).cuda(), torch.from_numpy(pf).cuda()
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = s.solve_device(d0, d1); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    recs = BatchSolver.records_from_bytes(out[3])
    cyc = np.array([r["device_ns"] for r in recs]) / 1e6
    top = np.argsort(-cyc)[:3]
    print(json.dumps(dict(rank=rank, step_ms=round(1e3 * dt, 1), balanced_ms=round(float(cyc.sum()) / 148, 1), max_ms=round(float(cyc.max()), 1),
          over300=int((cyc > 300).sum()), top=[dict(b=int(b), ms=round(float(cyc[b]), 1), scp=recs[b]["scp_iterations"], admm=recs[b]["admm_iterations"],
          infeas=recs[b]["qp_infeasible"], unsolved=recs[b]["qp_unsolved"], pol_att=recs[b]["polish_attempts"], pol_rounds=recs[b]["polish_rounds"], pol_ms=round(recs[b]["cycles_polish"] / 1.965e6, 1), scp_conv=recs[b]["converged"]) for b in top])), flush=True)
