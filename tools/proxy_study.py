"""Does a cheap difficulty proxy (conflicts of the straight-line minimum-acceleration guess) predict per-scenario solve time?
Saves per-scenario SM cycles and proxies for the scenario sets of ranks 0..7 (one GPU)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))
import torch
import bench
from path_planning.solvers.batch import BatchSolver
w = bench.WORKLOAD
B, N, K, R = w["scenarios_per_gpu"], w["n_agents"], 50, w["min_distance"]
s = BatchSolver(N, w["time_horizon"], w["time_step"], R, w["space"])
tau = np.arange(K) / K
sm = 3 * tau ** 2 - 2 * tau ** 3
out = {}
for rank in range(8):
    p0, pf = bench.make_scenarios(rank * B, B, N, R)
    res = s.solve_device(torch.from_numpy(p0).cuda(), torch.from_numpy(pf).cuda())
    torch.cuda.synchronize()
    recs = BatchSolver.records_from_bytes(res[3])
    cyc = np.array([r["cycles_total"] for r in recs], dtype=np.float64)
    scp = np.array([r["scp_iterations"] for r in recs])
    P = p0[:, :, None, :] + (pf - p0)[:, :, None, :] * sm[None, None, :, None]          # (B,N,K,2)
    d = np.linalg.norm(P[:, :, None] - P[:, None, :], axis=-1)                           # (B,N,N,K)
    iu = np.triu_indices(N, 1)
    dd = d[:, iu[0], iu[1], :]                                                           # (B,P,K)
    out[f"cyc{rank}"] = cyc; out[f"scp{rank}"] = scp
    out[f"cnt{rank}"] = (dd < R).sum(axis=(1, 2)); out[f"pen{rank}"] = np.clip(R - dd, 0, None).sum(axis=(1, 2))
    out[f"pairs{rank}"] = (dd.min(axis=2) < R).sum(axis=1); out[f"near{rank}"] = (dd < R + 0.5).sum(axis=(1, 2))
np.savez(os.path.join(ROOT, "gpurun_out", "proxy_study.npz"), **out)
print("saved")
