"""Profiling aid: per-scenario solve cycles of config 2 and a list-scheduling simulation of the persistent
kernel's tail for different scenario orders (index order, longest-first, a cheap conflict-count predictor)."""
import heapq, json, random, sys
import numpy as np
sys.path.insert(0, "ba-path-planning_b200")
import torch
from path_planning.solvers.batch import BatchSolver
from path_planning.scenarios.position_generator import generate_positions

B, N = 1024, 25
starts, goals = [], []
for b in range(B):
    random.seed(10_000 + b); p0, pf = generate_positions(N, 0.8); starts.append(p0); goals.append(pf)
starts, goals = np.stack(starts), np.stack(goals)
s = BatchSolver(N, 10.0, 0.2, 0.8, [0, 0, 20, 20])
out = s.solve_device(torch.from_numpy(starts).cuda(), torch.from_numpy(goals).cuda()); torch.cuda.synchronize()
recs = BatchSolver.records_from_bytes(out[3])
cyc = np.array([r["cycles_total"] for r in recs], dtype=float)
# predictor: pairs whose straight constant-speed paths come within 1.5 R
t = np.linspace(0, 1, 51)[None, None, :, None]
path = starts[:, :, None, :] * (1 - t) + goals[:, :, None, :] * t          # (B,N,51,2)
d = np.linalg.norm(path[:, :, None] - path[:, None, :], axis=-1).min(axis=-1)  # (B,N,N)
pred = ((d < 1.2).sum(axis=(1, 2)) - N) / 2
def simulate(order, workers=148):
    h = [0.0] * workers; heapq.heapify(h)
    for b in order:
        t0 = heapq.heappop(h); heapq.heappush(h, t0 + cyc[b])
    return max(h)
idx = np.arange(B)
print("sum/148 (perfect)", cyc.sum() / 148 / 1.965e9, "s; max scenario", cyc.max() / 1.965e9)
print("index order     ", simulate(idx) / 1.965e9)
print("longest first   ", simulate(np.argsort(-cyc)) / 1.965e9)
print("predictor first ", simulate(np.argsort(-pred, kind="stable")) / 1.965e9, "corr", np.corrcoef(pred, cyc)[0, 1])
print("quantiles of cycles (s):", np.quantile(cyc, [0.1, 0.5, 0.9, 0.99, 1.0]) / 1.965e9)
