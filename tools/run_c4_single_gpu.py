"""Config 4's scenario size (1000 agents, K=100, bounded-travel generator) solved on ONE GPU with the whole-grid kernel."""
import sys, time, random, numpy as np
sys.path.insert(0, "ba-path-planning_b200")
import torch
from path_planning.solvers.batch import BatchSolver
from path_planning.scenarios.position_generator import generate_positions_large
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
random.seed(10_000)
p0, pf, space = generate_positions_large(N, 0.8, time_horizon=20.0)
s = BatchSolver(N, 20.0, 0.2, 0.8, space)
d0 = torch.from_numpy(p0[None]).cuda(); d1 = torch.from_numpy(pf[None]).cuda()
torch.cuda.synchronize(); t0 = time.perf_counter()
out = s.solve_device(d0, d1); torch.cuda.synchronize(); dt = time.perf_counter() - t0
r = BatchSolver.records_from_bytes(out[3])[0]
print(f"N={N} K={s.K} time {dt:.2f}s", {k: r[k] for k in ("status", "scp_iterations", "converged", "admm_iterations", "qp_unsolved", "qp_infeasible", "polish_ok", "polish_attempts", "min_separation", "max_copies", "rebuilds")})
