import sys, time, random, numpy as np
sys.path.insert(0, "ba-path-planning_b200")
import torch
from path_planning.solvers.batch import BatchSolver
from path_planning.scenarios.position_generator import generate_positions, generate_positions_large
def run(name, B, N, T, gen, **kw):
    starts, goals = [], []
    space = [0,0,20,20]
    for b in range(B):
        random.seed(10_000 + b)
        if gen == "ref": p0, pf = generate_positions(N, 0.8)
        else: p0, pf, space = generate_positions_large(N, 0.8, time_horizon=T)
        starts.append(p0); goals.append(pf)
    s = BatchSolver(N, T, 0.2, 0.8, space, **kw)
    d0 = torch.from_numpy(np.stack(starts)).cuda(); d1 = torch.from_numpy(np.stack(goals)).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = s.solve_device(d0, d1); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    recs = BatchSolver.records_from_bytes(out[3])
    it = np.array([r["admm_iterations"] for r in recs]); sc = np.array([r["scp_iterations"] for r in recs])
    print(f"{name}: B={B} N={N} K={s.K} time {dt:.3f}s -> {B/dt:.1f} scen/s; scp iters mean {sc.mean():.1f}; admm mean {it.mean():.0f} max {it.max()}; unsolved {sum(r['qp_unsolved'] for r in recs)}; minsep pass {sum(r['min_separation']>=0.79 for r in recs)}/{B}; polish ok {sum(r['polish_ok'] for r in recs)}/{sum(r['polish_attempts'] for r in recs)}; copies max {max(r['max_copies'] for r in recs)}; cyc/it {sum(r['cycles_admm'] for r in recs)/max(1,it.sum()):.0f}; conv {sum(r['converged'] for r in recs)}", flush=True)
run("C5-50", 296, 50, 10.0, "ref")
run("C5-100", 148, 100, 10.0, "large")
run("C3-200", 1, 200, 20.0, "large")
run("C3-200x8", 8, 200, 20.0, "large")
