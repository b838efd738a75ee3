import sys, os, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))
import torch
from path_planning.solvers.stream import StreamSolver
from path_planning.solvers.batch import BatchSolver
keys = ["status","scp_iterations","converged","initial_feasible","admm_iterations","qp_unsolved","qp_infeasible","rebuilds","max_copies","min_separation","objective","pri_res","dua_res","rel_steps"]
for name in sys.argv[1:]:
    g = np.load(os.path.join(ROOT, "tests/golden", name))
    N, h, R, space, T = int(g["N"]), float(g["h"]), float(g["R"]), list(g["space"]), float(g["T"])
    print("==", name, "golden iterations", int(g["iterations"]), "rel", list(np.round(g["rel_steps"],5)), "obj", float(g["objective"]))
    for kw in ({}, {"stall_window":0}, {"stall_window":0,"max_admm_iter":20000}, {"adapt_every":0, "stall_window":0}):
        s = StreamSolver(N, T, h, R, space, n_scenarios=1, **kw)
        traj, recs = s.solve(g["p0"][None], g["pf"][None])
        r = recs[0]
        perr = np.linalg.norm(traj["positions"][0]-g["positions"])/np.linalg.norm(g["positions"])
        print(" stream", kw, {k:r[k] for k in keys}, "perr %.2e"%perr, "ms", s.last_device_ms)
        s.close()
    c = BatchSolver(N, T, h, R, space, polish=0)
    traj, recs = c.solve(g["p0"][None], g["pf"][None]); r = recs[0]
    perr = np.linalg.norm(traj["positions"][0]-g["positions"])/np.linalg.norm(g["positions"])
    print(" cta polish=0", {k:r[k] for k in keys}, "perr %.2e"%perr)
