"""ORACLE / TEST INFRASTRUCTURE ONLY -- generates tests/golden/c2_outcomes.npz.

Run in the build container (needs /root/reference):
    python oracle/make_outcomes.py [first_seed] [count] [workers]

What it records: the OUTCOME of the VERBATIM reference (scp.py loaded by
oracle/ref_loader.py, osqp shim at the reference's own settings -- OSQP defaults
eps 1e-3 for QP #0, warm_start=True / max_iter=10000 for every later QP,
scp.py:360, 442) on the benchmark's own scenarios: config 2 (25 agents, T=10,
h=0.2, R=0.8, 20 x 20 m), scenario b generated after random.seed(10_000 + b).
Unlike tests/golden/n*.npz these cases are NOT filtered for "every subproblem
solved": scenarios whose linearised subproblems are primal infeasible or run
into max_iter are exactly what this fixture is for.  The driver follows
compute_trajectories_batch.py:46-55 (try / except -> status "error").

Per scenario: batch status, exception text, SCP iterations, the OSQP status and
iteration count of every QP, rel-step sequence, whether the returned trajectory
is finite, min separation + pass/fail (scp.py:610), dynamics residual + pass/fail
(SURVEY.md 8c, 1e-3), objective, and the trajectory itself.

tests/golden/c2_outcomes_alt.npz is the same run with ONE OSQP setting changed
(SCP_OUTCOME_OVERRIDES='{"adaptive_rho_interval": 50}'; real OSQP 0.6 derives that
interval from its measured setup time, i.e. it differs from run to run): it measures how
far the reference is from ITSELF on scenarios whose subproblems end at max_iter.
"""

from __future__ import annotations

import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = dict(N=25, T=10.0, h=0.2, R=0.8, space=[0, 0, 20, 20], max_iterations=15)
# other sizes, e.g. config 1: SCP_OUTCOME_CFG='{"N": 10, "T": 100.0, "space": [0, 0, 200, 200]}' SCP_OUTCOME_SEEDS=0
CFG.update(json.loads(os.environ.get("SCP_OUTCOME_CFG", "{}")))
MAXQ = 16          # QP #0 + at most 15 avoidance QPs


def run_one(seed):
    from oracle import ref_loader, scp_oracle

    ref = ref_loader.load_reference()
    import osqp

    osqp.OVERRIDES.clear()
    osqp.OVERRIDES.update(json.loads(os.environ.get("SCP_OUTCOME_OVERRIDES", "{}")))   # e.g. a second OSQP setting
    osqp.STATS.clear()
    N, T, h, R, space = CFG["N"], CFG["T"], CFG["h"], CFG["R"], CFG["space"]
    K = int(T / h)
    random.seed(seed)
    np.random.seed(seed)
    p0, pf = ref.scenarios.position_generator.generate_positions(N, R)
    status, err, tr = "success", "", None
    t0 = time.perf_counter()
    with ref_loader.quiet() as buf:
        try:
            s = ref.solvers.scp.SCP(n_vehicles=N, time_horizon=T, time_step=h, min_distance=R, space_dims=space)
            s.set_initial_states(p0)
            s.set_final_states(pf)
            tr = s.generate_trajectories(max_iterations=CFG["max_iterations"])
        except Exception as e:  # compute_trajectories_batch.py:50-54
            status, err = "error", f"{type(e).__name__}: {e}"
    wall = time.perf_counter() - t0
    rels = [float(x) for x in buf.getvalue().splitlines() if x and (x[0].isdigit() or x.startswith("nan"))]
    stats = list(osqp.STATS)
    out = dict(seed=seed, status=status, error=err, time_sec=wall, p0=np.asarray(p0), pf=np.asarray(pf),
               qp_status=np.full(MAXQ, 0, np.int32), qp_iter=np.full(MAXQ, 0, np.int32), n_qp=len(stats),
               rel_steps=np.full(MAXQ - 1, np.nan), scp_iterations=len(rels),
               positions=np.full((N, K, 2), np.nan), accelerations=np.full((N, K, 2), np.nan),
               finite=False, min_separation=np.nan, minsep_pass=False, dyn_residual=np.nan, dyn_pass=False,
               objective=np.nan)
    for i, st in enumerate(stats[:MAXQ]):
        out["qp_status"][i] = st["status"]
        out["qp_iter"][i] = st["iter"]
    out["rel_steps"][: len(rels)] = rels[: MAXQ - 1]
    if tr is not None:
        pos, acc = np.asarray(tr["positions"], float), np.asarray(tr["accelerations"], float)
        out["positions"], out["accelerations"] = pos, acc
        out["finite"] = bool(np.isfinite(pos).all() and np.isfinite(acc).all())
        if out["finite"]:
            z = np.zeros((N, 2))
            out["min_separation"] = scp_oracle.min_separation(pos)
            out["dyn_residual"] = scp_oracle.dynamics_residual(acc, p0, z, pf, z, h, space, positions=pos)
            out["objective"] = float((acc ** 2).sum())
            out["minsep_pass"] = bool(out["min_separation"] >= R - 0.01)
            out["dyn_pass"] = bool(out["dyn_residual"] <= 1e-3)
    return out


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "tests", "golden", "c2_outcomes.npz")
    import multiprocessing as mp

    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    seeds = list(range(first, first + count))
    if os.environ.get("SCP_OUTCOME_SEEDS"):
        seeds = [int(x) for x in os.environ["SCP_OUTCOME_SEEDS"].split(",")]
    t0 = time.time()
    res = []
    with mp.get_context("spawn").Pool(workers) as pool:
        for r in pool.imap(run_one, seeds):
            res.append(r)
            print(f"seed {r['seed']}: {r['status']} finite={r['finite']} scp={r['scp_iterations']} "
                  f"qp_status={list(r['qp_status'][:r['n_qp']])} minsep={r['min_separation']:.4f} "
                  f"{r['time_sec']:.0f}s  [{time.time()-t0:.0f}s]", flush=True)
    keys = [k for k in res[0] if k not in ("error", "status")]
    np.savez_compressed(path, config=np.array([CFG["N"], CFG["T"], CFG["h"], CFG["R"], *CFG["space"], CFG["max_iterations"]], float),
                        status=np.array([r["status"] for r in res]), error=np.array([r["error"] for r in res]),
                        **{k: np.array([r[k] for r in res]) for k in keys})
    print("wrote", path)


if __name__ == "__main__":
    main()
