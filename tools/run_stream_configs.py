"""Streaming solver on the larger BASELINE.json configurations (one GPU): config 3 (one 200-agent scenario),
config 5 style batches (B x {50,100,200} agents, K=50) and, for comparison, the one-CTA-per-scenario solver.
One JSON line per case; `--cases` picks them (name:B:N:T:generator)."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))

import torch  # noqa: E402

from path_planning.scenarios.position_generator import generate_positions, generate_positions_large  # noqa: E402
from path_planning.solvers.batch import BatchSolver  # noqa: E402
from path_planning.solvers.stream import StreamSolver  # noqa: E402

DEFAULT = ["C2-25:256:25:10:ref", "C5-50:256:50:10:large", "C5-100:256:100:10:large", "C5-200:128:200:10:large",
           "C3-200:1:200:20:large"]


def scenarios(B, N, T, gen):
    starts, goals, space = [], [], [0, 0, 20, 20]
    for b in range(B):
        random.seed(10_000 + b)
        if gen == "ref":
            p0, pf = generate_positions(N, 0.8)
        else:
            p0, pf, space = generate_positions_large(N, 0.8, time_horizon=T)
        starts.append(p0)
        goals.append(pf)
    return np.stack(starts), np.stack(goals), space


def summarise(recs):
    it = np.array([r["admm_iterations"] for r in recs])
    return dict(scp_iterations_mean=float(np.mean([r["scp_iterations"] for r in recs])), admm_iterations_mean=float(it.mean()),
                admm_iterations_max=int(it.max()), qp_unsolved=int(sum(r["qp_unsolved"] for r in recs)),
                converged=int(sum(r["converged"] or r["initial_feasible"] for r in recs)),
                minsep_pass=int(sum(r["min_separation"] >= 0.79 for r in recs)), max_copies=int(max(r["max_copies"] for r in recs)),
                overflow=int(sum(bool(r.get("reserved2", 0) & 4) for r in recs)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="*", default=DEFAULT)
    ap.add_argument("--compare-cta", action="store_true", help="also run the one-CTA-per-scenario solver")
    ap.add_argument("--set", action="append", default=[])
    a = ap.parse_args()
    settings = {}
    for kv in a.set:
        k, v = kv.split("=")
        settings[k] = float(v) if "." in v or "e" in v else int(v)
    torch.cuda.set_device(0)
    for case in a.cases:
        name, B, N, T, gen = case.split(":")
        B, N, T = int(B), int(N), float(T)
        p0, pf, space = scenarios(B, N, T, gen)
        d0, d1 = torch.from_numpy(p0).cuda(), torch.from_numpy(pf).cuda()
        s = StreamSolver(N, T, 0.2, 0.8, space, n_scenarios=B, **settings)
        best = None
        for _ in range(2):
            out = s.solve_device(d0, d1)
            best = s.last_device_ms if best is None else min(best, s.last_device_ms)
        recs = StreamSolver.records_from_bytes(out[3])
        line = dict(case=name, solver="stream", B=B, N=N, K=s.K, device_ms=best, scenarios_per_s=1e3 * B / best,
                    macro_steps=s.last_macro_steps, **summarise(recs))
        print(json.dumps(line), flush=True)
        pos_s = out[1].cpu().numpy()
        s.close()
        if a.compare_cta:
            c = BatchSolver(N, T, 0.2, 0.8, space)
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                oc = c.solve_device(d0, d1)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            rc = BatchSolver.records_from_bytes(oc[3])
            pos_c = oc[1].cpu().numpy()
            both = [b for b in range(B) if recs[b]["qp_unsolved"] == 0 and rc[b]["qp_unsolved"] == 0
                    and recs[b]["scp_iterations"] == rc[b]["scp_iterations"]]
            err = max((float(np.linalg.norm(pos_s[b] - pos_c[b]) / np.linalg.norm(pos_c[b])) for b in both), default=None)
            print(json.dumps(dict(case=name, solver="cta", B=B, N=N, K=c.K, wall_ms=1e3 * dt, scenarios_per_s=B / dt,
                                  compared=len(both), max_rel_pos_diff_vs_stream=err, **summarise(rc))), flush=True)


if __name__ == "__main__":
    main()
