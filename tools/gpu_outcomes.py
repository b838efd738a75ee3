#!/usr/bin/env python
"""Solves config-2 scenarios (seeds 10000+b) through the C-ABI host call and dumps the per-scenario
records (+ trajectories of the first --dump scenarios) to an .npz for offline comparison with
tests/golden/c2_outcomes.npz (the reference's outcomes on the same seeds).

    python tools/gpu_outcomes.py --count 1024 --dump 64 --out gpurun_out/gpu_outcomes_default.npz [--set k=v ...]
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ba-path-planning_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=1024)
    ap.add_argument("--dump", type=int, default=64)
    ap.add_argument("--agents", type=int, default=25)
    ap.add_argument("--seed-base", type=int, default=10_000)
    ap.add_argument("--out", default="gpurun_out/gpu_outcomes.npz")
    ap.add_argument("--set", action="append", default=[])
    a = ap.parse_args()
    from path_planning import _capi
    from path_planning.scenarios.position_generator import generate_positions

    lib = _capi.load()
    N, T, h, R, space = a.agents, 10.0, 0.2, 0.8, [0, 0, 20, 20]
    K = int(T / h)
    B = a.count
    p0 = np.empty((B, N, 2)); pf = np.empty((B, N, 2))
    for b in range(B):
        random.seed(a.seed_base + b)
        p0[b], pf[b] = generate_positions(N, R)
    prob = _capi.default_problem(N, T, h, R, space)
    for kv in a.set:
        k, v = kv.split("=")
        setattr(prob, k, float(v) if ("." in v or "e" in v) else int(v))
    z = np.zeros_like(p0)
    acc = np.empty((B, N, K, 2)); pos = np.empty((B, N, K, 2)); vel = np.empty((B, N, K, 2))
    rec = (_capi.Record * B)()
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    for rep in range(2):
        t0 = time.perf_counter()
        _capi.check(lib.scp_b200_solve_batch_host(C.byref(prob), B, ptr(p0), ptr(z), ptr(pf), ptr(z), ptr(acc), ptr(pos),
                                                  ptr(vel), C.cast(rec, C.c_void_p), 0))
        wall = time.perf_counter() - t0
    recs = [_capi.record_to_dict(r) for r in rec]
    keys = ["status", "scp_iterations", "converged", "initial_feasible", "admm_iterations", "qp_unsolved", "qp_infeasible",
            "polish_ok", "polish_attempts", "min_separation", "objective", "cycles_total", "cycles_admm", "cycles_polish",
            "reserved2", "rebuilds", "max_copies", "device_ns", "polish_rounds", "cycles_pbuild", "cycles_psolve", "cycles_peval",
            "cycles_papply"]
    out = {k: np.array([r[k] for r in recs]) for k in keys}
    rel = np.full((B, 32), np.nan)
    for b, r in enumerate(recs):
        rel[b, : len(r["rel_steps"])] = r["rel_steps"]
    d = min(a.dump, B)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    np.savez_compressed(a.out, rel_steps=rel, positions=pos[:d], accelerations=acc[:d], p0=p0[:d], pf=pf[:d], wall=wall,
                        settings=np.array(a.set), **out)
    uns = out["qp_unsolved"] > 0
    print(f"{a.set}: {B} scenarios in {wall*1e3:.0f} ms ({B/wall:.0f}/s); unsolved-subproblem scenarios {int(uns.sum())}, "
          f"infeasible-flagged {int((out['qp_infeasible']>0).sum())}, status!=0 {int((out['status']!=0).sum())}, "
          f"minsep fail {int((out['min_separation'] < R-0.01).sum())}, admm/scen {out['admm_iterations'].mean():.0f}, "
          f"max cycles {out['cycles_total'].max()/1.965e6:.0f} ms, sum cycles/148 {out['cycles_total'].sum()/1.965e6/148:.0f} ms")


if __name__ == "__main__":
    main()
