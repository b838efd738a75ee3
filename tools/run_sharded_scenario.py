"""One large scenario, agents sharded over the GPUs of the node (BASELINE.json config 4), or on one GPU.

  python tools/run_sharded_scenario.py --agents 1000                       # one GPU
  torchrun --nproc-per-node 8 tools/run_sharded_scenario.py --agents 1000  # agent-sharded, NCCL all-gather per iteration

Prints ONE JSON line (rank 0): device time of the whole SCP solve (max over ranks), SCP / ADMM iteration counts,
feasibility outcome.  --check also solves the scenario unsharded on every rank and compares.
"""
import argparse
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from path_planning.scenarios.position_generator import generate_positions_large  # noqa: E402
from path_planning.solvers.stream import StreamSolver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=1000)
    ap.add_argument("--horizon", type=float, default=20.0)
    ap.add_argument("--seed", type=int, default=10_000)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--peer", action="store_true", help="peer-memory exchange of the positions instead of NCCL per iteration")
    ap.add_argument("--set", action="append", default=[], help="solver setting name=value")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    settings = {}
    for kv in a.set:
        k, v = kv.split("=")
        settings[k] = float(v) if "." in v or "e" in v else int(v)
    random.seed(a.seed)
    N, T, h, R = a.agents, a.horizon, 0.2, 0.8
    p0, pf, space = generate_positions_large(N, R, time_horizon=T)
    d0, d1 = torch.from_numpy(p0[None]).cuda(), torch.from_numpy(pf[None]).cuda()
    s = StreamSolver(N, T, h, R, space, n_scenarios=1, sharded=world > 1, peer_exchange=a.peer, **settings)
    best = None
    for _ in range(a.repeats):
        acc, pos, vel, rec = s.solve_device(d0, d1)
        ms = torch.tensor([s.last_device_ms], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = float(ms) if best is None else min(best, float(ms))
    r = StreamSolver.records_from_bytes(rec)[0]
    ok = True
    if a.check:
        ref = StreamSolver(N, T, h, R, space, n_scenarios=1, **settings)
        acc1, pos1, vel1, rec1 = ref.solve_device(d0, d1)
        r1 = StreamSolver.records_from_bytes(rec1)[0]
        err = float((pos - pos1).abs().max())
        ok = err <= 1e-9 and r1["scp_iterations"] == r["scp_iterations"] and r1["admm_iterations"] == r["admm_iterations"]
        if rank == 0:
            print(f"sharded vs one GPU: max |dpos| = {err:.2e}, scp {r['scp_iterations']}/{r1['scp_iterations']}, "
                  f"admm {r['admm_iterations']}/{r1['admm_iterations']}", flush=True)
    if rank == 0:
        print(json.dumps({
            "case": "single scenario, streaming solver", "N": N, "K": s.K, "n_gpus": world, "exchange": s.exchange, "device_ms": best,
            "scp_iterations": r["scp_iterations"], "converged": r["converged"], "admm_iterations": r["admm_iterations"],
            "macro_steps": s.last_macro_steps, "qp_unsolved": r["qp_unsolved"], "rebuilds": r["rebuilds"],
            "max_copies": r["max_copies"], "min_separation": r["min_separation"], "objective": r["objective"],
            "status": r["status"], "candidate_overflow": bool(r.get("reserved2", 0) & 4),
            "us_per_admm_iteration": 1e3 * best / max(1, r["admm_iterations"])}), flush=True)
        if a.check and ok:
            print("SHARDED_CHECK_OK", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
