"""``compute-trajectories-batch``: timing benchmark over several N (reference
src/path_planning/cli/compute_trajectories_batch.py:14-173).  Same CONFIG keys, same
``run_single_trial`` record, same JSON/CSV schema (schema_version "1.0", CSV columns
N,trial_index,status,time_sec,K,T,h,error); fields are only added.

New: with ``CONFIG["batched"]`` (default True) all trials of one N are solved by ONE device
launch (scenario-level data parallelism).  ``time_sec`` keeps the reference's meaning -- the time
of THAT trial's solve (compute_trajectories_batch.py:46-55): it is the GPU time the scenario
itself occupied (``device_ns`` of its result record) plus its share of the host<->device copies,
so trials of one N keep their spread for the box-plot consumer; ``batch_time_sec`` (added) is the
wall time of the whole launch.  Launched under ``torch.distributed.run`` (WORLD_SIZE > 1) the
trials of every N are sharded over the ranks, one GPU each, with no data-path collective
(solvers/sharding.py); rank 0 gathers the records and writes the files.  ``CONFIG["rng_seed"]``
now seeds the stdlib ``random`` module that the scenario generator really uses (the reference's
np.random.seed does not reach it -- TODOs at compute_trajectories_batch.py:40,65); the seed is
stored per run.  A YAML file with the same keys (``compute-trajectories-batch config.yaml`` or
the SCP_BATCH_CONFIG environment variable; reference TODO at :12 and configs/info.txt) overrides
CONFIG; see configs/batch_default.yaml.
"""

import csv
import json
import os
import random
import sys
import time
from datetime import datetime
from pathlib import Path

import numpy as np

from ..scenarios.position_generator import generate_positions
from ..solvers.scp import SCP

CONFIG = {
    "Ns": [18, 20],
    "trials_per_N": 10,
    "time_horizon": 10.0,
    "time_step": 0.2,
    "min_distance": 0.8,
    "space_dims": [0, 0, 20, 20],
    "max_iterations": 15,
    "rng_seed": None,
    "results_dir": "data/trial_xxx",
    "batched": True,
}


def run_single_trial(N, cfg, rng):
    """One SCP solve for N vehicles -> result record (reference :28-67)."""
    solver = SCP(n_vehicles=N, time_horizon=cfg["time_horizon"], time_step=cfg["time_step"],
                 min_distance=cfg["min_distance"], space_dims=cfg["space_dims"])
    init_pos, final_pos = generate_positions(N, cfg["min_distance"])
    solver.set_initial_states(init_pos)
    solver.set_final_states(final_pos)
    t0 = time.perf_counter()
    status, err_msg = "success", None
    try:
        _ = solver.generate_trajectories(max_iterations=cfg["max_iterations"])
    except Exception as e:
        status, err_msg = "error", str(e)
    t1 = time.perf_counter()
    rec = solver.last_record or {}
    return {
        "N": N, "status": status, "time_sec": t1 - t0, "error": err_msg,
        "K": getattr(solver, "K", None), "T": getattr(solver, "T", cfg["time_horizon"]),
        "h": getattr(solver, "h", cfg["time_step"]),
        "scp_iterations": rec.get("scp_iterations"), "admm_iterations": rec.get("admm_iterations"),
        "min_separation": rec.get("min_separation"),
    }


def load_config(path):
    """YAML file -> dict of CONFIG overrides (unknown keys are rejected; reference TODO, :12)."""
    import yaml

    with open(path, "r", encoding="utf-8") as f:
        data = yaml.safe_load(f) or {}
    if not isinstance(data, dict):
        raise ValueError(f"{path}: expected a mapping of CONFIG keys")
    unknown = sorted(set(data) - set(CONFIG) - {"engine", "analysis"})
    if unknown:
        raise ValueError(f"{path}: unknown config keys {unknown}; known: {sorted(CONFIG)}")
    return data


def _dist_context():
    """(rank, world) of a torch.distributed.run launch; initialises the process group (gloo: only result records
    travel) and binds this process to its GPU.  (0, 1) when launched plainly."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
    return dist.get_rank(), dist.get_world_size()


def run_batch(N, cfg, n_trials, seeds=None):
    """All trials of one N in one device launch per rank.  Returns the list of per-trial records (every rank
    receives all of them; trials are sharded over the ranks of a torch.distributed job)."""
    from ..solvers.batch import BatchSolver
    from ..solvers.sharding import gather_records, shard_range

    rank, world = _dist_context()
    starts, goals, used = [], [], []
    for t in range(n_trials):
        if seeds is not None:
            random.seed(seeds[t])
        p0, pf = generate_positions(N, cfg["min_distance"])
        starts.append(p0)
        goals.append(pf)
        used.append(None if seeds is None else seeds[t])
    lo, hi = shard_range(n_trials, rank, world)
    if hi <= lo:
        return gather_records([])
    n_trials = hi - lo
    starts, goals, used = starts[lo:hi], goals[lo:hi], used[lo:hi]
    from ..solvers.scp import use_stream_engine

    K = int(cfg["time_horizon"] / cfg["time_step"])
    if use_stream_engine(cfg.get("engine", "auto"), N, K, n_trials):
        from ..solvers.stream import StreamSolver

        solver = StreamSolver(N, cfg["time_horizon"], cfg["time_step"], cfg["min_distance"], cfg["space_dims"],
                              n_scenarios=n_trials, max_scp_iter=cfg["max_iterations"])
    else:
        solver = BatchSolver(N, cfg["time_horizon"], cfg["time_step"], cfg["min_distance"], cfg["space_dims"],
                             max_scp_iter=cfg["max_iterations"])
    t0 = time.perf_counter()
    traj, recs = solver.solve(np.stack(starts), np.stack(goals))
    dt = time.perf_counter() - t0
    checks = [None] * n_trials
    if cfg.get("analysis", True) and traj and "positions" in traj:
        # post-solve analysis on the device (outside time_sec): continuous-time separation and dynamics residual
        from ..analysis import check_trajectories

        checks = check_trajectories(traj, np.stack(starts), np.stack(goals), cfg["time_step"], cfg["space_dims"],
                                    min_distance=cfg["min_distance"])
    # copies of one trial: 4 (N,2) inputs up, 3 (N,K,2) outputs + the record down, at the PCIe rate the batch saw
    # (bounded by what is left of the wall time after the longest scenario)
    longest = max(r["device_ns"] for r in recs) * 1e-9
    copy_share = max(0.0, dt - longest) / n_trials
    out = []
    for t, r in enumerate(recs):
        failed = r["status"] == 1
        out.append({
            "N": N, "status": "error" if failed else "success", "time_sec": r["device_ns"] * 1e-9 + copy_share,
            "error": "OSQP failed: initial QP not solved" if failed else None,
            "K": solver.K, "T": cfg["time_horizon"], "h": cfg["time_step"], "trial_index": lo + t,
            "batch_time_sec": dt, "device_time_sec": r["device_ns"] * 1e-9, "seed": used[t], "rank": rank,
            "scp_iterations": r["scp_iterations"], "admm_iterations": r["admm_iterations"],
            "qp_unsolved": r["qp_unsolved"], "min_separation": r["min_separation"],
            "min_separation_continuous": checks[t]["min_separation_continuous"] if checks[t] else None,
            "dynamics_residual": checks[t]["dynamics_residual"] if checks[t] else None,
        })
    return gather_records(out) if world > 1 else out


def summarize(runs, Ns):
    summary = {}
    for N in Ns:
        times = [r["time_sec"] for r in runs if r["N"] == N and r["status"] == "success"]
        errors = sum(1 for r in runs if r["N"] == N and r["status"] != "success")
        if times:
            summary[str(N)] = {
                "count": len(times), "errors": errors, "min": float(np.min(times)), "max": float(np.max(times)),
                "mean": float(np.mean(times)), "median": float(np.median(times)),
                "p25": float(np.percentile(times, 25)), "p75": float(np.percentile(times, 75)),
                "std": float(np.std(times, ddof=1)) if len(times) > 1 else 0.0,
            }
        else:
            summary[str(N)] = {"count": 0, "errors": errors, "min": None, "max": None, "mean": None,
                               "median": None, "p25": None, "p75": None, "std": None}
    return summary


def main(config=None):
    cfg = CONFIG.copy()
    yaml_path = os.environ.get("SCP_BATCH_CONFIG")
    if config is None and len(sys.argv) > 1 and sys.argv[1].endswith((".yaml", ".yml")):
        yaml_path = sys.argv[1]
    if yaml_path:
        cfg.update(load_config(yaml_path))
    if config:
        cfg.update(config)
    rank, world = _dist_context()
    if world > 1 and not cfg.get("batched", True):
        raise ValueError("a torch.distributed launch shards batched trials: set batched: true")
    if rank != 0:                       # the other ranks only solve their shard of every N
        for N in cfg["Ns"]:
            seeds = None
            if cfg["rng_seed"] is not None:
                seeds = [cfg["rng_seed"] + 1000 * N + t for t in range(cfg["trials_per_N"])]
            run_batch(N, cfg, cfg["trials_per_N"], seeds)
        return None
    Path(cfg["results_dir"]).mkdir(parents=True, exist_ok=True)
    stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
    json_path = Path(cfg["results_dir"]) / f"scp_benchmark_{stamp}.json"
    csv_path = Path(cfg["results_dir"]) / f"scp_benchmark_{stamp}.csv"
    if cfg["rng_seed"] is not None:
        np.random.seed(cfg["rng_seed"])
    print("------ WOW SCP Benchmark ------")
    print(f"Robot counts: {cfg['Ns']}, Trials per N: {cfg['trials_per_N']}")
    print(f"T={cfg['time_horizon']}s, h={cfg['time_step']}s, R={cfg['min_distance']}m, space={cfg['space_dims']}")
    print(f"Max SCP iterations: {cfg['max_iterations']}")
    print()
    all_results = {
        "meta": {"timestamp": stamp,
                 "description": "SCP timing benchmark for multiple N; each entry is a full solve wall time.",
                 "config": cfg, "schema_version": "1.0"},
        "runs": [], "summary": {},
    }
    for N in cfg["Ns"]:
        print(f"==> N = {N}")
        seeds = None
        if cfg["rng_seed"] is not None:
            seeds = [cfg["rng_seed"] + 1000 * N + t for t in range(cfg["trials_per_N"])]
        if cfg.get("batched", True):
            runs = run_batch(N, cfg, cfg["trials_per_N"], seeds)
        else:
            runs = []
            for trial in range(cfg["trials_per_N"]):
                if seeds is not None:
                    np.random.seed(seeds[trial])
                    random.seed(seeds[trial])
                res = run_single_trial(N, cfg, rng=np.random)
                res["trial_index"] = trial
                res["seed"] = None if seeds is None else seeds[trial]
                runs.append(res)
        for res in runs:
            all_results["runs"].append(res)
            status_str = "OK" if res["status"] == "success" else f"ERR ({res['error']})"
            print(f"  trial {res['trial_index']+1:02d}/{cfg['trials_per_N']}  time = {res['time_sec']:.3f}s  [{status_str}]")
        print()
    all_results["summary"] = summarize(all_results["runs"], cfg["Ns"])
    with open(json_path, "w", encoding="utf-8") as f:
        json.dump(all_results, f, indent=2)
    print(f"Saved JSON: {json_path}")
    fieldnames = ["N", "trial_index", "status", "time_sec", "K", "T", "h", "error"]
    with open(csv_path, "w", newline="", encoding="utf-8") as f:
        w = csv.DictWriter(f, fieldnames=fieldnames)
        w.writeheader()
        for r in all_results["runs"]:
            w.writerow({k: r.get(k, None) for k in fieldnames})
    print(f"Saved CSV:  {csv_path}")
    print("\nSummary (success-only times):")
    for N in cfg["Ns"]:
        s = all_results["summary"][str(N)]
        print(f"  N={N}: count={s['count']}, errors={s['errors']}, mean={s['mean']}, median={s['median']}, "
              f"p25={s['p25']}, p75={s['p75']}")
    return all_results


if __name__ == "__main__":
    main()
