"""Host-side logic on CPU: reference-facing class surface, batch records/summary schema, tables,
and the N>1 scenario sharding with a world_size-2 gloo group."""
import io
import os
import contextlib

import numpy as np
import pytest

from path_planning import SCP
from path_planning.cli import compute_trajectories_batch as ctb
from path_planning.solvers.sharding import shard_range


def test_scp_surface_matches_reference():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        s = SCP(n_vehicles=10, time_horizon=100, time_step=0.2, min_distance=0.8, space_dims=[0, 0, 200, 200])
    assert buf.getvalue().splitlines()[0] == "---=== SCP Problem initialized ===---"
    assert len(buf.getvalue().splitlines()) == 5
    assert (s.N, s.K, s.T, s.h, s.R) == (10, 500, 100, 0.2, 0.8)
    assert s.convergence_tolerance == 1.5e-2 and (s.vel_max, s.acc_max, s.jerk_max) == (2, 15.0, 20)
    assert np.array_equal(s.pos_max, [200, 200]) and s.trajectories is None
    s.set_initial_states(np.zeros((10, 2)))
    assert s.initial_velocities.shape == (20,)
    with pytest.raises(AssertionError):
        s.set_final_states(np.zeros((9, 2)))
    with pytest.raises(ValueError):
        s.visualize_trajectories()
    with pytest.raises(ValueError):
        s.visualize_time_snapshots()


def test_batch_config_and_summary_schema():
    assert set(["Ns", "trials_per_N", "time_horizon", "time_step", "min_distance", "space_dims", "max_iterations",
                "rng_seed", "results_dir"]) <= set(ctb.CONFIG)
    runs = [dict(N=5, status="success", time_sec=t) for t in (1.0, 2.0, 4.0)] + [dict(N=5, status="error", time_sec=9.0)]
    s = ctb.summarize(runs, [5, 7])
    assert s["5"]["count"] == 3 and s["5"]["errors"] == 1 and s["5"]["median"] == 2.0
    assert abs(s["5"]["std"] - np.std([1, 2, 4], ddof=1)) < 1e-12
    assert s["7"] == {"count": 0, "errors": 0, "min": None, "max": None, "mean": None, "median": None, "p25": None,
                      "p75": None, "std": None}


def test_shard_range_partitions():
    for n in (0, 1, 7, 1024, 1025):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from path_planning.solvers.sharding import solve_scenarios_sharded

    p0 = np.arange(7 * 3 * 2, dtype=float).reshape(7, 3, 2)

    def fake_solve(a, b):  # stands in for BatchSolver.solve on this rank's GPU
        return {"positions": a.copy()}, [{"objective": float(x.sum())} for x in a]

    traj, recs = solve_scenarios_sharded(fake_solve, p0, p0)
    q.put((rank, traj["positions"].shape[0], [(r["scenario_index"], r["rank"], r["objective"]) for r in recs]))
    dist.destroy_process_group()


def test_scenario_sharding_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    p0 = np.arange(7 * 3 * 2, dtype=float).reshape(7, 3, 2)
    assert [r[1] for r in res] == [4, 3]
    for _, _, recs in res:
        assert [r[0] for r in recs] == list(range(7))
        assert [r[1] for r in recs] == [0, 0, 0, 0, 1, 1, 1]
        assert np.allclose([r[2] for r in recs], p0.sum(axis=(1, 2)))


def test_agent_blocks_balance_pair_rows():
    from path_planning.solvers.sharded import agent_blocks, pair_index

    for n in (2, 25, 200, 1000):
        total = n * (n - 1) // 2
        for w in (1, 2, 4, 8):
            b = agent_blocks(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[g] <= b[g + 1] for g in range(w))
            ends = [pair_index(b[g], n) if b[g] < n else total for g in range(w + 1)]
            rows = [ends[g + 1] - ends[g] for g in range(w)]
            assert sum(rows) == total
            if n >= 200:
                assert max(rows) - min(rows) <= 2 * n          # balanced to within two agents' worth of rows
    # pair_index is the reference's lexicographic i<j order (scp.py:495-496)
    n = 7
    order = [(i, j) for i in range(n) for j in range(i + 1, n)]
    for p, (i, j) in enumerate(order):
        assert pair_index(i, n) + (j - i - 1) == p


def test_stream_agent_block_partition():
    """Equal blocks of ceil(N/world) agents (the per-iteration position all-gather needs equal slices); the tail
    rank may own fewer, never a negative count; every agent is owned exactly once."""
    from path_planning.solvers.stream import agent_block

    for n in (1, 5, 30, 200, 1000, 1001):
        for w in (1, 2, 3, 4, 8):
            blocks = [agent_block(n, r, w) for r in range(w)]
            per = -(-n // w)
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(lo <= hi and hi - lo <= per for lo, hi in blocks)
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(w - 1))
            assert sum(hi - lo for lo, hi in blocks) == n


def _id_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the plumbing StreamSolver uses to hand rank 0's 128-byte NCCL id to the other ranks
    box = [bytes(range(128)) if rank == 0 else bytes(128)]
    dist.broadcast_object_list(box, src=0)
    from path_planning.solvers.stream import agent_block

    q.put((rank, box[0] == bytes(range(128)), agent_block(30, rank, world)))
    dist.destroy_process_group()


def test_stream_id_broadcast_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True, (0, 15)), (1, True, (15, 30))]


def test_stream_solver_has_no_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from path_planning import _capi
    from path_planning.solvers.stream import StreamSolver

    with pytest.raises(_capi.ScpB200Error):
        StreamSolver(5, 10.0, 0.2, 0.8)


# ------------------------------------------------------------------ batch CLI: YAML config, rank sharding, time_sec
def test_yaml_config_loader(tmp_path):
    cfgs = os.path.join(os.path.dirname(ctb.__file__), "..", "configs")
    c = ctb.load_config(os.path.join(cfgs, "batch_default.yaml"))
    assert c == {k: ctb.CONFIG[k] for k in c} and set(c) == set(ctb.CONFIG)        # the default file IS the CONFIG dict
    c2 = ctb.load_config(os.path.join(cfgs, "config2_1024x25.yaml"))
    assert c2["Ns"] == [25] and c2["trials_per_N"] == 1024
    bad = tmp_path / "bad.yaml"
    bad.write_text("trials: 3\n")
    with pytest.raises(ValueError):
        ctb.load_config(str(bad))


class _FakeBatchSolver:
    """Stands in for BatchSolver on a machine without a GPU: scenario b 'costs' (1 + b) ms of device time."""

    def __init__(self, N, T, h, R, space, **kw):
        self.K = int(T / h)

    def solve(self, p0, pf):
        recs = [dict(status=0, device_ns=int(1e6 * (1 + p0[b, 0, 0])), scp_iterations=3, admm_iterations=100,
                     qp_unsolved=0, min_separation=0.8) for b in range(len(p0))]
        return {}, recs


def _cli_worker(rank, world, port, results_dir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import path_planning.solvers.batch as batch
    from path_planning.cli import compute_trajectories_batch as c

    batch.BatchSolver = _FakeBatchSolver
    c.generate_positions = lambda N, R: (np.full((N, 2), float(c.random.random())), np.zeros((N, 2)))
    out = c.main({"Ns": [4], "trials_per_N": 5, "rng_seed": 7, "results_dir": results_dir, "engine": "cta"})
    q.put((rank, None if out is None else [(r["trial_index"], r["rank"], r["seed"], r["time_sec"], r["batch_time_sec"]) for r in out["runs"]]))
    import torch.distributed as dist

    dist.destroy_process_group()


def test_batch_cli_shards_trials_over_ranks_gloo(tmp_path):
    """compute-trajectories-batch under a 2-rank launch: each rank solves its shard of the trials, rank 0 gathers all
    records and writes the files; time_sec is per trial (device time of that scenario + copy share), not batch/trials."""
    import glob
    import json

    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_cli_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert res[1] is None                                   # only rank 0 reports / writes
    runs = res[0]
    assert [r[0] for r in runs] == [0, 1, 2, 3, 4] and [r[1] for r in runs] == [0, 0, 0, 1, 1]
    assert [r[2] for r in runs] == [7 + 4000 + t for t in range(5)]
    times = [r[3] for r in runs]
    assert len(set(round(t, 9) for t in times)) > 1         # spread kept: not batch wall / trials
    assert all(0 < t <= r[4] + 1e-3 + 2e-3 for t, r in zip(times, runs))
    files = glob.glob(str(tmp_path / "scp_benchmark_*.json"))
    assert len(files) == 1
    data = json.load(open(files[0]))
    assert len(data["runs"]) == 5 and data["summary"]["4"]["count"] == 5 and data["meta"]["schema_version"] == "1.0"


def test_bench_parity_block_on_the_fixture_itself():
    """bench.py's parity bookkeeping, fed with the reference's own outcomes as if they were the device's: every
    confusion matrix is diagonal and every error is zero."""
    import importlib.util

    from conftest import GOLDEN, ROOT

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    f = np.load(os.path.join(GOLDEN, "c2_outcomes.npz"))
    n = 12
    qs = f["qp_status"]
    recs = [dict(status=0, qp_unsolved=int((qs[b, : f["n_qp"][b]] != 1).sum()), objective=float(f["objective"][b]),
                 scp_iterations=int(f["scp_iterations"][b])) for b in range(n)]
    w = bench.WORKLOADS["c2"]
    out = bench.parity_block(w, recs, f["accelerations"][:n], f["positions"][:n], f["p0"][:n], f["pf"][:n])
    assert out["scenarios"] == n and out["status"]["both_success"] == n
    for key in ("finite_trajectory", "dynamics_residual_pass_1e-3", "min_separation_pass_R-0.01"):
        assert out[key]["ref_pass_gpu_fail"] == 0 and out[key]["ref_fail_gpu_pass"] == 0, (key, out[key])
    assert out["both_all_solved"]["position_rel_err"]["max"] == 0.0
    assert out["scp_iterations_equal"]["all"] == n
    assert out["subproblems"]["both_all_solved"] + out["subproblems"]["both_unsolved"] == n
    # the bench's own restatement of the two feasibility checks agrees with the oracle's
    from oracle import scp_oracle

    z = np.zeros((25, 2))
    for b in range(3):
        a, p = f["accelerations"][b], f["positions"][b]
        assert abs(bench.dynamics_residual(a, p, f["p0"][b], f["pf"][b], 0.2, [0, 0, 20, 20])
                   - scp_oracle.dynamics_residual(a, f["p0"][b], z, f["pf"][b], z, 0.2, [0, 0, 20, 20], positions=p)) <= 1e-12
        assert abs(bench.min_separation(p) - scp_oracle.min_separation(p)) <= 1e-15


def test_pyproject_declares_the_reference_console_scripts():
    """reference pyproject.toml:52-54: the entrypoint names of this path exist after `pip install`."""
    import importlib
    import tomllib

    from conftest import ROOT

    cfg = tomllib.load(open(os.path.join(ROOT, "pyproject.toml"), "rb"))
    scripts = cfg["project"]["scripts"]
    assert scripts["compute-trajectories"] == "path_planning.cli.compute_trajectories:main"
    assert scripts["compute-trajectories-batch"] == "path_planning.cli.compute_trajectories_batch:main"
    for target in scripts.values():
        mod, fn = target.split(":")
        assert callable(getattr(importlib.import_module(mod), fn))
    assert cfg["tool"]["setuptools"]["packages"]["find"]["where"] == ["ba-path-planning_b200"]
