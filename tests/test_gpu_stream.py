"""GPU parity tests of the streaming solver (scp_b200_stream_*, csrc/scp_stream.cu) through the C ABI.

Same bar as test_gpu_parity.py: final positions <= 1e-3 relative and objective <= 1e-4 relative against the
KKT-certified golden fixtures, identical pass/fail on the minimum-separation and dynamics-residual checks.
The streaming solver ends its subproblems on the ADMM residual test (no polish), so agreement with the certified
minimisers is to solver tolerance, not to rounding; the SCP iteration count must still be the reference's.
"""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_cases, golden_limits

pytestmark = pytest.mark.gpu

POS_TOL, OBJ_TOL, DYN_TOL = 1e-3, 1e-4, 1e-3


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def _solve(p0, pf, T, h, R, space, **kw):
    from path_planning.solvers.stream import StreamSolver

    p0 = np.asarray(p0, dtype=np.float64)
    if p0.ndim == 2:
        p0, pf = p0[None], np.asarray(pf)[None]
    s = StreamSolver(p0.shape[1], T, h, R, space, n_scenarios=p0.shape[0], **kw)
    traj, recs = s.solve(p0, pf)
    ms = s.last_device_ms
    s.close()
    return traj, recs, ms


@pytest.mark.parametrize("path", golden_cases())
def test_stream_matches_golden(torch_cuda, path):
    from oracle import scp_oracle

    g = np.load(path)
    N, h, R, space = int(g["N"]), float(g["h"]), float(g["R"]), list(g["space"])
    if round(float(g["T"]) / h) > 128:
        pytest.skip("streaming solver: n_steps <= 128 (include/scp_b200.h)")
    lim = golden_limits(g)                               # lazy_rows stays on: a binding box class has to join
    traj, recs, ms = _solve(g["p0"], g["pf"], float(g["T"]), h, R, space, **lim)
    r, pos, acc = recs[0], traj["positions"][0], traj["accelerations"][0]
    assert r["status"] == 0 and r["qp_unsolved"] == 0
    assert r["scp_iterations"] == int(g["iterations"])
    assert r["initial_feasible"] == (int(g["iterations"]) == 0)
    assert np.allclose(r["rel_steps"], g["rel_steps"], rtol=2e-2, atol=5e-4)
    perr = np.linalg.norm(pos - g["positions"]) / np.linalg.norm(g["positions"])
    oerr = abs(r["objective"] - float(g["objective"])) / float(g["objective"])
    assert perr <= POS_TOL, perr
    assert oerr <= OBJ_TOL, oerr
    z = np.zeros((N, 2))
    assert (r["min_separation"] >= R - 0.01) == (float(g["min_separation"]) >= R - 0.01)
    assert abs(r["min_separation"] - scp_oracle.min_separation(pos)) <= 1e-12
    assert abs(r["objective"] - (acc ** 2).sum()) <= 1e-9 * r["objective"]
    dyn = scp_oracle.dynamics_residual(acc, g["p0"], z, g["pf"], z, h, space, positions=pos,
                                       vlim=lim.get("vel_limit", 2.0), alim=lim.get("acc_limit", 15.0),
                                       jlim=lim.get("jerk_limit", 20.0))
    assert dyn <= DYN_TOL
    loose = np.linalg.norm(g["loose_positions"] - g["positions"]) / np.linalg.norm(g["positions"])
    print(f"{os.path.basename(path)}: stream pos err {perr:.1e} (loose reference {loose:.1e}), objective err {oerr:.1e}, "
          f"ADMM iterations {r['admm_iterations']}, device {ms:.1f} ms")


def test_public_class_streaming_engine(torch_cuda):
    """`SCP(...).generate_trajectories()` with engine="stream": same contract (result dict, record, prints) as the
    default engine, trajectories within tolerance of the certified golden; "auto" picks it for large scenarios."""
    from path_planning import SCP
    from path_planning.solvers.scp import use_stream_engine

    g = np.load([p for p in golden_cases() if p.endswith("n8_s1.npz")][0])
    s = SCP(n_vehicles=int(g["N"]), time_horizon=float(g["T"]), time_step=float(g["h"]), min_distance=float(g["R"]),
            space_dims=list(g["space"]))
    s.verbose = False
    s.engine = "stream"
    s.set_initial_states(g["p0"])
    s.set_final_states(g["pf"])
    tr = s.generate_trajectories(max_iterations=15)
    assert set(tr) == {"positions", "velocities", "accelerations"} and tr["positions"].shape == (int(g["N"]), 50, 2)
    perr = np.linalg.norm(tr["positions"] - g["positions"]) / np.linalg.norm(g["positions"])
    assert perr <= POS_TOL and s.last_record["scp_iterations"] == int(g["iterations"])
    assert use_stream_engine("auto", 200, 100) and not use_stream_engine("auto", 10, 500) and not use_stream_engine("auto", 25, 50)


def test_stream_batch_equals_single(torch_cuda):
    """Scenarios of a batch do not influence each other: batch results == one-by-one results, bit for bit."""
    from path_planning.scenarios.position_generator import generate_positions

    starts, goals = [], []
    for b in range(6):
        random.seed(200 + b)
        p0, pf = generate_positions(8, 0.8)
        starts.append(p0)
        goals.append(pf)
    starts, goals = np.stack(starts), np.stack(goals)
    tb, rb, _ = _solve(starts, goals, 10.0, 0.2, 0.8, [0, 0, 20, 20])
    for b in (0, 3, 5):
        t1, r1, _ = _solve(starts[b], goals[b], 10.0, 0.2, 0.8, [0, 0, 20, 20])
        assert np.array_equal(t1["positions"][0], tb["positions"][b])
        assert r1[0]["admm_iterations"] == rb[b]["admm_iterations"] and r1[0]["scp_iterations"] == rb[b]["scp_iterations"]


def test_stream_agrees_with_cta_solver_on_c2_scenarios(torch_cuda):
    """32 of config 2's scenarios: wherever both solvers solved every subproblem, trajectories agree to tolerance
    and the feasibility verdicts are the same."""
    from oracle import scp_oracle
    from path_planning.scenarios.position_generator import generate_positions
    from path_planning.solvers.batch import BatchSolver

    B, N, R, h = 32, 25, 0.8, 0.2
    starts, goals = [], []
    for b in range(B):
        random.seed(10_000 + b)
        p0, pf = generate_positions(N, R)
        starts.append(p0)
        goals.append(pf)
    starts, goals = np.stack(starts), np.stack(goals)
    ts, rs, _ = _solve(starts, goals, 10.0, h, R, [0, 0, 20, 20])
    tc, rc = BatchSolver(N, 10.0, h, R, [0, 0, 20, 20]).solve(starts, goals)
    z = np.zeros((N, 2))
    n_cmp = 0
    for b in range(B):
        assert rs[b]["status"] == 0
        dyn = scp_oracle.dynamics_residual(ts["accelerations"][b], starts[b], z, goals[b], z, h, [0, 0, 20, 20],
                                           positions=ts["positions"][b])
        assert dyn <= DYN_TOL, (b, dyn)
        if rs[b]["qp_unsolved"] == 0 and rc[b]["qp_unsolved"] == 0 and rs[b]["scp_iterations"] == rc[b]["scp_iterations"]:
            perr = np.linalg.norm(ts["positions"][b] - tc["positions"][b]) / np.linalg.norm(tc["positions"][b])
            assert perr <= POS_TOL, (b, perr)
            assert (rs[b]["min_separation"] >= R - 0.01) == (rc[b]["min_separation"] >= R - 0.01)
            n_cmp += 1
    assert n_cmp >= B // 3


def test_stream_large_scenario_properties(torch_cuda):
    """One 100-agent scenario (K=100, bounded-travel generator): dynamics feasible, separation holds, records consistent."""
    from oracle import scp_oracle
    from path_planning.scenarios.position_generator import generate_positions_large

    random.seed(10_000)
    N, T, h, R = 100, 20.0, 0.2, 0.8
    p0, pf, space = generate_positions_large(N, R, time_horizon=T)
    traj, recs, ms = _solve(p0, pf, T, h, R, space)
    r = recs[0]
    z = np.zeros((N, 2))
    assert r["status"] == 0
    dyn = scp_oracle.dynamics_residual(traj["accelerations"][0], p0, z, pf, z, h, space, positions=traj["positions"][0])
    assert dyn <= DYN_TOL
    assert abs(r["min_separation"] - scp_oracle.min_separation(traj["positions"][0])) <= 1e-12
    if r["converged"] and r["qp_unsolved"] == 0:
        assert r["min_separation"] >= R - 0.01
    print(f"100 agents: {r['scp_iterations']} SCP iterations, {r['admm_iterations']} ADMM iterations, device {ms:.1f} ms, "
          f"min separation {r['min_separation']:.4f}")


@pytest.mark.parametrize("N,K", [(1, 20), (2, 15), (6, 31), (6, 33), (5, 64), (4, 127), (4, 128)])
def test_stream_step_counts_and_tiny_scenarios(torch_cuda, N, K):
    """Every lane layout of k_iter (EPL = 1..4, odd K = scalar rows, K = 128 = all lanes full) and the degenerate
    sizes (one agent: no pair rows at all) against the one-CTA solver, which ends on KKT certificates."""
    from oracle import scp_oracle
    from path_planning.solvers.batch import BatchSolver

    rng = np.random.default_rng(100 * N + K)
    h, space = 0.2, [0, 0, 12, 12]
    T = K * h + 1e-9
    # starts on the left, goals on the right in reversed order: paths cross; displacement L per axis stays inside what
    # |v| <= 2 allows within the horizon (rest-to-rest peak velocity 1.5 d / T), R shrinks with L so starts are apart
    L = min(8.0, 0.4 * T)
    R = min(0.8, 0.6 * L / max(N, 2))
    ys = np.linspace(6.0 - L / 2, 6.0 + L / 2, N) if N > 1 else np.array([6.0])
    p0 = np.stack([np.full(N, 2.0) + rng.uniform(-0.02, 0.02, N), ys], axis=1)
    pf = np.stack([np.full(N, 2.0 + L) + rng.uniform(-0.02, 0.02, N), ys[::-1]], axis=1)
    traj, recs, _ = _solve(p0, pf, T, h, R, space)
    assert traj["positions"].shape == (1, N, K, 2)
    r = recs[0]
    tc, rc = BatchSolver(N, T, h, R, space).solve(p0[None], pf[None])
    z = np.zeros((N, 2))
    assert r["status"] == rc[0]["status"] == 0
    assert r["initial_feasible"] == rc[0]["initial_feasible"]
    dyn = scp_oracle.dynamics_residual(traj["accelerations"][0], p0, z, pf, z, h, space, positions=traj["positions"][0])
    assert dyn <= DYN_TOL
    if N > 1:
        assert abs(r["min_separation"] - scp_oracle.min_separation(traj["positions"][0])) <= 1e-12
    if r["qp_unsolved"] == 0 and rc[0]["qp_unsolved"] == 0 and r["scp_iterations"] == rc[0]["scp_iterations"]:
        perr = np.linalg.norm(traj["positions"][0] - tc["positions"][0]) / np.linalg.norm(tc["positions"][0])
        assert perr <= POS_TOL, perr
        assert (r["min_separation"] >= R - 0.01) == (rc[0]["min_separation"] >= R - 0.01)
    else:
        assert N > 1          # a single agent has nothing that could stay unsolved


def test_stream_failure_paths(torch_cuda):
    """scp.py:846-865 demo inputs (first avoidance QP infeasible: warning + loop continues, scp.py:446-449) and two
    agents that start closer than R (k = 0 rows infeasible): no exception, no hang, the record says what happened."""
    p0 = np.array([[-2.0, -2], [0, -2], [2, -2]])
    pf = np.array([[2.0, 2], [0, 2], [-2, 2]])
    traj, recs, _ = _solve(p0, pf, 3.0, 0.2, 0.5, [-5, -5, 500, 200], max_admm_iter=3000)
    r = recs[0]
    assert r["status"] == 0 and not r["initial_feasible"]
    assert r["first_violation"][1:] == (0, 1) and r["first_violation_dist"] < 0.49
    assert r["qp_unsolved"] >= 1 and r["scp_iterations"] >= 1
    assert np.isfinite(traj["accelerations"]).all() and r["min_separation"] < 0.5 - 0.01
    p0 = np.array([[5.0, 5.0], [5.3, 5.0]])
    pf = np.array([[15.0, 5.0], [15.0, 8.0]])
    _, recs, _ = _solve(p0, pf, 10.0, 0.2, 0.8, [0, 0, 20, 20], max_admm_iter=2000)
    assert recs[0]["status"] == 2 and recs[0]["first_violation"] == (0, 0, 1)


def test_stream_agent_sharded_matches_single_gpu(torch_cuda):
    """World-size-2 agent-sharded solve (NCCL all-gather of positions per ADMM iteration) == the one-GPU solve."""
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29631", os.path.join(ROOT, "tools", "run_sharded_scenario.py"), "--agents", "30", "--check"],
        capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_CHECK_OK" in out.stdout


def test_stream_agent_sharded_peer_exchange_matches_single_gpu(torch_cuda):
    """Same as above with the peer-memory exchange (NVLink stores + flags inside k_iter) instead of NCCL per iteration."""
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29633", os.path.join(ROOT, "tools", "run_sharded_scenario.py"), "--agents", "30", "--check", "--peer"],
        capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_CHECK_OK" in out.stdout and '"exchange": "peer"' in out.stdout
