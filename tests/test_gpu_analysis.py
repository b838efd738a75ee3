"""GPU tests of the steps either side of the solve: the device scenario generator against the
host generators' acceptance rules, and the post-solve analysis kernel against the oracle's checks and a dense numerical
evaluation of the continuous-time separation."""
import numpy as np
import pytest

from conftest import golden_cases, golden_limits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def _pair_min(points):
    iu, ju = np.triu_indices(points.shape[0], 1)
    return np.linalg.norm(points[iu] - points[ju], axis=-1).min()


def test_device_generator_reference_layout_rules(torch_cuda):
    """Layout of reference position_generator.py:18-75: starts ON the four corner circles, goals on the diamond border
    (or, 10 %, on the circles), pairwise spacing >= min_distance in each set; reproducible per (seed, index)."""
    from path_planning.scenarios.device_generator import generate_scenarios_device

    B, N, R = 256, 25, 0.8
    p0, pf, ok = generate_scenarios_device(B, N, R, "reference", seed=7)
    p0, pf, ok = p0.cpu().numpy(), pf.cpu().numpy(), ok.cpu().numpy()
    assert ok.mean() >= 0.9                     # the host generator also fails now and then at 25 agents / 1000 attempts
    centers = np.array([[3.5, 3.5], [16.5, 3.5], [3.5, 16.5], [16.5, 16.5]])
    dsz = 6.0 / np.sqrt(2.0)
    on_diamond = 0
    for b in np.flatnonzero(ok):
        assert _pair_min(p0[b]) >= R - 1e-12 and _pair_min(pf[b]) >= R - 1e-12
        d0 = np.linalg.norm(p0[b][:, None, :] - centers[None], axis=-1).min(axis=1)
        assert np.abs(d0 - 2.5).max() <= 1e-9
        df = np.linalg.norm(pf[b][:, None, :] - centers[None], axis=-1).min(axis=1)
        diamond = np.abs(np.abs(pf[b][:, 0] - 10.0) + np.abs(pf[b][:, 1] - 10.0) - dsz) <= 1e-9
        assert np.all(diamond | (np.abs(df - 2.5) <= 1e-9))
        on_diamond += diamond.sum()
    # 90 % of the goal DRAWS aim at the diamond; its 24 m border is crowded at 25 x 0.8 m, so fewer of the accepted ones do
    assert 0.5 <= on_diamond / (ok.sum() * N) <= 0.97
    q0, qf, ok2 = generate_scenarios_device(B, N, R, "reference", seed=7)
    assert np.array_equal(q0.cpu().numpy()[ok], p0[ok]) and np.array_equal(ok2.cpu().numpy(), ok)
    r0, _, _ = generate_scenarios_device(B, N, R, "reference", seed=8)
    assert not np.array_equal(r0.cpu().numpy(), p0)
    # scenario index, not position in the batch, selects the stream: a shard of the batch reproduces its slice
    s0, _, _ = generate_scenarios_device(64, N, R, "reference", seed=7, first_scenario=100)
    assert np.array_equal(s0.cpu().numpy(), p0[100:164])


def test_device_generator_large_layout_rules_and_solve(torch_cuda):
    """Bounded-travel layout (generate_positions_large): inside the arena with a 1 m rim, spacing >= 1.25 R in both sets,
    travel in [0.5, 1] x 0.4 v_max T; the generated scenarios go straight into the batched solver on the device."""
    from path_planning.scenarios.device_generator import generate_scenarios_device, space_dims_for
    from path_planning.solvers.batch import BatchSolver

    B, N, R, T = 64, 100, 0.8, 10.0
    p0, pf, ok = generate_scenarios_device(B, N, R, "large", seed=3, time_horizon=T)
    assert bool(ok.all())
    a, g = p0.cpu().numpy(), pf.cpu().numpy()
    side = space_dims_for("large", N)[2]
    assert abs(side - np.sqrt(16.0 * N)) < 1e-12
    for b in range(B):
        assert _pair_min(a[b]) >= 1.25 * R - 1e-12 and _pair_min(g[b]) >= 1.25 * R - 1e-12
        assert a[b].min() >= 1.0 and a[b].max() <= side - 1.0 and g[b].min() >= 1.0 and g[b].max() <= side - 1.0
        d = np.linalg.norm(g[b] - a[b], axis=1)
        assert d.min() >= 0.5 * 0.4 * 2.0 * T - 1e-9 and d.max() <= 0.4 * 2.0 * T + 1e-9
    s = BatchSolver(N, T, 0.2, R, space_dims_for("large", N))
    acc, pos, vel, rec = s.solve_device(p0[:8], pf[:8])
    torch_cuda.cuda.synchronize()
    recs = BatchSolver.records_from_bytes(rec)
    assert all(r["status"] == 0 and r["min_separation"] >= R - 0.01 for r in recs)


def _dense_continuous_min(acc, pos, vel, h, sub=400):
    """min over pairs and t of |p_i(t) - p_j(t)| by dense sampling of the constant-acceleration segments."""
    N, K, _ = pos.shape
    t = np.linspace(0.0, h, sub + 1)[None, None, :, None]
    seg = pos[:, :-1, None, :] + vel[:, :-1, None, :] * t + 0.5 * acc[:, :-1, None, :] * t * t     # (N, K-1, sub+1, 2)
    best = np.inf
    for i in range(N):
        d = np.linalg.norm(seg[i][None] - seg[i + 1:], axis=-1)
        if d.size:
            best = min(best, d.min())
    return best


@pytest.mark.parametrize("path", golden_cases(max_agents=25))
def test_check_kernel_matches_oracle_and_dense_sampling(torch_cuda, path):
    from oracle import scp_oracle
    from path_planning.analysis import check_trajectories

    g = np.load(path)
    N, h, R, space = int(g["N"]), float(g["h"]), float(g["R"]), list(g["space"])
    lim = golden_limits(g)
    tr = {k: g[k] for k in ("positions", "velocities", "accelerations")}
    r = check_trajectories(tr, g["p0"], g["pf"], h, space, min_distance=R, **lim)[0]
    assert abs(r["min_separation"] - scp_oracle.min_separation(g["positions"])) <= 1e-12
    z = np.zeros((N, 2))
    dyn = scp_oracle.dynamics_residual(g["accelerations"], g["p0"], z, g["pf"], z, h, space, positions=g["positions"],
                                       vlim=lim.get("vel_limit", 2.0), alim=lim.get("acc_limit", 15.0), jlim=lim.get("jerk_limit", 20.0))
    assert r["dynamics_pass"] == (dyn <= 1e-3) and r["dynamics_residual"] <= 1e-8 and dyn <= 1e-8
    if N > 1:
        dense = _dense_continuous_min(g["accelerations"], g["positions"], g["velocities"], h)
        assert r["min_separation_continuous"] <= r["min_separation"] + 1e-15
        assert r["min_separation_continuous"] <= dense + 1e-12              # the closed form is never above a sampled value
        assert dense - r["min_separation_continuous"] <= 1e-5               # ... and the dense sampling converges to it
        assert 0.0 <= r["min_separation_continuous_time"] <= (g["positions"].shape[1] - 1) * h + 1e-12
    print(f"{path.split('/')[-1]}: sampled min separation {r['min_separation']:.6f}, continuous {r['min_separation_continuous']:.6f} "
          f"at t = {r['min_separation_continuous_time']:.3f} s")


def test_check_kernel_flags_violations(torch_cuda):
    """Perturbed trajectories: a broken recursion, a box violation and a missed goal are each reported in their field."""
    from path_planning.analysis import check_trajectories

    g = np.load([p for p in golden_cases() if p.endswith("n5_s0.npz")][0])
    h, space = float(g["h"]), list(g["space"])
    base = {k: g[k].copy() for k in ("positions", "velocities", "accelerations")}
    ok = check_trajectories(base, g["p0"], g["pf"], h, space)[0]
    assert ok["dynamics_residual"] <= 1e-8
    t = {k: v.copy() for k, v in base.items()}
    t["positions"][2, 17, 0] += 0.25
    r = check_trajectories(t, g["p0"], g["pf"], h, space)[0]
    assert abs(r["dynamics_violation"] - 0.25) <= 1e-9 and r["box_violation"] <= 1e-9 and not r["dynamics_pass"]
    r = check_trajectories(base, g["p0"], g["pf"], h, space, acc_limit=0.5)[0]
    assert abs(r["box_violation"] - (np.abs(g["accelerations"]).max() - 0.5)) <= 1e-12
    r = check_trajectories(base, g["p0"], g["pf"] + 0.5, h, space)[0]
    assert abs(r["terminal_violation"] - 0.5) <= 1e-8


def test_check_kernel_batch_on_reference_outcomes(torch_cuda):
    """A batch call: the verbatim reference's own trajectories on the benchmark seeds (tests/golden/c2_outcomes.npz);
    the device analysis reproduces the min-separation / dynamics pass-fail flags recorded with the oracle's checks."""
    import os

    from conftest import GOLDEN
    from path_planning.analysis import check_trajectories

    f = np.load(os.path.join(GOLDEN, "c2_outcomes.npz"))
    acc, pos = f["accelerations"], f["positions"]
    vel = np.zeros_like(pos)
    vel[:, :, 1:] = 0.2 * np.cumsum(acc, axis=2)[:, :, :-1]
    res = check_trajectories({"positions": pos, "velocities": vel, "accelerations": acc}, f["p0"], f["pf"], 0.2,
                             [0, 0, 20, 20], min_distance=0.8)
    assert len(res) == len(f["seed"])
    for b, r in enumerate(res):
        assert abs(r["min_separation"] - float(f["min_separation"][b])) <= 1e-12
        assert r["min_separation_pass"] == bool(f["minsep_pass"][b])
        assert r["dynamics_pass"] == bool(f["dyn_pass"][b])
        assert r["min_separation_continuous"] <= r["min_separation"] + 1e-15
