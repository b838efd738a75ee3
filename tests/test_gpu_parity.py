"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libscp_b200.so), against
the oracle (restated reference, KKT-certified subproblems) and the committed golden fixtures.

Tolerances are the north star's: final positions <= 1e-3 relative, objective <= 1e-4 relative,
identical pass/fail of the minimum-separation check (>= R - 0.01, scp.py:610) and of the
dynamics-residual check (<= 1e-3).  Index/ordering work (row order, first
violation) is compared exactly; fp64 kernels that restate closed-form maps to 1e-12.
"""
import ctypes as C
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, active_box_classes, golden_cases, golden_limits

pytestmark = pytest.mark.gpu

POS_TOL, OBJ_TOL, DYN_TOL = 1e-3, 1e-4, 1e-3


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def lib():
    from path_planning import _capi

    return _capi.load()


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


# ------------------------------------------------------------------ reconstruct (scp.py:371-397, 559-595)
@pytest.mark.parametrize("B,N,K", [(1, 1, 2), (1, 3, 15), (3, 5, 50), (2, 10, 500), (1, 200, 100), (4, 7, 33)])
def test_reconstruct_matches_oracle(torch_cuda, lib, B, N, K):
    from oracle import scp_oracle
    from path_planning import _capi

    torch = torch_cuda
    rng = np.random.default_rng(B * 1000 + N * 10 + K)
    h = 0.2
    acc = rng.normal(size=(B, N, K, 2))
    p0 = rng.uniform(0, 20, (B, N, 2))
    v0 = rng.uniform(-1, 1, (B, N, 2))
    d_acc, d_p0, d_v0 = _dev(torch, acc), _dev(torch, p0), _dev(torch, v0)
    d_pos = torch.full((B, N, K, 2), float("nan"), dtype=torch.float64, device="cuda")
    d_vel = torch.full_like(d_pos, float("nan"))
    _capi.check(lib.scp_b200_reconstruct(d_acc.data_ptr(), d_p0.data_ptr(), d_v0.data_ptr(), B, N, K, h,
                                         d_pos.data_ptr(), d_vel.data_ptr(), None))
    torch.cuda.synchronize()
    for b in range(B):
        o = scp_oracle.ScpOracle(N, K * h + 1e-9, h, 0.8)
        o.K = K
        o.set_initial_states(p0[b], v0[b])
        pos, vel = o.states_from_accelerations(acc[b])
        scale = max(1.0, np.abs(pos).max())
        assert np.abs(d_pos[b].cpu().numpy() - pos).max() <= 1e-12 * scale * K
        assert np.abs(d_vel[b].cpu().numpy() - vel).max() <= 1e-12 * K


# ------------------------------------------------------------------ linearise (scp.py:453-557, 597-615)
def _linearize(torch, lib, pos, R, want_rows=True):
    from path_planning import _capi

    B, N, K, _ = pos.shape
    P = N * (N - 1) // 2
    d_pos = _dev(torch, pos)
    eta = torch.full((B, K, max(P, 1), 2), float("nan"), dtype=torch.float64, device="cuda")
    bound = torch.full((B, K, max(P, 1)), float("nan"), dtype=torch.float64, device="cuda")
    minsep = torch.zeros(B, dtype=torch.float64, device="cuda")
    first = torch.zeros((B, 3), dtype=torch.int32, device="cuda")
    _capi.check(lib.scp_b200_linearize(d_pos.data_ptr(), B, N, K, R, 0.01,
                                       eta.data_ptr() if want_rows else None, bound.data_ptr() if want_rows else None,
                                       minsep.data_ptr(), first.data_ptr(), None))
    torch.cuda.synchronize()
    return eta.cpu().numpy()[:, :, :P], bound.cpu().numpy()[:, :, :P], minsep.cpu().numpy(), first.cpu().numpy()


@pytest.mark.parametrize("B,N,K", [(1, 2, 2), (2, 5, 50), (1, 25, 50), (1, 64, 21), (1, 200, 100), (3, 10, 500)])
def test_linearize_rows_match_oracle_T3(torch_cuda, lib, B, N, K):
    from oracle import scp_oracle

    rng = np.random.default_rng(N * 7 + K)
    h, R = 0.2, 0.8
    acc = rng.normal(scale=0.3, size=(B, N, K, 2))
    p0 = rng.uniform(0, 20, (B, N, 2))
    pos = np.empty((B, N, K, 2))
    orc = []
    for b in range(B):
        o = scp_oracle.ScpOracle(N, K * h + 1e-9, h, R)
        o.K = K
        o.set_initial_states(p0[b])
        pos[b], _ = o.states_from_accelerations(acc[b])
        orc.append(o)
    eta, bound, minsep, first = _linearize(torch_cuda, lib, pos, R)
    iu, ju = np.triu_indices(N, 1)
    for b, o in enumerate(orc):
        d = np.transpose(pos[b][iu] - pos[b][ju], (1, 0, 2))          # (K,P,2), reference row order
        dist = np.hypot(d[..., 0], d[..., 1])
        assert np.abs(eta[b] - d / dist[..., None]).max() <= 1e-13
        assert np.abs(bound[b] - R).max() <= 1e-12                       # eta.d - dist == 0 up to rounding
        if N <= 25 and K <= 50:                                           # explicit rows of the oracle (scp.py:543-550)
            _, l_coll, _ = o.collision_rows(acc[b])
            shift = np.einsum("kpa,pa->kp", o._last_eta, p0[b][iu] - p0[b][ju])
            assert np.abs(bound[b] - (l_coll.reshape(K, -1) + shift)).max() <= 1e-11
        assert abs(minsep[b] - dist.min()) <= 1e-13
        feas = o.fast_check_avoidance(pos[b])
        if feas:
            assert tuple(first[b]) == (-1, -1, -1)
        else:
            assert tuple(first[b]) == tuple(o.record["first_violation"][:3])


def test_linearize_edge_cases(torch_cuda, lib):
    # N = 1: no rows, min separation = +inf, no violation
    _, _, minsep, first = _linearize(torch_cuda, lib, np.zeros((2, 1, 5, 2)), 0.8, want_rows=False)
    assert np.all(np.isinf(minsep)) and np.all(first == -1)
    # coincident agents: degenerate direction (deterministic stand-in for scp.py:503-507), bound = R - 1
    pos = np.zeros((1, 2, 3, 2))
    pos[0, 1, 1:] = [[3.0, 4.0], [0.3, 0.4]]
    eta, bound, minsep, first = _linearize(torch_cuda, lib, pos, 0.8)
    assert np.allclose(eta[0, 0, 0], [1.0, 0.0]) and abs(bound[0, 0, 0] - (0.8 - 1.0)) < 1e-15
    assert np.allclose(eta[0, 1, 0], [-0.6, -0.8]) and minsep[0] == 0.0
    assert tuple(first[0]) == (0, 0, 1)


# ------------------------------------------------------------------ full solve against the golden fixtures
def _solve_host(lib, p0, pf, T, h, R, space, **settings):
    from path_planning import _capi

    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    if p0.ndim == 2:
        p0, pf = p0[None], np.asarray(pf)[None]
    pf = np.ascontiguousarray(pf, dtype=np.float64)
    B, N, _ = p0.shape
    prob = _capi.default_problem(N, T, h, R, space)
    for k, v in settings.items():
        setattr(prob, k, v)
    K = prob.n_steps
    z = np.zeros_like(p0)
    acc, pos, vel = (np.full((B, N, K, 2), np.nan) for _ in range(3))
    recs = (_capi.Record * B)()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _capi.check(lib.scp_b200_solve_batch_host(C.byref(prob), B, ptr(p0), ptr(z), ptr(pf), ptr(z), ptr(acc), ptr(pos),
                                              ptr(vel), C.cast(recs, C.c_void_p), 0))
    return acc, pos, vel, [_capi.record_to_dict(r) for r in recs]


@pytest.mark.parametrize("path", golden_cases())
def test_solve_matches_golden(torch_cuda, lib, path):
    from oracle import scp_oracle

    g = np.load(path)
    N, h, R, space = int(g["N"]), float(g["h"]), float(g["R"]), list(g["space"])
    lim = golden_limits(g)
    acc, pos, vel, recs = _solve_host(lib, g["p0"], g["pf"], float(g["T"]), h, R, space, **lim)
    r = recs[0]
    assert r["status"] == 0 and r["qp_unsolved"] == 0
    assert r["scp_iterations"] == int(g["iterations"])
    assert r["initial_feasible"] == (int(g["iterations"]) == 0)
    assert np.allclose(r["rel_steps"], g["rel_steps"], rtol=1e-2, atol=2e-4)
    perr = np.linalg.norm(pos[0] - g["positions"]) / np.linalg.norm(g["positions"])
    oerr = abs(r["objective"] - float(g["objective"])) / float(g["objective"])
    assert perr <= POS_TOL, perr
    assert oerr <= OBJ_TOL, oerr
    z = np.zeros((N, 2))
    # same pass/fail on the two feasibility checks
    assert (r["min_separation"] >= R - 0.01) == (float(g["min_separation"]) >= R - 0.01)
    assert abs(r["min_separation"] - scp_oracle.min_separation(pos[0])) <= 1e-12
    dyn = scp_oracle.dynamics_residual(acc[0], g["p0"], z, g["pf"], z, h, space, positions=pos[0],
                                       vlim=lim.get("vel_limit", 2.0), alim=lim.get("acc_limit", 15.0),
                                       jlim=lim.get("jerk_limit", 20.0))
    assert dyn <= DYN_TOL
    # the box classes that bind at the reference's optimum bind here too (cases *_acc, *_jerk, *_pos; vel binds in several)
    got = dict(np.load(path))
    got.update(accelerations=acc[0], velocities=vel[0], positions=pos[0])

    class _G(dict):
        files = list(got)

    assert active_box_classes(_G(got), tol=1e-5) == active_box_classes(g, tol=1e-5)
    # the loose (eps 1e-3) reference itself sits further from the minimiser than we do
    loose = np.linalg.norm(g["loose_positions"] - g["positions"]) / np.linalg.norm(g["positions"])
    print(f"{path.split('/')[-1]}: pos err {perr:.1e} (loose reference {loose:.1e}), objective err {oerr:.1e}, "
          f"ADMM iterations {r['admm_iterations']}")


@pytest.mark.parametrize("name", ["n10_s2.npz", "n25_s3.npz"])
def test_team_mode_matches_golden(torch_cuda, lib, name):
    """Whole-grid cooperative kernel (the path large single scenarios take) forced on a small fixture."""
    import os

    from conftest import GOLDEN

    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    g = np.load(path)
    acc, pos, vel, recs = _solve_host(lib, g["p0"], g["pf"], float(g["T"]), float(g["h"]), float(g["R"]), list(g["space"]),
                                      team_mode=2)
    r = recs[0]
    assert r["status"] == 0 and r["scp_iterations"] == int(g["iterations"])
    assert np.linalg.norm(pos[0] - g["positions"]) / np.linalg.norm(g["positions"]) <= POS_TOL
    assert abs(r["objective"] - float(g["objective"])) <= OBJ_TOL * float(g["objective"])
    acc1, _, _, recs1 = _solve_host(lib, g["p0"], g["pf"], float(g["T"]), float(g["h"]), float(g["R"]), list(g["space"]),
                                    team_mode=1)
    assert recs1[0]["scp_iterations"] == r["scp_iterations"]
    assert np.abs(acc1 - acc).max() <= 1e-9 * max(1.0, np.abs(acc).max())


def test_solve_matches_live_oracle(torch_cuda, lib, truth_mode):
    """Seeded scenarios not in the fixtures, oracle run here (seconds)."""
    from oracle import scenarios, scp_oracle

    for seed, N in ((11, 4), (12, 6)):
        random.seed(seed)
        p0, pf = scenarios.generate_positions(N, 0.8)
        o = scp_oracle.ScpOracle(N, 10.0, 0.2, 0.8, [0, 0, 20, 20])
        o.set_initial_states(p0)
        o.set_final_states(pf)
        ref = o.generate_trajectories()
        acc, pos, vel, recs = _solve_host(lib, p0, pf, 10.0, 0.2, 0.8, [0, 0, 20, 20])
        assert recs[0]["scp_iterations"] == o.record["iterations"]
        assert np.linalg.norm(pos[0] - ref["positions"]) / np.linalg.norm(ref["positions"]) <= POS_TOL
        assert np.linalg.norm(vel[0] - ref["velocities"]) <= 1e-2 * max(1.0, np.linalg.norm(ref["velocities"]))
        obj = float((ref["accelerations"] ** 2).sum())
        assert abs(recs[0]["objective"] - obj) <= OBJ_TOL * obj


def test_initial_velocities_and_public_class(torch_cuda, lib, truth_mode):
    """SCP class surface (scp.py:99-180) with non-zero initial/final velocities."""
    from oracle import scp_oracle
    from path_planning import SCP

    p0 = np.array([[4.0, 4.0], [16.0, 5.0], [10.0, 15.0]])
    pf = np.array([[15.0, 14.0], [5.0, 13.0], [10.5, 4.0]])
    v0 = np.array([[0.3, 0.0], [-0.2, 0.1], [0.0, -0.4]])
    vf = np.array([[0.0, 0.2], [0.0, 0.0], [0.1, 0.0]])
    s = SCP(n_vehicles=3, time_horizon=14.0, time_step=0.2, min_distance=0.8)
    s.set_initial_states(p0, v0)
    s.set_final_states(pf, vf)
    tr = s.generate_trajectories(max_iterations=15)
    assert set(tr) == {"positions", "velocities", "accelerations"} and tr["positions"].shape == (3, s.K, 2)
    assert tr["positions"].dtype == np.float64 and s.trajectories is tr
    o = scp_oracle.ScpOracle(3, 14.0, 0.2, 0.8)
    o.set_initial_states(p0, v0)
    o.set_final_states(pf, vf)
    ref = o.generate_trajectories()
    assert s.last_record["scp_iterations"] == o.record["iterations"]
    assert np.linalg.norm(tr["positions"] - ref["positions"]) / np.linalg.norm(ref["positions"]) <= POS_TOL
    assert np.allclose(tr["positions"][:, 0], p0) and np.allclose(tr["velocities"][:, 0], v0)


def test_batch_equals_single_and_is_deterministic(torch_cuda, lib):
    from path_planning.scenarios.position_generator import generate_positions

    starts, goals = [], []
    for b in range(12):
        random.seed(500 + b)
        p0, pf = generate_positions(6, 0.8)
        starts.append(p0)
        goals.append(pf)
    starts, goals = np.stack(starts), np.stack(goals)
    acc_b, pos_b, _, recs_b = _solve_host(lib, starts, goals, 10.0, 0.2, 0.8, [0, 0, 20, 20])
    acc_b2, _, _, _ = _solve_host(lib, starts, goals, 10.0, 0.2, 0.8, [0, 0, 20, 20])
    assert np.array_equal(acc_b, acc_b2)                       # bit-reproducible
    for b in (0, 5, 11):
        acc_s, _, _, recs_s = _solve_host(lib, starts[b], goals[b], 10.0, 0.2, 0.8, [0, 0, 20, 20])
        assert np.array_equal(acc_s[0], acc_b[b])              # scenario result independent of the batch
        assert recs_s[0]["admm_iterations"] == recs_b[b]["admm_iterations"]


def test_failure_path_demo_inputs(torch_cuda, lib):
    """scp.py:846-865 demo: first avoidance QP is infeasible.  Reference: OSQP warning, loop continues with
    result.x (scp.py:446-449).  Here: the subproblem is counted in qp_unsolved, the loop continues, no exception."""
    p0 = np.array([[-2.0, -2], [0, -2], [2, -2]])
    pf = np.array([[2.0, 2], [0, 2], [-2, 2]])
    acc, pos, vel, recs = _solve_host(lib, p0, pf, 3.0, 0.2, 0.5, [-5, -5, 500, 200], max_admm_iter=3000)
    r = recs[0]
    assert r["status"] == 0 and not r["initial_feasible"]
    assert r["first_violation"][1:] == (0, 1) and r["first_violation_dist"] < 0.49
    assert r["qp_unsolved"] >= 1 and r["scp_iterations"] >= 1
    assert np.isfinite(acc).all() and r["min_separation"] < 0.5 - 0.01   # infeasible: separation not reached


def test_start_too_close_is_flagged(torch_cuda, lib):
    p0 = np.array([[5.0, 5.0], [5.3, 5.0]])
    pf = np.array([[15.0, 5.0], [15.0, 8.0]])
    _, _, _, recs = _solve_host(lib, p0, pf, 10.0, 0.2, 0.8, [0, 0, 20, 20], max_admm_iter=2000)
    assert recs[0]["status"] == 2 and recs[0]["first_violation"] == (0, 0, 1)


def test_c2_reference_outcomes_fixture(torch_cuda, lib):
    """The benchmark's own scenarios (config 2, seeds 10000..10063) against the outcomes of the VERBATIM reference at its
    own solver settings (tests/golden/c2_outcomes.npz, oracle/make_outcomes.py: OSQP defaults eps 1e-3, max_iter 10000,
    scp.py:360,442).  Unlike the n*.npz fixtures these are NOT filtered: 27 of the 64 contain a subproblem the
    reference ends as 'solved inaccurate' or at max_iter (scp.py:446-449: warning, iterate kept).

    Contract checked here:
      * same batch status ("success" on both sides for all 64, compute_trajectories_batch.py:50-54), finite
        trajectories, same pass on the dynamics-residual check (<= 1e-3);
      * minimum separation: the device result passes scp.py:610's R - 0.01 everywhere; the reference's eps-1e-3
        iterates sit up to its own primal tolerance (1e-3 (1 + 20 m)) below R, so its pass/fail is compared within
        that tolerance;
      * scenarios whose subproblems all solve on both sides: positions within 1e-2 relative of the LOOSE reference
        (median <= 2e-3) -- the loose reference is itself ~1e-3 from its own re-run (c2_outcomes_alt.npz) and
        3e-4..9e-4 from the certified minimiser the device path reproduces to 1e-12 (test_solve_matches_golden);
      * scenarios with an unsolved subproblem have no defined minimiser: trajectories are solver dependent on both
        sides, only bounded here (<= 5e-2 relative)."""
    from oracle import scp_oracle

    f = np.load(os.path.join(GOLDEN, "c2_outcomes.npz"))
    n, R, h, space = len(f["seed"]), 0.8, 0.2, [0, 0, 20, 20]
    acc, pos, vel, recs = _solve_host(lib, f["p0"], f["pf"], 10.0, h, R, space)
    z = np.zeros((25, 2))
    qs = f["qp_status"]
    ref_solved = np.array([(qs[b, : f["n_qp"][b]] == 1).all() for b in range(n)])
    gpu_solved = np.array([recs[b]["qp_unsolved"] == 0 for b in range(n)])
    slack = 1e-3 * (1 + 20.0)
    pe = np.empty(n)
    for b in range(n):
        assert (recs[b]["status"] == 0) == (str(f["status"][b]) == "success"), b
        assert np.isfinite(pos[b]).all() and np.isfinite(acc[b]).all() and bool(f["finite"][b])
        dyn = scp_oracle.dynamics_residual(acc[b], f["p0"][b], z, f["pf"][b], z, h, space, positions=pos[b])
        assert (dyn <= DYN_TOL) == bool(f["dyn_pass"][b]), (b, dyn)
        assert recs[b]["min_separation"] >= R - 0.01, (b, recs[b]["min_separation"])
        pe[b] = np.linalg.norm(pos[b] - f["positions"][b]) / np.linalg.norm(f["positions"][b])
    ref_sep_ok = f["min_separation"] >= R - 0.01 - slack
    assert ref_sep_ok.sum() >= n - 12, int(ref_sep_ok.sum())       # the rest: reference iterates beyond its own tolerance
    both = ref_solved & gpu_solved
    assert both.sum() >= 28 and (~ref_solved).sum() >= 20          # the fixture really contains both kinds
    assert pe[both].max() <= 1e-2 and np.median(pe[both]) <= 2e-3, (pe[both].max(), np.median(pe[both]))
    assert pe[~both].max() <= 5e-2, pe[~both].max()
    dscp = np.abs(np.array([recs[b]["scp_iterations"] for b in range(n)]) - f["scp_iterations"])
    assert dscp[both].max() <= 3
    assert abs(int((~gpu_solved).sum()) - int((~ref_solved).sum())) <= 8
    print(f"c2 outcomes: {int(both.sum())} scenarios solved on both sides, pos err median {np.median(pe[both]):.1e} max "
          f"{pe[both].max():.1e}; {int((~both).sum())} with an unsolved subproblem, pos err median {np.median(pe[~both]):.1e}")


# ------------------------------------------------------------------ BASELINE.json config sizes: properties
def test_c2_batch_properties(torch_cuda, lib):
    """128 of config 2's scenarios (25 agents, K=50, seeds 10000+b): size-independent properties --
    dynamics feasible, min separation passes wherever the loop converged, records consistent, and the device
    reconstruct/linearize kernels agree with the solver's own outputs."""
    from oracle import scp_oracle
    from path_planning import _capi
    from path_planning.scenarios.position_generator import generate_positions

    torch = torch_cuda
    B, N, R, h = 128, 25, 0.8, 0.2
    starts, goals = [], []
    for b in range(B):
        random.seed(10_000 + b)
        p0, pf = generate_positions(N, R)
        starts.append(p0)
        goals.append(pf)
    starts, goals = np.stack(starts), np.stack(goals)
    acc, pos, vel, recs = _solve_host(lib, starts, goals, 10.0, h, R, [0, 0, 20, 20])
    z = np.zeros((N, 2))
    n_ok = 0
    for b, r in enumerate(recs):
        assert r["status"] == 0
        dyn = scp_oracle.dynamics_residual(acc[b], starts[b], z, goals[b], z, h, [0, 0, 20, 20], positions=pos[b])
        assert dyn <= DYN_TOL, (b, dyn)
        assert abs(r["objective"] - (acc[b] ** 2).sum()) <= 1e-9 * r["objective"]
        assert abs(r["min_separation"] - scp_oracle.min_separation(pos[b])) <= 1e-12
        if r["converged"] and r["qp_unsolved"] == 0:
            assert r["min_separation"] >= R - 0.01, (b, r["min_separation"])
            n_ok += 1
    # ~40 % of these crowded scenarios contain a linearised subproblem that is primal INFEASIBLE (the reference
    # would get NaNs / 'maximum iterations reached' from OSQP there, scp.py:446-449); they are excluded above.
    assert n_ok >= 0.5 * B
    # device reconstruct on the solver's accelerations reproduces its positions / velocities
    d_acc, d_p0 = _dev(torch, acc), _dev(torch, starts)
    d_v0 = torch.zeros_like(d_p0)
    d_pos = torch.empty_like(d_acc)
    d_vel = torch.empty_like(d_acc)
    _capi.check(lib.scp_b200_reconstruct(d_acc.data_ptr(), d_p0.data_ptr(), d_v0.data_ptr(), B, N, 50, h,
                                         d_pos.data_ptr(), d_vel.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.abs(d_pos.cpu().numpy() - pos).max() <= 1e-10 and np.abs(d_vel.cpu().numpy() - vel).max() <= 1e-10
    _, _, minsep, _ = _linearize(torch, lib, pos, R, want_rows=False)
    assert np.abs(minsep - np.array([r["min_separation"] for r in recs])).max() <= 1e-12


def test_c3_sized_kernels_properties(torch_cuda, lib):
    """Config 3/4 sizes for the stand-alone kernels: 200 and 1000 agents x 100 steps.  Properties: every eta is a
    unit vector, eta.d == dist, antisymmetry under agent relabelling, and reconstruct is linear in the accelerations."""
    from path_planning import _capi

    torch = torch_cuda
    rng = np.random.default_rng(3)
    for N in (200, 1000):
        K = 100
        pos = rng.uniform(0, 100, (1, N, K, 2))
        eta, bound, minsep, _ = _linearize(torch, lib, pos, 0.8)
        assert np.abs(np.hypot(eta[..., 0], eta[..., 1]) - 1).max() <= 1e-14
        eta_r, _, minsep_r, _ = _linearize(torch, lib, pos[:, ::-1].copy(), 0.8)
        # relabel i -> N-1-i: row (i,j) becomes (N-1-j, N-1-i) with eta negated
        iu, ju = np.triu_indices(N, 1)
        idx = {}
        pidx = (N - 1 - ju) * (2 * N - (N - 1 - ju) - 1) // 2 + ((N - 1 - iu) - (N - 1 - ju) - 1)
        assert np.abs(eta_r[0][:, pidx] + eta[0]).max() <= 1e-14 and minsep[0] == minsep_r[0]
        del idx
    B, N, K, h = 2, 1000, 100, 0.2
    a1, a2 = rng.normal(size=(B, N, K, 2)), rng.normal(size=(B, N, K, 2))
    p0 = rng.uniform(0, 100, (B, N, 2))
    z = np.zeros_like(p0)

    def rec(a, p, v):
        d_pos = torch.empty((B, N, K, 2), dtype=torch.float64, device="cuda")
        d_vel = torch.empty_like(d_pos)
        _capi.check(lib.scp_b200_reconstruct(_dev(torch, a).data_ptr(), _dev(torch, p).data_ptr(),
                                             _dev(torch, v).data_ptr(), B, N, K, h, d_pos.data_ptr(), d_vel.data_ptr(), None))
        torch.cuda.synchronize()
        return d_pos.cpu().numpy(), d_vel.cpu().numpy()

    pa, va = rec(a1, p0, z)
    pb, vb = rec(a2, z, z)
    pc, vc = rec(a1 + 2 * a2, p0, z)
    assert np.abs(pc - (pa + 2 * pb)).max() <= 1e-9 and np.abs(vc - (va + 2 * vb)).max() <= 1e-10


def test_c1_default_cli_problem_properties(torch_cuda, lib):
    """BASELINE.json configs[0]: compute-trajectories defaults (10 agents, T=100 s, h=0.2 s -> K=500, 200 x 200 m).
    The oracle needs a 10000 x 10000 dense factorisation per rho update for this size, so only properties are
    checked: feasibility of the returned trajectory and agreement between the whole-grid and one-CTA kernels."""
    from oracle import scp_oracle
    from path_planning.scenarios.position_generator import generate_positions

    random.seed(0)
    p0, pf = generate_positions(10, 0.8)
    space = [0, 0, 200, 200]
    acc, pos, vel, recs = _solve_host(lib, p0, pf, 100, 0.2, 0.8, space)
    r = recs[0]
    assert pos.shape == (1, 10, 500, 2) and r["status"] == 0 and r["converged"]
    z = np.zeros((10, 2))
    assert scp_oracle.dynamics_residual(acc[0], p0, z, pf, z, 0.2, space, positions=pos[0]) <= DYN_TOL
    assert r["min_separation"] >= 0.8 - 0.01
    acc1, _, _, recs1 = _solve_host(lib, p0, pf, 100, 0.2, 0.8, space, team_mode=1)
    assert recs1[0]["scp_iterations"] == r["scp_iterations"] and np.abs(acc1 - acc).max() <= 1e-9


def test_batch_cli_records_and_files(torch_cuda, lib, tmp_path):
    """compute-trajectories-batch (reference cli/compute_trajectories_batch.py:70-173): JSON/CSV schema and records."""
    import csv
    import json

    from path_planning.cli import compute_trajectories_batch as ctb

    cfg = dict(Ns=[5, 6], trials_per_N=3, results_dir=str(tmp_path), rng_seed=7)
    res = ctb.main(cfg)
    assert set(res) == {"meta", "runs", "summary"} and res["meta"]["schema_version"] == "1.0"
    assert len(res["runs"]) == 6 and set(res["summary"]) == {"5", "6"}
    for run in res["runs"]:
        assert {"N", "status", "time_sec", "error", "K", "T", "h", "trial_index"} <= set(run)
        assert run["status"] == "success" and run["K"] == 50 and run["time_sec"] > 0
    files = sorted(p.name for p in tmp_path.iterdir())
    assert len(files) == 2 and files[0].endswith(".csv") and files[1].endswith(".json")
    rows = list(csv.DictReader(open(tmp_path / files[0])))
    assert list(rows[0].keys()) == ["N", "trial_index", "status", "time_sec", "K", "T", "h", "error"] and len(rows) == 6
    assert json.load(open(tmp_path / files[1]))["summary"]["5"]["count"] == 3
    # the unbatched path (run_single_trial, reference :28-67) gives the same record keys
    single = ctb.run_single_trial(5, {**ctb.CONFIG, **cfg}, rng=np.random)
    assert {"N", "status", "time_sec", "error", "K", "T", "h"} <= set(single) and single["status"] == "success"


def test_linearize_range_shards_concatenate_to_full(torch_cuda, lib):
    """Agent-sharded linearisation (config 4's pairwise step): the per-rank pair ranges of scp_b200_linearize_range,
    run one after the other on one GPU, reproduce the full kernel's rows bit for bit and its reductions via MIN."""
    from path_planning import _capi
    from path_planning.solvers.sharded import agent_blocks, pair_index

    torch = torch_cuda
    rng = np.random.default_rng(11)
    N, K, R = 300, 40, 0.8
    pos = rng.uniform(0, 60, (1, N, K, 2))
    eta_full, bound_full, minsep_full, first_full = _linearize(torch, lib, pos, R)
    P = N * (N - 1) // 2
    d_pos = _dev(torch, pos)
    for world in (2, 8):
        b = agent_blocks(N, world)
        etas, bounds, mins, firsts = [], [], [], []
        for g in range(world):
            pb = pair_index(b[g], N)
            pe = pair_index(b[g + 1], N) if b[g + 1] < N else P
            rows = pe - pb
            eta = torch.full((K, max(rows, 1), 2), float("nan"), dtype=torch.float64, device="cuda")
            bound = torch.full((K, max(rows, 1)), float("nan"), dtype=torch.float64, device="cuda")
            ms = torch.zeros(1, dtype=torch.float64, device="cuda")
            fi = torch.zeros(3, dtype=torch.int32, device="cuda")
            _capi.check(lib.scp_b200_linearize_range(d_pos.data_ptr(), 1, N, K, R, 0.01, pb, pe, eta.data_ptr(),
                                                     bound.data_ptr(), ms.data_ptr(), fi.data_ptr(), None))
            torch.cuda.synchronize()
            etas.append(eta.cpu().numpy()[:, :rows]); bounds.append(bound.cpu().numpy()[:, :rows])
            mins.append(float(ms[0])); firsts.append(tuple(int(v) for v in fi.tolist()))
        assert np.array_equal(np.concatenate(etas, axis=1), eta_full[0])
        assert np.array_equal(np.concatenate(bounds, axis=1), bound_full[0])
        assert min(mins) == minsep_full[0]
        rows_idx = [k * P + pair_index(i, N) + (j - i - 1) for (k, i, j) in firsts if k >= 0]
        exp = tuple(first_full[0])
        assert (min(rows_idx) if rows_idx else None) == (exp[0] * P + pair_index(exp[1], N) + exp[2] - exp[1] - 1 if exp[0] >= 0 else None)


@pytest.mark.parametrize("name", ["c1_outcomes.npz", "n50_outcomes.npz"])
def test_reference_outcomes_other_sizes(torch_cuda, lib, name):
    """Config-1 size (10 agents, K=500, compute-trajectories defaults, random.seed(0)) and 50 agents (config-5 size, reference
    generator): the device path against the VERBATIM reference at its own settings (oracle/make_outcomes.py with
    SCP_OUTCOME_CFG).  A certified golden at these sizes did not finish in hours of CPU time, so the comparison is with the
    eps-1e-3 reference and the tolerance is its solver noise: same status, finite, dynamics-feasible, separation within the
    reference's primal tolerance, positions within 2e-2 relative (5e-2 when a subproblem is unsolved on either side)."""
    from oracle import scp_oracle

    path = os.path.join(GOLDEN, name)
    if not os.path.isfile(path):
        pytest.skip(f"{name} not generated (hours of CPU time in the build container)")
    f = np.load(path)
    N, T, h, R = int(f["config"][0]), float(f["config"][1]), float(f["config"][2]), float(f["config"][3])
    space = [float(x) for x in f["config"][4:8]]
    acc, pos, vel, recs = _solve_host(lib, f["p0"], f["pf"], T, h, R, space)
    z = np.zeros((N, 2))
    for b in range(len(f["seed"])):
        assert (recs[b]["status"] == 0) == (str(f["status"][b]) == "success")
        assert np.isfinite(pos[b]).all() and bool(f["finite"][b])
        dyn = scp_oracle.dynamics_residual(acc[b], f["p0"][b], z, f["pf"][b], z, h, space, positions=pos[b])
        assert (dyn <= DYN_TOL) == bool(f["dyn_pass"][b])
        # separation: same pass/fail, the reference's judged within its own primal tolerance.  (50 agents in the 20 x 20 m
        # arena, seed 10001: a subproblem ends at max_iter on the reference side and unsolved here; BOTH trajectories end
        # below R - 0.01 -- 0.767 m and 0.764 m -- which is the same outcome.)
        ref_sep_ok = float(f["min_separation"][b]) >= R - 0.01 - 1e-3 * (1 + max(space[2], space[3]))
        assert (recs[b]["min_separation"] >= R - 0.01) == ref_sep_ok, (b, recs[b]["min_separation"], float(f["min_separation"][b]))
        pe = np.linalg.norm(pos[b] - f["positions"][b]) / np.linalg.norm(f["positions"][b])
        solved = recs[b]["qp_unsolved"] == 0 and (f["qp_status"][b, : f["n_qp"][b]] == 1).all()
        assert pe <= (2e-2 if solved else 5e-2), (b, pe)
        print(f"{name} seed {int(f['seed'][b])}: SCP iterations {recs[b]['scp_iterations']} (reference {int(f['scp_iterations'][b])}), "
              f"positions {pe:.1e} from the eps-1e-3 reference, solved on both sides: {bool(solved)}")
