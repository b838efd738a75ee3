"""CPU tests of the oracle: SURVEY.md pins T1-T7 and the committed golden fixtures."""
import random

import numpy as np
import pytest

from conftest import golden_cases, golden_limits
from oracle import scenarios, scp_oracle


def _oracle(N=4, T=2.0, h=0.2, R=0.8, seed=5):
    rng = np.random.default_rng(seed)
    o = scp_oracle.ScpOracle(N, T, h, R, [0, 0, 20, 20])
    o.set_initial_states(rng.uniform(2, 18, (N, 2)), rng.uniform(-0.5, 0.5, (N, 2)))
    o.set_final_states(rng.uniform(2, 18, (N, 2)), rng.uniform(-0.5, 0.5, (N, 2)))
    o.precompute_constraint_matrices()
    return o, rng


def test_T1_shapes_and_nnz():
    o, _ = _oracle(N=3, T=2.4)
    N, K = o.N, o.K
    assert o.C_acc.shape == (2 * N * K, 2 * N * K) and o.C_acc.nnz == 2 * N * K
    assert o.C_jerk.shape == (2 * N * (K - 1), 2 * N * K) and o.C_jerk.nnz == 4 * N * (K - 1)
    assert o.C_vel.nnz == N * K * (K + 1) and o.C_pos.nnz == N * K * (K + 1)


def test_T2_scan_form_and_T6_state_map():
    o, rng = _oracle()
    N, K, h = o.N, o.K, o.h
    a = rng.normal(size=(N, K, 2))
    pos, vel = o.states_from_accelerations(a)
    v0 = o.initial_velocities.reshape(N, 2)
    p0 = o.initial_positions.reshape(N, 2)
    rv = (o.C_vel @ a.reshape(-1)).reshape(N, K, 2)
    rp = (o.C_pos @ a.reshape(-1)).reshape(N, K, 2)
    kk = np.arange(1, K)[None, :, None]
    assert np.allclose(vel[:, 1:], v0[:, None] + rv[:, :-1], atol=1e-13)
    assert np.allclose(pos[:, 1:], p0[:, None] + h * kk * v0[:, None] + rp[:, :-1], atol=1e-13)
    # explicit sums of scp.py:386-395
    i, k = 1, K - 1
    vk = v0[i] + h * a[i, :k].sum(0)
    pk = p0[i] + h * k * v0[i] + sum(h**2 * (k - j - 0.5) * a[i, j] for j in range(k))
    assert np.allclose(vel[i, k], vk, atol=1e-13) and np.allclose(pos[i, k], pk, atol=1e-13)


def test_T3_collision_rows_mean_eta_dot_separation():
    o, rng = _oracle(N=4)
    N, K = o.N, o.K
    a_prev = rng.normal(size=2 * N * K)
    A, l, u = o.collision_rows(a_prev)
    a = rng.normal(size=2 * N * K)
    pos, _ = o.states_from_accelerations(a)
    iu, ju = np.triu_indices(N, 1)
    eta = o._last_eta  # (K,P,2)
    lhs = (A @ a - l).reshape(K, -1)
    rhs = np.einsum("kpa,pka->kp", eta, pos[iu] - pos[ju]) - o.R
    assert np.allclose(lhs, rhs, atol=1e-11)
    assert np.all(np.isinf(u)) and A.nnz == 4 * iu.size * K * (K - 1) // 2


def test_T4_exactly_4N_equalities():
    o, _ = _oracle()
    _, l, u = o._stack_dynamics()
    assert int(np.sum(l == u)) == 4 * o.N


def test_generator_matches_golden_inputs():
    for f in golden_cases():
        g = np.load(f)
        if int(g["seed"]) < 0:
            continue                     # explicit start/goal positions (binding box-row cases), not generated
        random.seed(int(g["seed"]))
        p0, pf = scenarios.generate_positions(int(g["N"]), float(g["R"]))
        assert np.array_equal(p0, g["p0"]) and np.array_equal(pf, g["pf"])


def test_T7_end_to_end_seed0(truth_mode):
    random.seed(0)
    p0, pf = scenarios.generate_positions(5, 0.8)
    o = scp_oracle.ScpOracle(5, 10.0, 0.2, 0.8, [0, 0, 20, 20])
    o.set_initial_states(p0)
    o.set_final_states(pf)
    tr = o.generate_trajectories()
    assert o.record["iterations"] == 3
    assert np.allclose(o.record["rel_steps"], [0.9738, 0.6548, 0.00841], atol=2e-4)
    assert abs((tr["accelerations"] ** 2).sum() - 30.120795) < 1e-5
    assert abs(scp_oracle.min_separation(tr["positions"]) - 0.80255) < 1e-4
    assert all(q["cert"] <= 1e-9 for q in o.record["qp"])
    g = [f for f in golden_cases() if f.endswith("n5_s0.npz")]
    if g:
        gold = np.load(g[0])
        assert np.abs(tr["positions"] - gold["positions"]).max() < 1e-8


@pytest.mark.parametrize("path", golden_cases(max_agents=10))
def test_oracle_reproduces_golden(path, truth_mode):
    g = np.load(path)
    N = int(g["N"])
    o = scp_oracle.ScpOracle(N, float(g["T"]), float(g["h"]), float(g["R"]), list(g["space"]))
    for key, val in golden_limits(g).items():            # box limits are plain attributes, as in the reference (scp.py:67-74)
        name = key.split("_")[0]
        setattr(o, name + "_min", -val)
        setattr(o, name + "_max", val)
    o.set_initial_states(g["p0"])
    o.set_final_states(g["pf"])
    tr = o.generate_trajectories()
    assert o.record["iterations"] == int(g["iterations"])
    assert np.abs(tr["positions"] - g["positions"]).max() < 1e-8
    assert np.abs(tr["accelerations"] - g["accelerations"]).max() < 1e-8


def test_failure_path_demo_inputs_first_qp_infeasible():
    """scp.py:846-865 demo inputs: the first linearised avoidance QP is primal infeasible
    (SURVEY.md section 8c); the shim reports it and the loop carries on like the reference."""
    from oracle import scp_oracle as so

    osqp = so._osqp()
    osqp.OVERRIDES.clear()
    osqp.STATS.clear()
    o = so.ScpOracle(3, 3.0, 0.2, 0.5, [-5, -5, 500, 200])
    o.set_initial_states(np.array([[-2.0, -2], [0, -2], [2, -2]]))
    o.set_final_states(np.array([[2.0, 2], [0, 2], [-2, 2]]))
    o.precompute_constraint_matrices()
    a0 = o.solve_initial_trajectory()
    assert abs((a0**2).sum() - 178.5714) < 1e-2
    pos, _ = o.states_from_accelerations(a0)
    assert not o.fast_check_avoidance(pos)
    k, i, j, d = o.record["first_violation"]
    assert (i, j) == (0, 1) and d < 0.49
    o.solve_with_avoidance(a0)
    assert osqp.STATS[-1]["status"] in (-3, 3, -2)


def test_dynamics_residual_on_golden():
    for f in golden_cases():
        g = np.load(f)
        N = int(g["N"])
        z = np.zeros((N, 2))
        r = scp_oracle.dynamics_residual(g["accelerations"], g["p0"], z, g["pf"], z, float(g["h"]), list(g["space"]),
                                         positions=g["positions"])
        assert r <= 1e-8
        assert scp_oracle.min_separation(g["positions"]) >= float(g["R"]) - 0.01


def test_golden_set_covers_binding_box_rows():
    """per box class (jerk, acc, vel, pos) there is a fixture in which that class binds at the
    KKT-certified optimum; every fixture carries certificates <= 1e-9."""
    import os

    from conftest import active_box_classes, golden_cases

    active = {}
    for p in golden_cases():
        g = np.load(p)
        assert float(np.max(g["certs"])) <= 1e-9, p
        for cls in active_box_classes(g, tol=1e-6):
            active.setdefault(cls, []).append(os.path.basename(p))
    assert {"jerk", "acc", "vel", "pos"} <= set(active), active
