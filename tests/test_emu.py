"""CPU check of the KERNEL SOURCE logic: csrc/scp_device.inl compiled as plain C++ (tests/emu, test-only)
against the certified golden fixtures.  The same source runs on the GPU in test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import golden_cases, golden_limits

emu = pytest.importorskip("emu_binding")


@pytest.mark.parametrize("path", golden_cases(max_agents=8))
def test_emulated_kernel_matches_golden(path):
    g = np.load(path)
    tr, recs = emu.solve(g["p0"], g["pf"], float(g["T"]), float(g["h"]), float(g["R"]), list(g["space"]), **golden_limits(g))
    r = recs[0]
    assert r["status"] == 0 and r["scp_iterations"] == int(g["iterations"])
    assert np.allclose(r["rel_steps"], g["rel_steps"], rtol=5e-3, atol=1e-4)
    perr = np.linalg.norm(tr["positions"][0] - g["positions"]) / np.linalg.norm(g["positions"])
    assert perr <= 1e-3
    assert abs(r["objective"] - float(g["objective"])) <= 1e-4 * float(g["objective"])


def test_emulated_single_agent_and_two_far_agents():
    tr, recs = emu.solve(np.array([[2.0, 2.0]]), np.array([[8.0, 9.0]]), 4.0, 0.2, 0.8, [0, 0, 20, 20])
    assert recs[0]["status"] == 0 and recs[0]["scp_iterations"] == 0 and recs[0]["initial_feasible"]
    # closed form: min-norm rest-to-rest move, objective = 12 d^2 / (T^3) * h-discretisation ~ check terminal state
    a = tr["accelerations"][0, 0]
    K, h = a.shape[0], 0.2
    c1 = np.cumsum(a, 0)
    c2 = np.cumsum(c1, 0)
    assert np.allclose(h * c1[-1], 0, atol=1e-6)
    assert np.allclose(np.array([2.0, 2.0]) + h * h * (c2[-1] - 0.5 * c1[-1]), [8.0, 9.0], atol=1e-6)
