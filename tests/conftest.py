import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ba-path-planning_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def golden_cases(max_agents=None):
    import glob

    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "n*_s*.npz"))):
        n = int(os.path.basename(f).split("_")[0][1:])
        if max_agents is None or n <= max_agents:
            out.append(f)
    return out


@pytest.fixture
def truth_mode():
    """Oracle QP back-end in truth mode: ADMM 1e-5 + active-set refinement + KKT certificate."""
    from oracle import scp_oracle

    osqp = scp_oracle._osqp()
    osqp.OVERRIDES.clear()
    osqp.OVERRIDES.update(eps_abs=1e-5, eps_rel=1e-5, max_iter=200000, certify=True)
    osqp.STATS.clear()
    yield osqp
    osqp.OVERRIDES.clear()
