"""No-GPU checks of the boundary: the library loads, exports every symbol include/*.h declares,
struct layouts agree, argument errors are reported, and nothing computes on the CPU."""
import ctypes as C
import glob
import os
import re

import pytest

from conftest import ROOT
from path_planning import _capi


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        names |= set(re.findall(r"\b(scp_b200_[a-z_0-9]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_capi.exported_symbols())


def test_struct_layouts_and_defaults():
    lib = _capi.load()
    assert lib.scp_b200_sizeof_problem() == C.sizeof(_capi.Problem)
    assert lib.scp_b200_sizeof_record() == C.sizeof(_capi.Record)
    p = _capi.default_problem(25, 10.0, 0.2, 0.8, [0, 0, 20, 20])
    assert (p.n_agents, p.n_steps) == (25, 50)
    # K = int(T/h) with the reference's float truncation (scp.py:43)
    for T, h in ((3.0, 0.1), (10.0, 0.2), (100, 0.2), (20.0, 0.2)):
        assert _capi.default_problem(2, T, h, 0.5).n_steps == int(T / h)
    assert (p.vel_limit, p.acc_limit, p.jerk_limit, p.scp_tolerance, p.max_scp_iter) == (2.0, 15.0, 20.0, 1.5e-2, 15)
    assert lib.scp_b200_workspace_bytes(C.byref(p), 2) > lib.scp_b200_workspace_bytes(C.byref(p), 1) > 0
    assert lib.scp_b200_tables_bytes(C.byref(p)) == (2 * 50 * 50 + 5 * 50) * 8


def test_argument_errors_are_reported_without_a_gpu():
    lib = _capi.load()
    p = _capi.default_problem(25, 10.0, 0.2, 0.8)
    p.n_steps = 1
    rc = lib.scp_b200_solve_batch(C.byref(p), 4, None, None, None, None, None, None, 0, 1, None, None, None, None, None)
    assert rc != 0 and b"n_steps" in lib.scp_b200_last_error()
    with pytest.raises(_capi.ScpB200Error):
        _capi.check(rc)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from path_planning.solvers.batch import BatchSolver

    with pytest.raises(_capi.ScpB200Error):
        BatchSolver(5, 10.0, 0.2, 0.8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ba-path-planning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".inl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
                assert "scp_emu" not in src


def test_stream_default_problem_and_engine_rule():
    """Host-only pieces of the streaming solver: its default settings (fixed rho scaled with the horizon, relaxation,
    lazy rows) and the rule that sends single large scenarios to it."""
    from path_planning.solvers.scp import use_stream_engine

    p = _capi.default_problem(25, 10.0, 0.2, 0.8, [0, 0, 20, 20], stream=True)          # K = 50
    assert (p.n_agents, p.n_steps, p.polish, p.adapt_every, p.relax_pct, p.lazy_rows) == (25, 50, 0, 0, 160, 1)
    assert (p.eps_abs, p.eps_rel, p.max_admm_iter, p.check_every, p.stall_window) == (1e-4, 1e-4, 20000, 50, 1000)
    assert p.rho0 == 1.0 and p.momentum_pct == 0 and p.warm_duals == 0
    assert _capi.default_problem(200, 20.0, 0.2, 0.8, stream=True).rho0 == pytest.approx(0.25)    # K = 100
    assert _capi.default_problem(10, 100.0, 0.2, 0.8, stream=True).rho0 == pytest.approx(0.1)     # K = 500: clamped
    # the one-CTA defaults are untouched by the streaming ones
    q = _capi.default_problem(25, 10.0, 0.2, 0.8)
    assert (q.polish, q.adapt_every, q.relax_pct, q.max_admm_iter, q.check_every) == (1, 100, 0, 5000, 25)
    # engine rule: one (or a few) large scenarios with K <= 128 -> streaming; batches, K = 500 and small ones -> one CTA
    assert use_stream_engine("auto", 200, 100) and use_stream_engine("auto", 1000, 100, 1)
    assert not use_stream_engine("auto", 200, 100, 1024) and not use_stream_engine("auto", 10, 500)
    assert not use_stream_engine("auto", 25, 50) and use_stream_engine("stream", 5, 50) and not use_stream_engine("cta", 1000, 100)


def test_stream_create_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call: bad sizes are reported through the error string."""
    lib = _capi.load()
    p = _capi.default_problem(5, 10.0, 0.2, 0.8, stream=True)
    h = C.c_void_p()
    p.n_steps = 200
    assert lib.scp_b200_stream_create(C.byref(p), 1, 16, 0, 1, None, C.byref(h)) != 0
    assert b"n_steps" in lib.scp_b200_last_error()
    p.n_steps = 50
    assert lib.scp_b200_stream_create(C.byref(p), 2, 16, 0, 2, None, C.byref(h)) != 0      # sharding solves ONE scenario
    assert b"ONE scenario" in lib.scp_b200_last_error()
    assert lib.scp_b200_stream_create(C.byref(p), 1, 16, 3, 2, None, C.byref(h)) != 0      # rank outside world
    assert h.value is None


def test_analysis_and_device_generator_have_no_cpu_path():
    """The steps either side of the solve (scp_b200_check_batch, scp_b200_generate_scenarios) are device kernels:
    without a CUDA device the host wrappers raise, they do not fall back to numpy."""
    import numpy as np
    import torch

    from path_planning import _capi
    from path_planning.analysis import FIELDS, check_trajectories
    from path_planning.scenarios.device_generator import generate_scenarios_device, space_dims_for

    assert C.sizeof(_capi.Check) == 8 * len(FIELDS) == 64
    assert space_dims_for("reference", 25) == [0.0, 0.0, 20.0, 20.0] and abs(space_dims_for("large", 100)[2] - 40.0) < 1e-12
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    z = np.zeros((2, 5, 2))
    with pytest.raises(_capi.ScpB200Error):
        check_trajectories({"positions": z, "velocities": z, "accelerations": z}, z[:, 0], z[:, 0], 0.2)
    with pytest.raises(_capi.ScpB200Error):
        generate_scenarios_device(4, 10, 0.8)
    # argument errors are reported through the C ABI without touching the GPU
    lib = _capi.load()
    assert lib.scp_b200_generate_scenarios(1, 0, 0, 0.8, 10.0, 2.0, 0, 0, 0, None, None, None, None) != 0
    assert lib.scp_b200_check_batch(None, 1, None, None, None, None, None, None, None, None, None) != 0
