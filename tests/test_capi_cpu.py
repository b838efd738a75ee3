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
