"""TEST-ONLY: ctypes access to tests/emu/libscp_emu.so (host emulation of the kernel source)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, os.path.join(_ROOT, "ba-path-planning_b200"))
from path_planning import _capi  # noqa: E402  (struct layouts only; the CUDA library is not loaded)


def build():
    so = os.path.join(_HERE, "libscp_emu.so")
    srcs = [os.path.join(_HERE, "scp_emu.cpp"),
            os.path.join(_ROOT, "ba-path-planning_b200", "csrc", "scp_device.inl"),
            os.path.join(_ROOT, "ba-path-planning_b200", "csrc", "scp_tables.h"),
            os.path.join(_ROOT, "ba-path-planning_b200", "csrc", "scp_defaults.h"),
            os.path.join(_ROOT, "include", "scp_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, srcs[0]])
    return so


def load():
    lib = C.CDLL(build())
    lib.scp_emu_default_problem.argtypes = [C.POINTER(_capi.Problem), C.c_int, C.c_double, C.c_double, C.c_double]
    lib.scp_emu_solve_batch.argtypes = [C.POINTER(_capi.Problem), C.c_int] + [C.c_void_p] * 8 + [C.c_int]
    lib.scp_emu_solve_batch.restype = C.c_int
    return lib


def solve(p0, pf, T, h, R, space, v0=None, vf=None, nthreads=64, **overrides):
    lib = load()
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    if p0.ndim == 2:
        p0 = p0[None]
    pf = np.ascontiguousarray(pf, dtype=np.float64).reshape(p0.shape)
    B, N, _ = p0.shape
    v0 = np.zeros_like(p0) if v0 is None else np.ascontiguousarray(v0, dtype=np.float64).reshape(p0.shape)
    vf = np.zeros_like(p0) if vf is None else np.ascontiguousarray(vf, dtype=np.float64).reshape(p0.shape)
    prob = _capi.Problem()
    lib.scp_emu_default_problem(C.byref(prob), N, T, h, R)
    for i in range(4):
        prob.space[i] = float(space[i])
    for k, v in overrides.items():
        setattr(prob, k, v)
    K = prob.n_steps
    acc = np.zeros((B, N, K, 2)); pos = np.zeros((B, N, K, 2)); vel = np.zeros((B, N, K, 2))
    recs = (_capi.Record * B)()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.scp_emu_solve_batch(C.byref(prob), B, ptr(p0), ptr(v0), ptr(pf), ptr(vf), ptr(acc), ptr(pos), ptr(vel),
                                 C.cast(recs, C.c_void_p), nthreads)
    assert rc == 0
    return dict(accelerations=acc, positions=pos, velocities=vel), [_capi.record_to_dict(r) for r in recs]
