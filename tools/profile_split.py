#!/usr/bin/env python
"""Diagnostic (library built with `make -C ba-path-planning_b200/csrc PROFILE_SPLIT=1`): where the ADMM cycles of the
one-CTA solver go -- fused iterations, their collision rows, check iterations -- on config-2 scenarios."""
import ctypes as C
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))
from path_planning import _capi  # noqa: E402
from path_planning.scenarios.position_generator import generate_positions, generate_positions_large  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
B = int(sys.argv[2]) if len(sys.argv) > 2 else 296
lib = _capi.load()
p0 = np.empty((B, N, 2)); pf = np.empty((B, N, 2)); space = [0, 0, 20, 20]
for b in range(B):
    random.seed(10_000 + b)
    if N <= 25:
        p0[b], pf[b] = generate_positions(N, 0.8)
    else:
        p0[b], pf[b], space = generate_positions_large(N, 0.8, time_horizon=10.0)
prob = _capi.default_problem(N, 10.0, 0.2, 0.8, space)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    setattr(prob, k, float(v) if ("." in v or "e" in v) else int(v))
K = 50
z = np.zeros_like(p0)
acc = np.empty((B, N, K, 2)); pos = np.empty((B, N, K, 2)); vel = np.empty((B, N, K, 2))
rec = (_capi.Record * B)()
ptr = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
for _ in range(2):
    _capi.check(lib.scp_b200_solve_batch_host(C.byref(prob), B, ptr(p0), ptr(z), ptr(pf), ptr(z), ptr(acc), ptr(pos), ptr(vel),
                                              C.cast(rec, C.c_void_p), 0))
it = sum(r.admm_iterations for r in rec)
tot = sum(r.cycles_total for r in rec); admm = sum(r.cycles_admm for r in rec); pol = sum(r.cycles_polish for r in rec)
fused = sum(r.rel_step[29] for r in rec); colx = sum(r.rel_step[30] for r in rec); chk = sum(r.rel_step[31] for r in rec)
n_chk = it / 25.0
print(f"N={N} B={B}: ADMM iterations {it}, cycles total {tot:.3g}: admm {admm/tot:.3f} polish {pol/tot:.3f}")
print(f"  per ADMM iteration (all): {admm/it:.0f} cycles")
print(f"  fused iterations: {fused/(it-n_chk):.0f} cycles each ({fused/admm:.3f} of ADMM cycles)")
print(f"  collision rows of fused iterations: {colx/(it-n_chk):.0f} cycles each ({colx/admm:.3f})")
print(f"  check iterations (whole pass incl. residuals, certificates, signature): {chk/max(n_chk,1):.0f} cycles each ({chk/admm:.3f})")
print(f"  rest of ADMM time (operator factorisations, setup): {(admm-fused-colx-chk)/admm:.3f}")
