"""Post-solve analysis on the device (`scp_b200_check_batch`, include/scp_b200.h): minimum separation at the samples --
the quantity `_fast_check_avoidance_constraints` tests (reference scp.py:597-615) and `print_distance_analysis` reports
(reference scenarios/position_generator.py:173-205) --, minimum separation in continuous time (between samples the
reference's dynamics are constant-acceleration segments, scp.py:371-397, so the check is exact, not a finer sampling)
and the dynamics residual of SURVEY.md 8(c).  PyTorch is plumbing only; there is no CPU path."""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

FIELDS = [n for n, _ in _capi.Check._fields_]


def check_trajectories(trajectories, initial_positions, final_positions, time_step, space_dims=None,
                       initial_velocities=None, final_velocities=None, vel_limit=2.0, acc_limit=15.0, jerk_limit=20.0,
                       min_distance=None, device=None):
    """trajectories: the result dict of SCP.generate_trajectories (arrays (N,K,2)) or a batch of them ((B,N,K,2));
    numpy arrays or CUDA tensors.  Returns one dict per scenario with the fields of `scp_b200_check` plus, when
    `min_distance` is given, the pass/fail flags of the reference's threshold R - 0.01 (scp.py:610)."""
    import torch

    if not torch.cuda.is_available():
        raise _capi.ScpB200Error("no CUDA device: the analysis kernels have no CPU path")
    lib = _capi.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def up(a, shape=None):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        t = t.to(dev, dtype=torch.float64).contiguous()
        return t.reshape(shape) if shape is not None else t

    pos = up(trajectories["positions"])
    if pos.dim() == 3:
        pos = pos[None]
    B, N, K, _ = pos.shape
    acc, vel = up(trajectories["accelerations"], (B, N, K, 2)), up(trajectories["velocities"], (B, N, K, 2))
    p0, pf = up(initial_positions, (B, N, 2)), up(final_positions, (B, N, 2))
    zero = torch.zeros_like(p0)
    v0 = zero if initial_velocities is None else up(initial_velocities, (B, N, 2))
    vf = zero if final_velocities is None else up(final_velocities, (B, N, 2))
    prob = _capi.default_problem(N, K * time_step, time_step, min_distance or 0.0, space_dims, lib=lib)
    prob.n_steps = K
    prob.vel_limit, prob.acc_limit, prob.jerk_limit = float(vel_limit), float(acc_limit), float(jerk_limit)
    out = torch.empty(B * C.sizeof(_capi.Check), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _capi.check(lib.scp_b200_check_batch(C.byref(prob), B, acc.data_ptr(), pos.data_ptr(), vel.data_ptr(), p0.data_ptr(),
                                             v0.data_ptr(), pf.data_ptr(), vf.data_ptr(), out.data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream))
    vals = np.frombuffer(out.cpu().numpy().tobytes(), dtype=np.float64).reshape(B, len(FIELDS))
    res = []
    for b in range(B):
        d = dict(zip(FIELDS, (float(x) for x in vals[b])))
        d["dynamics_pass"] = d["dynamics_residual"] <= 1e-3
        if min_distance is not None:
            d["min_separation_pass"] = d["min_separation"] >= min_distance - 0.01
            d["min_separation_continuous_pass"] = d["min_separation_continuous"] >= min_distance - 0.01
        res.append(d)
    return res
