"""Batched scenario generation on the device (`scp_b200_generate_scenarios`, include/scp_b200.h): thousands of
start/goal sets by rejection sampling, one CTA per scenario, without the O(N^2) Python loop per scenario of the host
generators (reference scenarios/position_generator.py:44-75, 247-248).  Same layouts and acceptance rules as
`generate_positions` (layout "reference") and `generate_positions_large` (layout "large"); the random stream is a
counter-based hash keyed by (seed, scenario index), NOT Python's `random`: use the host generators when the reference's
exact draws are needed (the benchmark's seeded workloads do)."""

from __future__ import annotations

import ctypes as C
import math

from .. import _capi

LAYOUTS = {"reference": 0, "large": 1}


def space_dims_for(layout, n_vehicles):
    """Arena of a layout: the reference's 20 x 20 m box, or side sqrt(16 N) for the bounded-travel layout."""
    if LAYOUTS[layout] == 0:
        return [0.0, 0.0, 20.0, 20.0]
    side = math.sqrt(16.0 * n_vehicles)
    return [0.0, 0.0, side, side]


def generate_scenarios_device(n_scenarios, n_vehicles, min_distance=0.4, layout="reference", seed=0, first_scenario=0,
                              time_horizon=10.0, vel_limit=2.0, max_attempts=0, device=None):
    """Returns (initial (B,N,2), final (B,N,2), ok (B,) bool) as CUDA tensors, asynchronously on the current stream.
    ok[b] is False where the attempts ran out (the reference raises ValueError there, position_generator.py:58-59)."""
    import torch

    if not torch.cuda.is_available():
        raise _capi.ScpB200Error("no CUDA device: the device generator has no CPU path")
    lib = _capi.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    p0 = torch.empty((n_scenarios, n_vehicles, 2), dtype=torch.float64, device=dev)
    pf = torch.empty_like(p0)
    st = torch.zeros(n_scenarios, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(lib.scp_b200_generate_scenarios(
            int(n_scenarios), int(n_vehicles), LAYOUTS[layout], float(min_distance), float(time_horizon), float(vel_limit),
            C.c_uint64(int(seed) & (2 ** 64 - 1)), int(first_scenario), int(max_attempts), p0.data_ptr(), pf.data_ptr(),
            st.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return p0, pf, st.bool()
