"""Batched device solve: B independent scenarios in one kernel launch.

This is where the scenario-level data parallelism of the reference's batch
driver (cli/compute_trajectories_batch.py:103-112, a sequential loop there) is
introduced.  PyTorch is plumbing only: it owns the device buffers and the
stream; all arithmetic happens in libscp_b200.so (include/scp_b200.h).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _capi


def _require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise _capi.ScpB200Error("no CUDA device: the SCP solver has no CPU path")
    return torch


class BatchSolver:
    """Reusable solver for one problem definition (N, K, h, R, limits, settings)."""

    def __init__(self, n_vehicles, time_horizon, time_step, min_distance, space_dims=None, device=None,
                 **settings):
        torch = _require_cuda()
        self.lib = _capi.load()
        self.problem = _capi.default_problem(n_vehicles, time_horizon, time_step, min_distance, space_dims)
        for k, v in settings.items():
            if not hasattr(self.problem, k):
                raise TypeError(f"unknown solver setting {k!r}")
            setattr(self.problem, k, v)
        self.N, self.K = int(self.problem.n_agents), int(self.problem.n_steps)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            tb = self.lib.scp_b200_tables_bytes(C.byref(self.problem))
            self.tables = torch.empty(tb, dtype=torch.uint8, device=self.device)
            _capi.check(self.lib.scp_b200_build_tables(C.byref(self.problem), self.tables.data_ptr(),
                                                       torch.cuda.current_stream(self.device).cuda_stream))
            self.max_slots = int(self.lib.scp_b200_default_slots(C.byref(self.problem)))
        self._ws = {}          # lane -> (workspace tensor, slots)

    def _workspace(self, slots, lane=0):
        import torch

        ws, have = self._ws.get(lane, (None, 0))
        if ws is None or slots > have:
            nbytes = self.lib.scp_b200_workspace_bytes(C.byref(self.problem), slots)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[lane] = (ws, slots)
        return ws

    def solve_device(self, p0, pf, v0=None, vf=None, lane=0):
        """p0, pf (and optional v0, vf): float64 CUDA tensors (B,N,2).  Returns device tensors
        acc,pos,vel (B,N,K,2) and a uint8 tensor holding B result records; asynchronous on
        the current stream.  `lane` selects the workspace: batches issued on different streams with different
        lanes run concurrently (the solver's CTAs retire as their queue drains, so the tail of one batch -- a few long
        scenarios -- overlaps the head of the next)."""
        import torch

        B = p0.shape[0]
        assert p0.shape == (B, self.N, 2) and p0.dtype == torch.float64 and p0.is_cuda
        z = None
        if v0 is None or vf is None:
            z = torch.zeros_like(p0)
        v0 = z if v0 is None else v0
        vf = z if vf is None else vf
        p0, pf, v0, vf = (t.contiguous() for t in (p0, pf, v0, vf))
        out = torch.empty((3, B, self.N, self.K, 2), dtype=torch.float64, device=self.device)
        rec = torch.empty(B * C.sizeof(_capi.Record), dtype=torch.uint8, device=self.device)
        slots = min(self.max_slots, B)
        ws = self._workspace(slots, lane)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            _capi.check(self.lib.scp_b200_solve_batch(
                C.byref(self.problem), B, p0.data_ptr(), v0.data_ptr(), pf.data_ptr(), vf.data_ptr(),
                self.tables.data_ptr(), ws.data_ptr(), ws.numel(), slots, out[0].data_ptr(), out[1].data_ptr(),
                out[2].data_ptr(), rec.data_ptr(), st))
        return out[0], out[1], out[2], rec

    @staticmethod
    def records_from_bytes(rec_u8):
        raw = rec_u8.cpu().numpy().tobytes()
        n = len(raw) // C.sizeof(_capi.Record)
        arr = (_capi.Record * n).from_buffer_copy(raw)
        return [_capi.record_to_dict(r) for r in arr]

    def solve(self, initial_positions, final_positions, initial_velocities=None, final_velocities=None):
        """Host numpy in, host numpy out (one H2D per input, one D2H per output)."""
        import torch

        def up(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, self.N, 2)
            return torch.from_numpy(a).to(self.device, non_blocking=True)

        acc, pos, vel, rec = self.solve_device(up(initial_positions), up(final_positions),
                                               up(initial_velocities), up(final_velocities))
        torch.cuda.synchronize(self.device)
        return ({"positions": pos.cpu().numpy(), "velocities": vel.cpu().numpy(),
                 "accelerations": acc.cpu().numpy()}, self.records_from_bytes(rec))


def solve_scenarios(initial_positions, final_positions, time_horizon, time_step, min_distance, space_dims=None,
                    device=None, **settings):
    """Convenience: solve a (B,N,2) batch of scenarios once."""
    p0 = np.asarray(initial_positions, dtype=np.float64)
    if p0.ndim == 2:
        p0 = p0[None]
    s = BatchSolver(p0.shape[1], time_horizon, time_step, min_distance, space_dims, device, **settings)
    return s.solve(p0, final_positions)
