"""Streaming solver: host side of `scp_b200_stream_*` (include/scp_b200.h).

Two uses of the same device code:

* ``StreamSolver(..., n_scenarios=B)`` -- a batch of scenarios whose per-scenario state does not fit one CTA's
  shared memory (BASELINE.json configs 3 and 5).  Same contract as ``BatchSolver.solve_device``.
* ``StreamSolver(..., group=<torch.distributed group>)`` -- ONE scenario whose agents are sharded over the GPUs of
  a node (config 4).  Rank g owns a block of agents; every ADMM iteration ends with an NCCL all-gather of the
  positions, issued by the library on its own communicator (the unique id travels through torch.distributed).
  Every rank passes the same inputs and receives the full result.

PyTorch is plumbing only (device buffers, the current stream, the id broadcast).  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _capi
from .batch import BatchSolver, _require_cuda


def agent_block(n_agents: int, rank: int, world: int):
    """Agents [lo, hi) owned by `rank`: equal blocks of ceil(N/world) (the all-gather needs equal slices)."""
    per = -(-n_agents // world)
    return min(rank * per, n_agents), min((rank + 1) * per, n_agents)


class StreamSolver:
    def __init__(self, n_vehicles, time_horizon, time_step, min_distance, space_dims=None, n_scenarios=1,
                 max_candidates=16, group=None, sharded=False, peer_exchange=False, **settings):
        torch = _require_cuda()
        self.torch = torch
        self.lib = _capi.load()
        self.problem = _capi.default_problem(n_vehicles, time_horizon, time_step, min_distance, space_dims, stream=True)
        for k, v in settings.items():
            if not hasattr(self.problem, k):
                raise TypeError(f"unknown solver setting {k!r}")
            setattr(self.problem, k, v)
        self.N, self.K, self.B = int(self.problem.n_agents), int(self.problem.n_steps), int(n_scenarios)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.rank, self.world = 0, 1
        idbuf = None
        if sharded or group is not None:
            import torch.distributed as dist

            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            if self.world > 1:
                raw = bytearray(128)
                if self.rank == 0:
                    buf = (C.c_char * 128)()
                    _capi.check(self.lib.scp_b200_nccl_unique_id(C.cast(buf, C.c_void_p)))
                    raw = bytearray(buf.raw)
                box = [bytes(raw)]
                dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                idbuf = (C.c_char * 128).from_buffer_copy(box[0])
        self.lo, self.hi = agent_block(self.N, self.rank, self.world)
        self.max_candidates = int(max_candidates)
        self._h = None
        self._create(idbuf)
        self.exchange = "none" if self.world == 1 else "nccl"
        if self.world > 1 and peer_exchange:
            # peer-memory exchange of the position slices (NVLink stores + flags) instead of an NCCL call per iteration
            import torch.distributed as dist

            mine = (C.c_char * 192)()
            ok = self.lib.scp_b200_stream_ipc_handles(self._h, C.cast(mine, C.c_void_p)) == 0
            box = [None] * self.world
            dist.all_gather_object(box, bytes(mine.raw) if ok else None, group=group)
            if all(b is not None for b in box):
                blob = (C.c_char * (192 * self.world)).from_buffer_copy(b"".join(box))
                ok = self.lib.scp_b200_stream_ipc_connect(self._h, C.cast(blob, C.c_void_p)) == 0
            else:
                ok = False
            flags = [None] * self.world
            dist.all_gather_object(flags, bool(ok), group=group)
            if all(flags):
                self.exchange = "peer"
            elif ok:
                raise _capi.ScpB200Error("peer exchange connected on some ranks only")
            dist.barrier(group=group)
        self.last_device_ms = None
        self.last_macro_steps = None

    MAX_CANDIDATES_LIMIT = 48          # MAXC_MAX in csrc/scp_stream.cu

    def _create(self, idbuf=None):
        if self._h:
            self.lib.scp_b200_stream_destroy(self._h)
            self._h = None
        h = C.c_void_p()
        _capi.check(self.lib.scp_b200_stream_create(
            C.byref(self.problem), self.B, self.max_candidates, self.rank, self.world,
            C.cast(idbuf, C.c_void_p) if idbuf is not None else None, C.byref(h)))
        self._h = h

    def _overflow_retry(self, records):
        """True when a scenario dropped collision rows for lack of candidate slots and the solver was rebuilt with
        twice the capacity (single-GPU solver only; the caller solves again).  Otherwise warns if rows were dropped."""
        if not any(r["candidate_overflow"] for r in records):
            return False
        cap = min(self.MAX_CANDIDATES_LIMIT, max(self.N - 1, 1))
        if self.world == 1 and self.max_candidates < cap:
            self.max_candidates = min(cap, 2 * self.max_candidates)
            self._create()
            return True
        import warnings

        n = sum(1 for r in records if r["candidate_overflow"])
        warnings.warn(f"streaming SCP solver: {n} scenario(s) dropped collision rows (more than {self.max_candidates} "
                      "partners per agent and step); their subproblems are reported as unsolved", RuntimeWarning)
        return False

    def close(self):
        if getattr(self, "_h", None):
            self.lib.scp_b200_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve_device(self, p0, pf, v0=None, vf=None):
        """(B,N,2) float64 CUDA tensors in; acc,pos,vel (B,N,K,2) and a uint8 tensor of B records out (blocking)."""
        torch = self.torch
        B = p0.shape[0]
        assert B == self.B and p0.shape == (B, self.N, 2) and p0.dtype == torch.float64 and p0.is_cuda
        z = torch.zeros_like(p0) if (v0 is None or vf is None) else None
        v0 = z if v0 is None else v0
        vf = z if vf is None else vf
        p0, pf, v0, vf = (t.contiguous() for t in (p0, pf, v0, vf))
        out = torch.empty((3, B, self.N, self.K, 2), dtype=torch.float64, device=self.device)
        rec = torch.empty(B * C.sizeof(_capi.Record), dtype=torch.uint8, device=self.device)
        ms, steps = C.c_float(0.0), C.c_int64(0)
        st = torch.cuda.current_stream(self.device).cuda_stream
        while True:
            _capi.check(self.lib.scp_b200_stream_solve(
                self._h, p0.data_ptr(), v0.data_ptr(), pf.data_ptr(), vf.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                out[2].data_ptr(), rec.data_ptr(), st, C.byref(ms), C.byref(steps)))
            if not self._overflow_retry(self.records_from_bytes(rec)):
                break
        self.last_device_ms, self.last_macro_steps = float(ms.value), int(steps.value)
        return out[0], out[1], out[2], rec

    records_from_bytes = staticmethod(BatchSolver.records_from_bytes)

    def solve(self, initial_positions, final_positions, initial_velocities=None, final_velocities=None):
        """Host numpy in, host numpy out through `scp_b200_stream_solve_host`."""
        def buf(a):
            if a is None:
                return np.zeros((self.B, self.N, 2))
            return np.ascontiguousarray(a, dtype=np.float64).reshape(self.B, self.N, 2)

        p0, pf, v0, vf = buf(initial_positions), buf(final_positions), buf(initial_velocities), buf(final_velocities)
        acc, pos, vel = (np.empty((self.B, self.N, self.K, 2)) for _ in range(3))
        recs = (_capi.Record * self.B)()
        ms, steps = C.c_float(0.0), C.c_int64(0)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        while True:
            _capi.check(self.lib.scp_b200_stream_solve_host(
                self._h, ptr(p0), ptr(v0), ptr(pf), ptr(vf), ptr(acc), ptr(pos), ptr(vel), C.cast(recs, C.c_void_p),
                C.byref(ms), C.byref(steps)))
            out = [_capi.record_to_dict(r) for r in recs]
            if not self._overflow_retry(out):
                break
        self.last_device_ms, self.last_macro_steps = float(ms.value), int(steps.value)
        return ({"positions": pos, "velocities": vel, "accelerations": acc}, out)
