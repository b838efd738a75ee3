"""Agent-sharded collision linearisation of ONE large scenario over the GPUs of a node (BASELINE.json config 4).

The reference builds the collision rows of `_add_collision_constraints` (scp.py:453-557) for all pairs in one
Python loop.  Here rank g owns a contiguous block of agents and the pair rows (i, j>i) of its agents; once per SCP
iteration every rank contributes the positions of its agents to an NCCL all-gather (N*K*2 doubles in total, 1.6 MB
at N=1000, K=100) and then runs `scp_b200_linearize_range` on its share of the pair indices; the minimum
separation and the first violating row (the gate of scp.py:597-615) are combined with a MIN all-reduce.
Blocks are balanced by pair count (agent i owns N-1-i rows per step), not by agent count.

This module shards the pairwise kernel alone -- the O(N^2 K) part of an SCP iteration.  The complete agent-sharded
solve (QP included: all-gather or peer-memory exchange of the positions per ADMM iteration) is
`solvers/stream.py::StreamSolver(..., sharded=True)` on top of `scp_b200_stream_*`.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _capi


def pair_index(i: int, n: int) -> int:
    """Index of pair (i, i+1) in the reference's i<j lexicographic order (scp.py:495-496)."""
    return i * (2 * n - i - 1) // 2


def agent_blocks(n_agents: int, world: int):
    """Agent boundaries [a_0=0, a_1, ..., a_world=N] so that every rank owns about the same number of pair rows."""
    total = n_agents * (n_agents - 1) // 2
    bounds = [0]
    for g in range(1, world):
        target = g * total / world
        i = bounds[-1]
        while i < n_agents and pair_index(i, n_agents) < target:
            i += 1
        bounds.append(i)
    bounds.append(n_agents)
    return bounds


class ShardedLinearizer:
    """One instance per rank; `positions_own` are this rank's agents' trajectories (n_own, K, 2) on its GPU."""

    def __init__(self, n_agents, n_steps, min_distance, feas_margin=0.01, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.N, self.K, self.R, self.margin = n_agents, n_steps, float(min_distance), float(feas_margin)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bounds = agent_blocks(n_agents, self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.p_begin = pair_index(self.lo, n_agents)
        self.p_end = pair_index(self.hi, n_agents) if self.hi < n_agents else n_agents * (n_agents - 1) // 2
        self.lib = _capi.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.pos_all = torch.empty((n_agents, n_steps, 2), dtype=torch.float64, device=dev)
        rows = self.p_end - self.p_begin
        self.eta = torch.empty((n_steps, max(rows, 1), 2), dtype=torch.float64, device=dev)
        self.bound = torch.empty((n_steps, max(rows, 1)), dtype=torch.float64, device=dev)
        self.red = torch.empty(2, dtype=torch.float64, device=dev)      # min separation, first violating row (as float)
        self.minsep = torch.empty(1, dtype=torch.float64, device=dev)
        self.first = torch.empty(3, dtype=torch.int32, device=dev)

    def gather_positions(self, positions_own):
        """NCCL all-gather of the trajectories (uneven blocks -> all_gather into views of the full array)."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            self.pos_all.copy_(positions_own)
            return self.pos_all
        chunks = [self.pos_all[self.bounds[g]:self.bounds[g + 1]] for g in range(self.world)]
        dist.all_gather(chunks, positions_own.contiguous(), group=self.group)
        return self.pos_all

    def linearize(self, positions_own):
        """Returns this rank's rows eta (K,rows,2), bound (K,rows) and a 2-element device tensor with the GLOBAL
        minimum separation and first violating row index (decode() turns it into (k,i,j))."""
        torch, dist = self.torch, self.dist
        pos = self.gather_positions(positions_own)
        st = torch.cuda.current_stream().cuda_stream
        _capi.check(self.lib.scp_b200_linearize_range(
            pos.data_ptr(), 1, self.N, self.K, self.R, self.margin, self.p_begin, self.p_end, self.eta.data_ptr(),
            self.bound.data_ptr(), self.minsep.data_ptr(), self.first.data_ptr(), st))
        # reductions stay on the device (no host sync inside an SCP iteration): row = k P + p(i, j), +inf when none
        P = self.N * (self.N - 1) // 2
        f = self.first.to(torch.float64)
        k, i, j = f[0], f[1], f[2]
        row = k * P + (i * (2 * self.N - i - 1)) * 0.5 + (j - i - 1)          # exact in fp64 for N <= 2^20
        self.red[0] = self.minsep[0]
        self.red[1] = torch.where(k >= 0, row, torch.full_like(row, float("inf")))
        if self.world > 1:
            dist.all_reduce(self.red, op=dist.ReduceOp.MIN, group=self.group)
        return self.eta, self.bound, self.red

    def decode(self, red):
        """Host view of the reductions: (min separation, first violating (k, i, j) or None).  Synchronises."""
        minsep, row = (float(v) for v in red.tolist())
        if not np.isfinite(row):
            return minsep, None
        P = self.N * (self.N - 1) // 2
        k, p = divmod(int(row), P)
        i = 0
        while pair_index(i + 1, self.N) <= p:
            i += 1
        return minsep, (k, i, i + 1 + p - pair_index(i, self.N))
