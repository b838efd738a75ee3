"""Multi-GPU execution of the batch entry point: scenarios are independent units
(reference cli/compute_trajectories_batch.py:103-112 loops over them sequentially), so they
are sharded over ranks with NO data-path collective; only the small result records are
gathered.  One process per GPU, torch.distributed for the plumbing."""

from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local_records, group=None):
    """All ranks receive the concatenation (rank order = scenario order) of the per-scenario records."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return list(local_records)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(local_records), group=group)
    return [r for part in out for r in part]


def solve_scenarios_sharded(solve_fn, initial_positions, final_positions, group=None):
    """solve_fn(p0_slice, pf_slice) -> (trajectories dict, records) on this rank's GPU.
    Returns this rank's trajectories and the global record list."""
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(len(initial_positions), rank, world)
    traj, recs = solve_fn(initial_positions[lo:hi], final_positions[lo:hi])
    for i, r in enumerate(recs):
        r["scenario_index"] = lo + i
        r["rank"] = rank
    return traj, gather_records(recs, group)
