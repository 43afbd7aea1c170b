"""Multi-GPU: episodes shard trivially.  One process per GPU; rank g owns the contiguous
episode range [g E / G, (g + 1) E / G); there is NO data-path collective.  Only the final metric gather and
the reduction of the aggregate row cross ranks (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import _abi as abi


def shard_range(n_episodes, rank, world_size):
    """Contiguous, balanced: the first ``n_episodes % world_size`` ranks get one extra episode."""
    base, extra = divmod(int(n_episodes), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_episodes, world_size):
    return [shard_range(n_episodes, r, world_size)[1] - shard_range(n_episodes, r, world_size)[0]
            for r in range(world_size)]


def combine_aggregates(rows):
    """Reduce per-rank aggregate rows (G, BGR_N_AGG) in RANK ORDER (deterministic): sums, except the
    minimum lap time."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, abi.N_AGG)
    out = np.zeros(abi.N_AGG)
    for r in rows:                       # fixed order: rank 0, 1, ...
        out += r
    out[abi.A_MIN_LAP_TIME] = rows[:, abi.A_MIN_LAP_TIME].min()
    return out


def gather_results(metrics, status, aggregate, n_episodes, group=None, dst=0, rows=True):
    """The final cross-rank step: GATHER the per-episode ``metrics`` / ``status`` blocks of every rank onto rank
    ``dst`` (tensors on this rank's device; ragged shards are padded to the largest one) and REDUCE the aggregate rows
    in rank order.  Returns ``(metrics (E, 12), status (E, 10), aggregate (8,) float64 tensor)``; the two big blocks
    are ``None`` on every rank but ``dst`` -- only one rank holds O(E) memory.  The aggregate row (64 B per rank,
    all-gathered and summed in a fixed order) is identical on every rank.

    ``dst=None``: all-gather instead (every rank gets the blocks; O(world) memory per rank).
    ``rows=False``: reduce only -- for callers that want the aggregate and leave the rows where they are."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    agg_rows = torch.empty((world, abi.N_AGG), dtype=torch.float64, device=aggregate.device)
    dist.all_gather_into_tensor(agg_rows, aggregate.reshape(1, abi.N_AGG).contiguous(), group=group)
    total = torch.zeros(abi.N_AGG, dtype=torch.float64, device=aggregate.device)
    for r in range(world):               # fixed order
        total = total + agg_rows[r]
    total[abi.A_MIN_LAP_TIME] = agg_rows[:, abi.A_MIN_LAP_TIME].min()
    if not rows:
        return None, None, total

    sizes = shard_sizes(n_episodes, world)
    cap = max(sizes)

    def _gather(t, width):
        pad = t
        if t.shape[0] != cap:
            pad = torch.zeros((cap, width), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
        if dst is None:
            out = torch.empty((world * cap, width), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
        else:
            out = torch.empty((world * cap, width), dtype=t.dtype, device=t.device) if rank == dst else None
            dist.gather(pad.contiguous(), list(out.view(world, cap, width).unbind(0)) if rank == dst else None,
                        dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
            if rank != dst:
                return None
        if all(sz == cap for sz in sizes):
            return out
        return torch.cat([out[r * cap: r * cap + sizes[r]] for r in range(world)], dim=0)

    return _gather(metrics, abi.N_METRICS), _gather(status, abi.N_STATUS), total
