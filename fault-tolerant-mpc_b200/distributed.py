"""Multi-GPU partitioning of a Monte-Carlo scenario batch.

Every MPC instance (fault scenario x initial state) is independent, so the batch is split into contiguous
blocks -- rank r owns instances [lo_r, hi_r) -- with NO collective inside a solve or a closed-loop rollout.
The only exchange is one all-gather of the per-instance closed-loop results at the end
(`torch.distributed.all_gather_into_tensor`; NCCL on the GPUs, gloo in the CPU tests).
The reference has no counterpart: it is single-process (examples/sim.py runs one instance).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RESULT_WIDTH = 16      # per-instance result record: final robot state (13) | cumulative cost | status | steps done


def shard_bounds(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of rank `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    if not (0 <= rank < world) or batch < 0:
        raise ValueError(f"bad shard request batch={batch} rank={rank} world={world}")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(final_state: torch.Tensor, cost: torch.Tensor, status: torch.Tensor, steps: torch.Tensor) -> torch.Tensor:
    """[b,13], [b], [b], [b] -> [b,16] fp64 record (status / steps are small integers, exact in fp64)."""
    rec = torch.empty(final_state.shape[0], RESULT_WIDTH, dtype=torch.float64, device=final_state.device)
    rec[:, 0:13] = final_state
    rec[:, 13] = cost
    rec[:, 14] = status.to(torch.float64)
    rec[:, 15] = steps.to(torch.float64)
    return rec


def gather_results(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gather the per-rank result records into the global [batch, RESULT_WIDTH] tensor, in instance order.
    `local` holds the records of this rank's shard (shard_bounds).  One collective; shards are padded to the
    largest shard so that all_gather_into_tensor (equal sizes) applies."""
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[0] != batch:
            raise ValueError("single-process gather needs the whole batch")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(batch, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} records, its shard has {hi - lo}")
    cap = (batch + world - 1) // world
    send = torch.zeros(cap, local.shape[1], dtype=local.dtype, device=local.device)
    send[: hi - lo] = local
    recv = torch.empty(world * cap, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = torch.empty(batch, local.shape[1], dtype=local.dtype, device=local.device)
    for r in range(world):
        a, b = shard_bounds(batch, r, world)
        out[a:b] = recv[r * cap: r * cap + (b - a)]
    return out
