"""Multi-GPU plumbing: clips shard by contiguous ranges of the global clip index (every clip is independent in
every stage), weights are replicated, and the only exchange is an all-gather of the packed
(B_local, C + 1) fp32 records [mean_logits | uncertainty].  Device-agnostic torch.distributed code: NCCL over
NVLink on the GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous share; the first n % world ranks get one extra clip."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(n_clips), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_counts(n_clips: int, world_size: int) -> List[int]:
    return [shard_range(n_clips, r, world_size)[1] - shard_range(n_clips, r, world_size)[0]
            for r in range(world_size)]


def all_gather_packed(packed_local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """packed_local: this rank's (count_rank, W) records in global clip order -> (n_clips, W) on every rank.
    One collective: shards are padded to the largest count (they differ by at most one row)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if packed_local.shape[0] != n_clips:
            raise ValueError("single process must hold every clip")
        return packed_local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = shard_counts(n_clips, world)
    if packed_local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {packed_local.shape[0]} rows, expected {counts[rank]}")
    W = packed_local.shape[1]
    cmax = max(counts)
    send = packed_local
    if counts[rank] != cmax:
        send = torch.zeros(cmax, W, dtype=packed_local.dtype, device=packed_local.device)
        send[:counts[rank]] = packed_local
    recv = torch.empty(world * cmax, W, dtype=packed_local.dtype, device=packed_local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(c == cmax for c in counts):
        return recv
    recv = recv.view(world, cmax, W)
    return torch.cat([recv[r, :counts[r]] for r in range(world)], dim=0)


def labels_from_packed(packed: torch.Tensor) -> torch.Tensor:
    """argmax over the mean-logit columns (first maximum wins, like torch.argmax)."""
    return packed[:, :-1].argmax(dim=1)
