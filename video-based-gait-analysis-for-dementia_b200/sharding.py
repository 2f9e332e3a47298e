"""Multi-GPU plumbing: one process per GPU, sequences sharded contiguously, weights replicated.

Sequences never interact anywhere on the path, so there is no data-path
collective: each rank runs the whole head on its own block of sequences.  The only exchange the
north_star names is the final gather of joints / meshes, done here with NCCL
(torch.distributed); it is optional and off the compute path.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(num_seqs: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of sequences for `rank`; the first `num_seqs % world_size`
    ranks take one extra."""
    if world_size < 1 or not (0 <= rank < world_size) or num_seqs < 0:
        raise ValueError(f"bad shard request: num_seqs={num_seqs} world_size={world_size} rank={rank}")
    base, rem = divmod(num_seqs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(num_seqs: int, world_size: int) -> list[int]:
    return [shard_bounds(num_seqs, world_size, r)[1] - shard_bounds(num_seqs, world_size, r)[0]
            for r in range(world_size)]


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; initialises the process
    group (NCCL when CUDA is available, else gloo) when WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def gather_sequences(local: torch.Tensor, num_seqs: int, group=None) -> torch.Tensor:
    """All-gather per-rank blocks (S_r, ...) of a sequence-sharded tensor into (num_seqs, ...)
    on every rank, in sequence order.  Even shards use one all_gather_into_tensor; uneven
    shards are padded to the largest block."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local.shape[0] != num_seqs:
            raise ValueError(f"single process holds {local.shape[0]} of {num_seqs} sequences")
        return local
    world = dist.get_world_size(group)
    counts = shard_counts(num_seqs, world)
    rank = dist.get_rank(group)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} sequences, expected {counts[rank]}")
    local = local.contiguous()
    if len(set(counts)) == 1:
        out = torch.empty((num_seqs,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    big = max(counts)
    pad = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host float over all ranks (used for the bench's max-over-ranks step time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
