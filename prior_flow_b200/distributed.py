"""Multi-GPU plumbing: one process per GPU (torchrun), pairs sharded across ranks.

The hot path shards by image pair — every op on it is per-sample — so inference uses NO
collective; ranks only meet at the timing barrier.  Training replaces the reference's single-process
`nn.DataParallel` (train_flow.py:95-96) with DDP over NCCL: one 33 MB gradient all-reduce per step.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class Context:
    rank: int
    local_rank: int
    world: int
    device: torch.device


def init_from_env(backend: Optional[str] = None) -> Context:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and initialises the process group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    device = torch.device("cuda", local) if backend == "nccl" else torch.device("cpu")
    if backend == "nccl":
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend, rank=rank, world_size=world)
    return Context(rank, local, world, device)


def shard_pairs(total_pairs: int, rank: int, world: int) -> List[int]:
    """Indices of the image pairs rank `rank` owns: contiguous blocks, sizes differing by at most one."""
    base, extra = divmod(total_pairs, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def max_over_ranks(value: float, ctx: Context) -> float:
    """A multi-GPU time is the slowest rank's."""
    if ctx.world == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def ddp_loss_scale(ctx: Context) -> float:
    """The reference sums the loss over pixels and DataParallel sums replica gradients (train_flow.py:69,96); DDP
    averages them.  Multiplying the per-rank loss by world restores the reference's effective step."""
    return float(ctx.world)


def wrap_ddp(model: torch.nn.Module, ctx: Context) -> torch.nn.Module:
    if ctx.world == 1:
        return model
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=[ctx.local_rank] if ctx.device.type == "cuda" else None,
                                                     gradient_as_bucket_view=True)
