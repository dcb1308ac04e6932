"""Data-parallel plumbing: one process per GPU, replicated weights, per-rank image shards, one
all-reduce (sum) of the flat gradient buffer per step; the 1/world_size mean is folded into the
optimiser kernel (wap_adadelta_l2_step grad_scale).  This reproduces the semantics of the
reference's SyncReplicasOptimizer (model-distributed.py:704-717: mean of the N replica gradients
applied once) without its parameter server / gRPC mechanism.  Batch-norm batch statistics stay
per-replica, as in the reference (model-distributed.py:140); moving statistics are averaged across
ranks by `sync_moving_stats` before evaluation/checkpointing."""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun); returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            os.environ.setdefault("NCCL_NET_DISABLE", "1")      # one node: NVLink / NVSwitch only
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_batches(batches: Sequence, rank: int, world: int) -> List:
    """Disjoint round-robin shard of the (area-sorted) batch list, equal length on every rank
    (trailing remainder dropped) so that every rank performs the same number of all-reduces.
    (The reference slices an independently shuffled list per worker, model-distributed.py:793-798,
    so its shards overlap; disjoint shards are the intended semantics.)"""
    n = (len(batches) // world) * world
    return [batches[i] for i in range(rank, n, world)]


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def sync_moving_stats(moving: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(moving, op=dist.ReduceOp.SUM)
        moving.div_(dist.get_world_size())
    return moving


def broadcast_params(flat: torch.Tensor, src: int = 0) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat
