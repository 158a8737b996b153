"""Multi-GPU plumbing of the calibration path; the reference has no counterpart.

One process per GPU. Calibration steps are sharded round-robin over ranks (tokens are
independent rows of the SYRK), each rank accumulates a partial d x d fp32 covariance per layer,
and the only data-path exchange is one NCCL reduction of that matrix to the layer's owner
(`layer_idx % world`), which runs the eigensolve and broadcasts the top-k eigenvector block.
Everything here works on whatever device the tensors live on (NCCL for CUDA tensors; the CPU
`gloo` tests drive the same functions with host tensors).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def default_group():
    """The world group when torch.distributed is initialised with more than one rank, else None."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.group.WORLD
    return None


def rank_and_world(group) -> tuple[int, int]:
    if group is None:
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def steps_of_rank(num_steps: int, rank: int, world: int) -> list[int]:
    """Calibration steps a rank runs: i = rank (mod world)."""
    return [i for i in range(num_steps) if i % world == rank]


def owner_of(layer_index: int, world: int) -> int:
    """NVSwitch is uniform, so eigensolve ownership is plain round-robin."""
    return layer_index % world


def _flush(acc) -> None:
    """Staged (deferred) batches must be folded into C before C crosses ranks."""
    flush = getattr(acc, "flush", None)
    if flush is not None:
        flush()


def allreduce_accumulator(acc, group) -> None:
    """Sum the partial covariance (and column sums, step counts) over ranks, in place."""
    if group is None:
        return
    _flush(acc)
    steps = torch.tensor([acc.steps], dtype=torch.int64, device=acc.C.device)
    dist.all_reduce(acc.C, op=dist.ReduceOp.SUM, group=group)
    if getattr(acc, "colsum", None) is not None:
        dist.all_reduce(acc.colsum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(steps, op=dist.ReduceOp.SUM, group=group)
    acc.steps = int(steps.item())


def reduce_accumulator_to(acc, owner: int, group) -> None:
    """Sum the partial covariance onto `owner` only (half the traffic of an all-reduce); the step
    count is summed everywhere so every rank agrees on it."""
    if group is None:
        return
    _flush(acc)
    steps = torch.tensor([acc.steps], dtype=torch.int64, device=acc.C.device)
    dst = dist.get_global_rank(group, owner)
    dist.reduce(acc.C, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if getattr(acc, "colsum", None) is not None:
        dist.reduce(acc.colsum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(steps, op=dist.ReduceOp.SUM, group=group)
    acc.steps = int(steps.item())


def owner_computes(layer_index: int, group, compute: Callable[[], torch.Tensor], acc,
                   shape: tuple[int, int], dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Reduce `acc` to the layer's owner, let the owner run `compute()` (finalize + eigensolve on
    its now-complete accumulator) and broadcast the [d, k] result to every rank."""
    if group is None:
        return compute()
    rank, world = rank_and_world(group)
    owner = owner_of(layer_index, world)
    reduce_accumulator_to(acc, owner, group)
    if rank == owner:
        out = compute().contiguous()
        assert tuple(out.shape) == tuple(shape), (out.shape, shape)
    else:
        out = torch.empty(shape, dtype=dtype, device=acc.C.device)
    dist.broadcast(out, src=dist.get_global_rank(group, owner), group=group)
    return out


def mean_over_ranks(t: torch.Tensor, group) -> torch.Tensor:
    """Average a small metric vector over ranks (used when metric batches are sharded)."""
    if group is None:
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t / dist.get_world_size(group)


def max_over_ranks(value: float, device: torch.device, group=None) -> float:
    """Max of a host scalar over ranks (timing)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
