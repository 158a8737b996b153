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


def resolve_group(process_group):
    """The explicit opt-in of the drivers: None -> single GPU; "world" -> the default group when
    it has more than one rank; a ProcessGroup -> itself (None when it has one rank)."""
    if process_group is None:
        return None
    if isinstance(process_group, str):
        if process_group != "world":
            raise ValueError(f"process_group must be None, 'world' or a ProcessGroup, got {process_group!r}")
        return default_group()
    return process_group if dist.get_world_size(process_group) > 1 else None


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


TRIANGLE_BANDS = 8  # row bands of the packed lower-triangle exchange: 1/2 + 1/(2*8) = 56 % of d^2


def _bands(d: int) -> list[tuple[int, int]]:
    step = max(1, -(-d // TRIANGLE_BANDS))
    return [(r0, min(d, r0 + step)) for r0 in range(0, d, step)]


def start_reduce_lower(acc, owner: int, group, total_steps: Optional[int] = None) -> list:
    """Asynchronously sum the LOWER TRIANGLE of the partial covariance onto `owner` (the SYRK kernel
    leaves the strict upper triangle unspecified until finalize, include/ptdeco_b200.h, so sending
    it would move garbage): the triangle goes in row bands C[r0:r1, :r1] staged contiguously.
    Returns the handles `finish_reduce_lower` consumes. Nothing blocks the host: the caller can
    keep enqueueing eigensolves of earlier layers while NCCL moves this one. `total_steps` (the
    calibration's step count, known to every rank) avoids a host-synchronising count exchange."""
    if group is None:
        return []
    _flush(acc)
    dst = dist.get_global_rank(group, owner)
    me = dist.get_rank(group)
    handles = []
    d = acc.C.shape[0]
    for r0, r1 in _bands(d):
        buf = acc.C[r0:r1, :r1].contiguous()
        work = dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group, async_op=True)
        handles.append((work, buf, r0, r1))
    if getattr(acc, "colsum", None) is not None:
        handles.append((dist.reduce(acc.colsum, dst=dst, op=dist.ReduceOp.SUM, group=group,
                                    async_op=True), None, 0, 0))
    if total_steps is not None:
        acc.steps = int(total_steps)
    else:
        steps = torch.tensor([acc.steps], dtype=torch.int64, device=acc.C.device)
        dist.all_reduce(steps, op=dist.ReduceOp.SUM, group=group)
        acc.steps = int(steps.item())
    acc._reduce_owner = (me == owner)
    return handles


def finish_reduce_lower(acc, handles: list) -> None:
    """Wait (stream-side) for `start_reduce_lower` and, on the owner, put the summed bands back."""
    for work, buf, r0, r1 in handles:
        work.wait()
        if buf is not None and getattr(acc, "_reduce_owner", False):
            acc.C[r0:r1, :r1].copy_(buf)


def reduce_accumulator_to(acc, owner: int, group, total_steps: Optional[int] = None) -> None:
    """Sum the partial covariance onto `owner` only: lower triangle, half the bytes of a full
    d x d reduce and a quarter of an all-reduce."""
    if group is None:
        return
    finish_reduce_lower(acc, start_reduce_lower(acc, owner, group, total_steps))


def gather_shards_to(acc, owner: int, group, total_steps: int) -> None:
    """Reproducibility mode of the exchange (CovarianceAccumulator with canonical shards): instead
    of an NCCL reduction, whose summation order follows the ring / tree of the day and the world
    size, the owner receives every non-empty shard C_v from its holder (v mod world) and adds them
    in index order -- the arithmetic of a single GPU holding all V shards. Full matrices, point to
    point; every rank walks (v = 0..V-1) in the same order, so the sends and receives pair up."""
    rank, world = rank_and_world(group)
    if acc.shard_world != world:
        raise ValueError(f"accumulator was sharded for {acc.shard_world} ranks, group has {world}")
    if rank == owner:
        acc.C.zero_()
    tmp = None
    for v in range(min(acc.shards, int(total_steps))):  # shard v is non-empty iff step v exists
        holder = v % world
        mine = acc.shard_C[v]
        if holder == owner:
            if rank == owner:
                acc.C += mine
        elif rank == holder:
            dist.send(mine, dst=dist.get_global_rank(group, owner), group=group)
        elif rank == owner:
            if tmp is None:
                tmp = torch.empty_like(acc.C)
            dist.recv(tmp, src=dist.get_global_rank(group, holder), group=group)
            acc.C += tmp
        if rank == holder:
            acc.shard_C[v] = None
    acc._collapsed = True
    acc.steps = int(total_steps)


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


def balanced_owners(costs: list[float], world: int) -> list[int]:
    """Owner of every job by greedy longest-processing-time assignment (deterministic: every rank
    computes the same table). Plain round-robin hands whole CLASSES of jobs to the same ranks when
    the job list is periodic -- a decoder's accumulators come as (q/k/v, o, gate/up, down) per layer,
    so with 2 or 8 ranks the gate/up units (three eigensolves each) all landed on the same ranks."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owners = [0] * len(costs)
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owners[i] = r
        load[r] += costs[i]
    return owners


def owners_compute_pipelined(jobs: list, group, total_steps: Optional[int] = None,
                             dtype: torch.dtype = torch.float32,
                             costs: Optional[list[float]] = None) -> list[torch.Tensor]:
    """All accumulators of a calibration split at once: jobs[i] = (acc, compute, shape); owners
    are balanced by `costs` (relative eigensolve cost per job; round-robin without). Three phases,
    each without a host synchronisation:
      1. every reduction (lower-triangle bands) is enqueued up front; NCCL runs them back to back;
      2. every rank runs the eigensolves of ITS jobs back to back;
      3. the [d, k] results are broadcast in job order.
    Phases 2 and 3 are deliberately NOT interleaved: the eigensolver's tridiagonalisation kernels
    are cooperative launches that need every SM, and a broadcast kernel spinning on a peer that is
    still computing keeps SMs occupied -- interleaved, the ranks' eigensolves ran one after the
    other (measured at 2 GPUs: 9.1 s instead of 4.4 s for the 8B-shape split).
    Returns the results in job order on every rank; they are valid for work enqueued on the
    current stream after this function returns."""
    if group is None:
        return [compute() for _, compute, _ in jobs]
    rank, world = rank_and_world(group)
    owners = (balanced_owners(costs, world) if costs is not None
              else [owner_of(i, world) for i in range(len(jobs))])
    if any(getattr(acc, "shards", 1) > 1 for acc, _, _ in jobs):
        if total_steps is None:
            raise ValueError("canonical shards need the calibration's total step count")
        for i, (acc, _, _) in enumerate(jobs):
            gather_shards_to(acc, owners[i], group, total_steps)
    else:
        pending = [start_reduce_lower(acc, owners[i], group, total_steps)
                   for i, (acc, _, _) in enumerate(jobs)]
        for i, (acc, _, _) in enumerate(jobs):
            finish_reduce_lower(acc, pending[i])
        pending.clear()
    outs: list = [None] * len(jobs)
    for i, (acc, compute, shape) in enumerate(jobs):
        if rank == owners[i]:
            out = compute().contiguous()
            assert tuple(out.shape) == tuple(shape), (out.shape, shape)
            outs[i] = out
    works = []
    for i, (acc, _, shape) in enumerate(jobs):
        owner = owners[i]
        if outs[i] is None:
            outs[i] = torch.empty(shape, dtype=dtype, device=acc.C.device)
        works.append(dist.broadcast(outs[i], src=dist.get_global_rank(group, owner), group=group,
                                    async_op=True))
    for w in works:
        w.wait()
    return outs


def check_identical_batches(batch, group) -> None:
    """The sharded paths assume every rank draws the SAME batches from its iterators (each rank
    skips the ones that are not its share). A cheap guard: compare a checksum of one batch across
    ranks and fail loudly on rank-sharded loaders instead of silently calibrating on 1/world of
    the data with diverging replicas."""
    if group is None:
        return
    tensors = [batch] if isinstance(batch, torch.Tensor) else [
        v for _, v in sorted(batch.items()) if isinstance(v, torch.Tensor)] if isinstance(batch, dict) else []
    if not tensors:
        return
    dev = tensors[0].device
    sig = torch.stack([t.double().sum().to(dev) + t.numel() for t in tensors]).sum().reshape(1)
    world = dist.get_world_size(group)
    if sig.device.type == "cpu" and dist.get_backend(group) == "nccl":
        sig = sig.cuda()
    gathered = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(gathered, sig, group=group)
    vals = [float(g.item()) for g in gathered]
    if any(v != vals[0] for v in vals):
        raise RuntimeError(
            "ptdeco_b200 multi-GPU decomposition needs identical data / metric iterators on every "
            f"rank (batch checksums differ across ranks: {vals}); do not shard the loaders by rank")


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
