"""dwain (Decomposing Weights Algorithm - an Iterative techNique) on the sm_100a kernels.

Drop-in for the reference's src/ptdeco/dwain/decomposition.py (cited as D:line): same public
entry points, same private primitives the reference's tests import (`_wrap_in_place`,
`_unwrap_in_place`, `_compute_covariance_matrix_decomposition`), reversed layer order, geometric
rank descent where the smallest accepted rank wins, optional one-pass covariance precompute in
splits, per-layer `finetune_fn`, same `decompose_config`.

  D:152      Eyyt += einsum(y, y) / N          -> tcgen05 SYRK, fp32 accumulation of exact bf16
                                                  products (the reference rounds each per-step
                                                  product to bf16 for bf16 models);
                                                  for in < out layers the INPUT covariance is
                                                  accumulated and C = W S W^T is solved as an
                                                  in x in problem (same eigenvectors)
  D:158-162  damping + torch.linalg.eigh       -> ptdeco_cov_finalize + ptdeco_eigh (top-k)
  D:193-204  CovarianceComputingLinearModule   -> layer forward on the tcgen05 GEMM engine + SYRK
  D:427-429  U = W^T uk ; (U V)^T              -> ptdeco_gemm
  D:269-277  NSR on logits                     -> ptdeco_nsr_metric

Multi-GPU: when torch.distributed is initialised the calibration steps are sharded over ranks and
only the d x d fp32 covariances cross NVLink (ptdeco_b200.parallel). CUDA only; no CPU path.
"""
from __future__ import annotations

import collections.abc
import logging
import os
import time
from typing import Any, Optional

import torch

from .. import _native as nat
from .. import _wrap, linalg, parallel, utils

__all__ = ["decompose_in_place", "is_decomposeable_module"]

EIGEN_DAMPEN_FACTOR = 0.01  # D:14

logger = logging.getLogger("ptdeco.dwain.decomposition")


class WrappedDWAINModule(_wrap.WrappedModule):
    pass


class WrappedDWAINLinear(_wrap.WrappedLinear, WrappedDWAINModule):
    pass


class WrappedDWAINConv2d1x1(_wrap.WrappedConv2d1x1, WrappedDWAINModule):
    pass


is_decomposeable_module = _wrap.is_decomposeable_module
_is_num_params_reduced = _wrap.is_num_params_reduced


def decompose_in_place(
    *,
    module: torch.nn.Module,
    device: torch.device,
    data_iterator: collections.abc.Iterator[dict[str, torch.Tensor]],
    loss_fn: collections.abc.Callable[[dict[str, torch.Tensor], torch.Tensor], torch.Tensor],
    num_data_steps: int,
    metric_iterator: collections.abc.Iterator[dict[str, torch.Tensor]],
    num_metric_steps: int,
    blacklisted_module_names: Optional[list[str]] = None,
    nsr_final_threshold: float,
    finetune_fn: collections.abc.Callable[[torch.nn.Module, torch.device, list[str]], torch.nn.Module],
    min_rank: int = 32,
    trade_off_factor: float = 0.5,
    reduction_factor: float = 0.5,
    max_accepted_ppl_diff: float = 0.1,
    decompose_in_float64: bool = True,
    precomputing_covariance_num_splits: Optional[int] = None,
    trace: Optional[list] = None,
    process_group: Any = None,
) -> dict[str, Any]:
    """D:677-800. Returns the decompose_config (insertion order = reversed module order).

    `process_group` (not in the reference) opts into the multi-GPU path: pass a torch.distributed
    group (or "world") when EVERY rank runs this call on its own replica of the model with
    IDENTICAL data / metric iterators and a deterministic `finetune_fn`. Calibration steps, rank
    trials and metric batches are then sharded over the ranks and every rank ends with the same
    modules. Without it the call is single-GPU even when torch.distributed is initialised."""
    start_time = time.perf_counter()
    device = torch.device(device)
    if device.type != "cuda":
        raise nat.NativeError("ptdeco_b200.dwain runs on CUDA (sm_100a) only; there is no CPU path")
    nat.lib()
    num_params = utils.get_num_params(module)
    current_params = num_params
    if blacklisted_module_names is None:
        blacklisted_module_names = []
    names = _get_decomposeable_submodule_names(module, blacklisted_module_names)
    n = len(names)
    n_decomposed = 0
    logger.info("\n".join([f"There are {n} linear modules that can be decomposed:"]
                          + [f"  {i}. {nm}" for i, nm in enumerate(names, start=1)]))

    decompose_config: dict[str, Any] = {}
    decomposed_submodules: list[str] = []
    group = parallel.resolve_group(process_group)

    if precomputing_covariance_num_splits is not None and precomputing_covariance_num_splits > 0:
        u_dict = _precompute_covariance_matrix_decompositions_in_splits(
            module=module, modules_to_decompose=names,
            num_splits=precomputing_covariance_num_splits, data_iterator=data_iterator,
            num_data_steps=num_data_steps, device=device,
            decompose_in_float64=decompose_in_float64, reduction_factor=reduction_factor,
            group=group)
    else:
        logger.info("Skipping precomputing convariance matrices")
        u_dict = {}
    utils.relieve_gpu_memory_pressure()
    pair_state = _wrap.PairState(module)

    for i, name in enumerate(reversed(names), start=1):
        logger.info(f"PROCESSING {name} MODULE {i} OUT OF {n}")
        with torch.no_grad():
            logger.info(f"start reserved gpu mem={utils.get_gpu_reserved_memory_gb():.2f} GB")
            result = _process_module(
                root_module=module, decomposed_submodule_name=name, data_iterator=data_iterator,
                loss_fn=loss_fn, metric_iterator=metric_iterator,
                nsr_final_threshold=nsr_final_threshold, num_data_steps=num_data_steps,
                num_metric_steps=num_metric_steps, device=device, num_params=num_params,
                trade_off_factor=trade_off_factor, reduction_factor=reduction_factor,
                max_accepted_ppl_diff=max_accepted_ppl_diff, min_rank=min_rank,
                decompose_in_float64=decompose_in_float64,
                u_matrix=u_dict.pop(name) if len(u_dict) > 0 else None, trace=trace,
                pair_state=pair_state, group=group)
            logger.info(f"stop reserved gpu mem={utils.get_gpu_reserved_memory_gb():.2f} GB")
        current_params -= result.get("drop_in_params", 0)
        logger.info(f"CURRENT PARAMS IN M: {current_params / 1e6}")
        new_module = result["decomposed_module"]
        proportion = result["proportion"]
        if new_module is not None:
            decomposed_submodules.append(name)
            utils.replace_submodule_in_place(module, name, new_module)
            module = finetune_fn(module, device, decomposed_submodules)
            utils.relieve_gpu_memory_pressure()
            module_config = utils.get_module_config(new_module)
            _add_meta_to_module_config(module_config, result)
            decompose_config[name] = module_config
            logger.info(f"{name} decomposed with rank {proportion=:.4f}")
            n_decomposed += 1
        utils.relieve_gpu_memory_pressure()

    logger.info(f"Decomposed {n_decomposed} out of {n} modules")
    logger.info(f"Decomposition took {time.perf_counter() - start_time:.1f} seconds")
    return decompose_config


def _process_module(
    *,
    root_module: torch.nn.Module,
    decomposed_submodule_name: str,
    data_iterator: collections.abc.Iterator[dict[str, torch.Tensor]],
    loss_fn: collections.abc.Callable[[dict[str, torch.Tensor], torch.Tensor], torch.Tensor],
    nsr_final_threshold: float,
    num_data_steps: int,
    num_metric_steps: int,
    device: torch.device,
    metric_iterator: collections.abc.Iterator[dict[str, torch.Tensor]],
    num_params: int,
    min_rank: int = 32,
    trade_off_factor: float,
    reduction_factor: float,
    max_accepted_ppl_diff: float,
    decompose_in_float64: bool = True,
    u_matrix: Optional[torch.Tensor] = None,
    trace: Optional[list] = None,
    pair_state: Optional[_wrap.PairState] = None,
    group: Any = None,
) -> dict[str, Any]:
    """D:333-537."""
    indent = "    "
    target = root_module.get_submodule(decomposed_submodule_name)
    orig_device = target.weight.device
    orig_dtype = target.weight.dtype
    decomposed_type = utils.get_type_name(target)
    _wrap_in_place(root_module, decomposed_submodule_name)
    wrapper = root_module.get_submodule(decomposed_submodule_name)
    assert isinstance(wrapper, WrappedDWAINModule)
    orig_weight = wrapper.get_weight_copy()
    nat.require_cuda(orig_weight, f"weight of {decomposed_submodule_name}")
    dim_out, dim_in = orig_weight.shape
    full_rank = min(dim_in, dim_out)
    msg_prefix = f"Processing {decomposed_submodule_name}:"

    if full_rank == 1:
        _unwrap_in_place(root_module, decomposed_submodule_name)
        logger.info(f"{msg_prefix} Module has rank 1, not decomposing")
        return {"proportion": 1.0, "nsr_final": 0.0, "ppl_final": 0.0, "decomposed_module": None}

    logger.info(f"{msg_prefix} {decomposed_type} weight_shape={tuple(orig_weight.shape)} "
                f"{orig_weight.dtype}")
    logger.info(f"{msg_prefix} {nsr_final_threshold=:.4f} {max_accepted_ppl_diff=:.4f}")

    if u_matrix is not None:
        logger.info(f"Using pre-computed u_matrix, {u_matrix.dtype=}")
    else:
        u_matrix = _compute_covariance_matrix_decomposition(
            root_module=root_module, decomposed_submodule_name=decomposed_submodule_name,
            data_iterator=data_iterator, weight=orig_weight, num_data_steps=num_data_steps,
            device=device, decompose_in_float64=decompose_in_float64,
            num_vectors=_max_rank_consumed(dim_in, dim_out, reduction_factor))
        logger.info(f"Computed u_matrix, {u_matrix.dtype=}")

    if pair_state is not None:
        pair_state.begin_layer()

    def factors(rank: int) -> tuple[torch.Tensor, torch.Tensor]:
        """uk [out, k] and W1 = uk^T W [k, in], both in the weight dtype like D:423-428."""
        uk = u_matrix[:, u_matrix.shape[1] - rank:].to(orig_dtype).to(device)
        return uk, linalg.factor_w1(orig_weight, uk)

    # ---- the candidate ranks are data independent (D:407-421): enumerate them first ...
    plan: list[tuple[int, int, float, float]] = []
    rank_new = full_rank
    skipped: list[int] = []
    while rank_new > min_rank:
        rank_new = int(rank_new * reduction_factor)
        previous_params = _get_params_for_proportion(1.0, dim_in, dim_out)
        current_params = _get_params_for_proportion(rank_new / full_rank, dim_in, dim_out)
        drop_in_params = previous_params - current_params
        fraction_removed = drop_in_params / num_params
        if drop_in_params == 0:
            skipped.append(rank_new)
            continue
        plan.append((rank_new, drop_in_params, fraction_removed, fraction_removed * trade_off_factor))
    for rn in skipped:
        logger.info(f"{indent}{rn=} does not lead to params drop, skipping")

    # ---- ... evaluate every candidate (the reference evaluates all of them too: the smallest
    # accepted rank wins even when a larger one was rejected). The model does not change inside
    # one layer's search, so with a process group the (trial, metric batch) pairs are dealt
    # round-robin to the ranks; every rank draws every metric batch, keeping iterator positions
    # those of the reference, and the per-trial sums meet in one all-reduce.
    my_rank, world = parallel.rank_and_world(group)
    # one slot per (trial, batch): the all-reduce only adds zeros to each value and the batch sum
    # is taken in the same order on any number of GPUs, so the measured metrics -- and the ranks
    # chosen from them -- do not depend on how the jobs were dealt
    slots = torch.zeros((max(1, len(plan)), num_metric_steps, 3), dtype=torch.float64, device=orig_device)
    job = 0
    for t, (rank_t, _, _, _) in enumerate(plan):
        batches = [next(metric_iterator) for _ in range(num_metric_steps)]
        mine = [(j - job, b) for j, b in enumerate(batches, start=job) if j % world == my_rank]
        job += num_metric_steps
        if not mine:
            continue
        # no K5 GEMM (D:429) and no weight copy: the wrapper runs the two-factor op for the trial
        uk, w1 = factors(rank_t)
        for b_idx, batch in mine:
            input_dict = utils.to_device(batch, device)
            nsr_sample, ppl_deco_sample, ppl_orig_sample = _compute_metrics(
                input_dict=input_dict, root_module=root_module, decomposed_submodule=wrapper,
                factors=(w1, uk), loss_fn=loss_fn, pair_state=pair_state)
            ppl_diff_sample = (ppl_deco_sample - ppl_orig_sample) / ppl_orig_sample
            slots[t, b_idx] = torch.stack([ppl_diff_sample.double(), nsr_sample.double(),
                                           ppl_deco_sample.double()])
    if group is not None and len(plan) > 0:
        torch.distributed.all_reduce(slots, group=group)  # every (trial, batch) lives on one rank
    results = slots.sum(dim=1) / num_metric_steps
    measured = results.tolist() if len(plan) > 0 else []  # the one host sync of the layer

    # ---- ... then apply the accept / reject rules in order (D:445-487)
    tried = len(plan) > 0
    rank_best = full_rank
    nsr_best, ppl_deco_best = 0.0, 0.0
    for i, ((rank_new, drop_in_params, fraction_removed, ppl_diff_threshold),
            (ppl_diff_new, nsr_new, ppl_deco_new)) in enumerate(zip(plan, measured), start=1):
        logger.info(f"{indent}{i=} {ppl_deco_new=:.4f} {ppl_diff_new=:.4f} "
                    f"{ppl_diff_threshold=:.4f} {fraction_removed=:.4f} {nsr_new=:.4f}")
        reject = f"{indent}{i=} REJECTING rank {rank_new}/{full_rank}"
        accepted = False
        if ppl_diff_new >= ppl_diff_threshold:
            logger.info(f"{reject} {ppl_diff_new=:.2f} >= {ppl_diff_threshold=:.2f}")
        elif ppl_diff_new >= max_accepted_ppl_diff:
            logger.info(f"{reject} {ppl_diff_new=:.3f} >= {max_accepted_ppl_diff:.3f}")
        elif nsr_new >= nsr_final_threshold:
            logger.info(f"{reject} {nsr_new=:.4f} >= {nsr_final_threshold=:.4f}")
        else:
            accepted = True
            rank_best, nsr_best, ppl_deco_best = rank_new, nsr_new, ppl_deco_new
            logger.info(f"{indent}{i=} ACCEPTING rank {rank_best}/{full_rank}")
        if trace is not None:
            trace.append({"name": decomposed_submodule_name, "rank": rank_new, "nsr": nsr_new,
                          "ppl_diff": ppl_diff_new, "ppl_deco": ppl_deco_new,
                          "ppl_diff_threshold": ppl_diff_threshold, "accepted": accepted})
        logger.info(f"{indent}{i=} {rank_new=}/{full_rank} {nsr_new=:.6f} {ppl_diff_new=:.6f}  "
                    f"{rank_best=} {nsr_best=:.6f} {ppl_deco_best=:.6f}")
        logger.info(f"{indent}---")

    decompose_decision = False
    proportion = 1.0
    if tried:
        proportion = rank_best / full_rank
        logger.info(f"{indent}i=FINAL rank={rank_best}/{full_rank} {proportion=:.4f} "
                    f"nsr={nsr_best:.6f} ppl={ppl_deco_best:.6f}")
        decompose_decision = _is_num_params_reduced(proportion, dim_in, dim_out)
        if not decompose_decision:
            logger.info(f"{indent}{proportion=:.4f} leads to num param increase, not decomposing")

    if tried and full_rank != rank_best and decompose_decision:
        uk, w1 = factors(rank_best)  # rebuilt at rank_best (D:507-511)
        new_module = wrapper.get_decomposed_module(u=w1, v=uk)
        new_module.to(orig_device)
        new_module.to(orig_dtype)
        drop_in_params = (_get_params_for_proportion(1.0, dim_in, dim_out)
                          - _get_params_for_proportion(proportion, dim_in, dim_out))
    else:
        proportion, nsr_best, ppl_deco_best, drop_in_params, new_module = 1.0, 0.0, 0.0, 0, None
        logger.info(f"{msg_prefix} Skipping module decomposition")
        _unwrap_in_place(root_module, decomposed_submodule_name)

    return {"proportion": proportion, "nsr_final": nsr_best, "ppl_final": ppl_deco_best,
            "drop_in_params": drop_in_params, "decomposed_module": new_module}


def _compute_metrics(
    *,
    input_dict: dict[str, torch.Tensor],
    root_module: torch.nn.Module,
    decomposed_submodule: torch.nn.Module,
    factors: tuple[torch.Tensor, torch.Tensor],
    loss_fn: collections.abc.Callable[[dict[str, torch.Tensor], torch.Tensor], torch.Tensor],
    pair_state: Optional[_wrap.PairState] = None,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """D:247-278 with the layer evaluated through the trial factors (W1 [k, in], uk [out, k])
    instead of a copied-in effective weight. With a `pair_state` that verified this layer the two
    forwards share one pass over the doubled batch (see _wrap.PairState); the losses are still
    taken per variant on the original batch."""
    assert isinstance(input_dict, dict)
    assert isinstance(decomposed_submodule, WrappedDWAINModule)
    root_module.eval()
    if pair_state is not None:
        y_deco, y_orig = pair_state.forward_pair(root_module, decomposed_submodule, input_dict, factors)
    else:
        decomposed_submodule.set_trial(*factors)
        try:
            y_deco = root_module(input_dict)
        finally:
            decomposed_submodule.clear_trial()
        y_orig = root_module(input_dict)
    loss_deco = loss_fn(input_dict, y_deco)
    loss_orig = loss_fn(input_dict, y_orig)
    nsr_final = utils.calc_per_channel_noise_to_signal_ratio(
        y=y_orig, x=y_deco, non_channel_dim=(0, 1), mode="mean")
    ppl_deco = torch.exp(loss_deco).mean()
    ppl_orig = torch.exp(loss_orig).mean()
    return nsr_final, ppl_deco, ppl_orig


def _get_params_for_proportion(proportion: float, in_features: int, out_features: int) -> int:
    """D:319-330 (int() truncation included)."""
    baseline = in_features * out_features
    original_rank = min(in_features, out_features)
    proposed = (in_features + out_features) * proportion * original_rank
    return int(proposed) if proposed < baseline else baseline


def _precompute_covariance_matrix_decompositions_in_splits(
    *,
    module: torch.nn.Module,
    modules_to_decompose: list[str],
    num_splits: int,
    num_data_steps: int,
    data_iterator: collections.abc.Iterator[dict[str, torch.Tensor]],
    device: torch.device,
    decompose_in_float64: bool,
    reduction_factor: float = 1.0,
    group=None,
) -> dict[str, torch.Tensor]:
    """D:636-674: chunks of len // num_splits names, plus one more partition for the remainder."""
    chunk_size = len(modules_to_decompose) // num_splits
    if chunk_size == 0:
        chunk_size = 1
        num_splits = len(modules_to_decompose)
    num_partitions = num_splits if len(modules_to_decompose) % num_splits == 0 else num_splits + 1
    u_dict: dict[str, torch.Tensor] = {}
    for p in range(num_partitions):
        sublist = modules_to_decompose[p * chunk_size:(p + 1) * chunk_size]
        logger.info(f"Pre computing covariance matrices for {len(sublist)} modules")
        u_dict.update(_precompute_covariance_matrix_decompositions(
            module=module, submodule_names=sublist, num_data_steps=num_data_steps,
            data_iterator=data_iterator, device=device, decompose_in_float64=decompose_in_float64,
            reduction_factor=reduction_factor, group=group))
    assert len(u_dict) == len(modules_to_decompose)
    return u_dict


def _precompute_covariance_matrix_decompositions(
    *,
    module: torch.nn.Module,
    submodule_names: list[str],
    num_data_steps: int,
    data_iterator: collections.abc.Iterator[dict[str, torch.Tensor]],
    device: torch.device,
    decompose_in_float64: bool,
    reduction_factor: float = 1.0,
    group=None,
) -> dict[str, torch.Tensor]:
    """D:580-633: one set of num_data_steps forwards updates the covariances of every listed
    Linear at once. Targets that read the SAME input tensor (q/k/v, gate/up) share one input-side
    accumulator (see CovarianceUnits). With a process group, rank r only runs the steps i = r
    (mod world) -- every rank still draws all num_data_steps batches, so iterator positions stay
    those of the reference -- the partial covariances are summed over NVLink (lower triangles,
    asynchronously) and the eigensolves are distributed round-robin over the accumulators, each
    owner broadcasting its top-k blocks while the other ranks are still solving theirs."""
    rank, world = parallel.rank_and_world(group)
    # deterministic mode: canonical shards make the covariance bits independent of the GPU count
    originals = _install_covariance_modules(
        module, submodule_names, decompose_in_float64, reduction_factor, share_inputs=True,
        shard_plan=(linalg.canonical_shards(world), rank, world))
    units = module.get_submodule(submodule_names[0]).units if submodule_names else None

    module.eval()
    first_batch = None
    ran = 0
    with torch.no_grad():
        for step in range(num_data_steps):
            batch = next(data_iterator)
            if step == 0:
                first_batch = batch
                parallel.check_identical_batches(batch, group)
            if step % world != rank:
                continue
            if units is not None:
                units.step = step
            _ = module(utils.to_device(batch, device))
            ran += 1
        if units is not None:
            if ran == 0 and first_batch is not None:
                # fewer steps than ranks: this rank still needs the layout of the accumulators
                units.dry = True
                _ = module(utils.to_device(first_batch, device))
            units.finish_probe()
    utils.relieve_gpu_memory_pressure()

    logger.info("Computing eigenvectors ...")
    u_dict: dict[str, torch.Tensor] = {}
    if units is not None:
        ks = {id(m): _max_rank_consumed(m.in_features, m.out_features, reduction_factor)
              for u in units.units for m in u.members}
        jobs = []
        for u in units.units:
            sizes = [(m.out_features, ks[id(m)]) for m in u.members]
            total = sum(a * b for a, b in sizes)
            jobs.append((u.acc, (lambda u=u: torch.cat(
                [m.get_eigenvectors(ks[id(m)]).reshape(-1) for m in u.members])), (total,)))
        flats = parallel.owners_compute_pipelined(jobs, group, total_steps=num_data_steps,
                                                  costs=[_unit_cost(u, ks) for u in units.units])
        by_module = {}
        for u, flat in zip(units.units, flats):
            off = 0
            for m in u.members:
                n_el = m.out_features * ks[id(m)]
                by_module[id(m)] = flat[off:off + n_el].reshape(m.out_features, ks[id(m)])
                off += n_el
        for name in submodule_names:
            u_dict[name] = by_module[id(module.get_submodule(name))]
    _restore_modules(module, originals)
    utils.relieve_gpu_memory_pressure()
    return u_dict


def _unit_cost(unit, ks: dict) -> float:
    """Relative eigensolve cost of one accumulator (d^3 per eigensolve) for balancing its owners."""
    if unit.kind == "output":
        return float(unit.members[0].out_features) ** 3
    cost, shared_eig = 0.0, False
    for m in unit.members:
        if linalg.use_input_side(m.in_features, m.out_features, ks[id(m)]):
            cost += float(m.in_features) ** 3  # eigh of the in x in Gram matrix (+ its GEMMs)
            shared_eig = True
        else:
            cost += float(m.out_features) ** 3 + 0.1 * float(m.in_features) ** 3  # C = W S W^T, eigh(C)
    return cost + (float(unit.members[0].in_features) ** 3 if shared_eig else 0.0)


def _install_covariance_modules(module: torch.nn.Module, submodule_names: list[str],
                                decompose_in_float64: bool,
                                reduction_factor: Optional[float] = None,
                                share_inputs: bool = True,
                                shard_plan: tuple[int, int, int] = (1, 0, 1)) -> dict[str, torch.nn.Module]:
    """D:592-603: swap every listed Linear for a CovarianceComputingLinearModule that shares its
    weight and bias. Returns the originals for _restore_modules. With `share_inputs` the modules
    of this call form one CovarianceUnits registry (reachable as `<module>.units`)."""
    originals: dict[str, torch.nn.Module] = {}
    units = CovarianceUnits(shard_plan) if share_inputs else None
    for name in submodule_names:
        old = module.get_submodule(name)
        if not isinstance(old, torch.nn.Linear):
            # the reference crashes here on 1x1 convs (D:194 with a 4-D weight)
            raise ValueError(f"covariance precompute supports Linear targets only, got {name}={old}")
        originals[name] = old
        logger.info(f"Replacing {name} by covariance computing wrapper")
        k = (None if reduction_factor is None
             else _max_rank_consumed(old.in_features, old.out_features, reduction_factor))
        utils.replace_submodule_in_place(
            module, name,
            CovarianceComputingLinearModule(old.weight, old.bias, decompose_in_float64, k, units=units))
    return originals


def _restore_modules(module: torch.nn.Module, originals: dict[str, torch.nn.Module]) -> None:
    """D:624-630."""
    for name, old in originals.items():
        logger.info(f"Replacing {name} by original linear")
        utils.replace_submodule_in_place(module, name, old)


class CovarianceUnit:
    """One covariance accumulator and the covariance modules that read it: kind "output" (C = E[y y^T]
    of one module, the reference formulation) or "input" (S = E[x x^T] of the tensor the members
    share; a member's C = W S W^T follows without touching the activations again)."""

    def __init__(self, kind: str, acc: linalg.CovarianceAccumulator, members: list) -> None:
        self.kind = kind
        self.acc = acc
        self.members = members
        self.last_call = -1
        self.last_x: Optional[torch.Tensor] = None
        self._s_final: Optional[torch.Tensor] = None
        self._s_eig: Optional[tuple[torch.Tensor, torch.Tensor]] = None

    def finalized_input_covariance(self) -> torch.Tensor:
        if self._s_final is None:
            # centring / damping act on C, not on S (damping is a multiple of I on C)
            self._s_final = self.acc.finalize(use_mean=False, damp_factor=0.0)
        return self._s_final

    def input_eigensystem(self) -> tuple[torch.Tensor, torch.Tensor]:
        """eigh(S), computed once for all members that take the in x in route (gate and up)."""
        if self._s_eig is None:
            self._s_eig = linalg.eigh(self.finalized_input_covariance())
        return self._s_eig


class CovarianceUnits:
    """Which covariance modules of one calibration split see the SAME input tensor.

    The reference gives every target its own d_out x d_out accumulator (D:166-208); q/k/v and
    gate/up of a decoder read one tensor each, so their input covariance S = E[x x^T] is the same
    matrix accumulated two or three times. The first forward is a probe: every module records the
    tensor OBJECT it received; modules that received the same object form one unit with a single
    input-side accumulator (Llama-3-8B decoder layer: 4 SYRK launches per step instead of 7, one of
    them shared by q/k/v and one by gate/up). Later forwards check the identity again and fail
    loudly if the model's dataflow changed. A module called twice within one forward (weight
    sharing) switches the whole split back to private accumulators."""

    def __init__(self, shard_plan: tuple[int, int, int] = (1, 0, 1)) -> None:
        self.shard_plan = shard_plan  # (canonical shards, rank, world) of the accumulators
        self.step: Optional[int] = None        # calibration step of the running forward (driver-set)
        self.probe_step: Optional[int] = None  # ... of the probe forward
        self.probing = True
        self.dry = False            # probe for the layout only, fold nothing in
        self.records: list = []     # (module, x, rows, y_rows) of the probe forward, in call order
        self.units: list[CovarianceUnit] = []
        self.calls = 0              # forward number, counted at the first module of the probe order
        self.first_module = None
        self.multi_call = False

    def on_forward(self, mod, x: torch.Tensor, rows: torch.Tensor, y_rows: torch.Tensor) -> None:
        if self.probing:
            seen = any(r[0] is mod for r in self.records)
            if self.first_module is None:
                self.first_module = mod
            elif mod is self.first_module:
                self.finish_probe()  # the first module again: the probe forward is over
            elif seen:
                self.multi_call = True  # weight sharing inside the model: no input sharing
            if self.probing:
                if not self.records:
                    self.probe_step = self.step
                self.records.append((mod, x, rows, y_rows))
                return
        if mod is self.first_module:
            self.calls += 1
        unit = mod.unit
        if unit is None:  # not seen by the probe (it ended early on a re-entrant first module)
            unit = mod.unit = self._private_unit(mod)
            self.units.append(unit)
        if unit.kind == "output":
            _update_Eyyt_in_place(unit.acc, y_rows, step=self.step)
        elif len(unit.members) == 1:
            _update_Eyyt_in_place(unit.acc, rows, step=self.step)
        elif unit.last_call != self.calls:
            _update_Eyyt_in_place(unit.acc, rows, step=self.step)
            unit.last_call, unit.last_x = self.calls, x
        elif x is not unit.last_x:
            raise RuntimeError(
                "ptdeco_b200: modules that shared one input tensor in the first calibration forward "
                f"received different tensors later ({len(unit.members)} members); "
                "set PTDECO_B200_SHARE_INPUTS=0")

    def _accumulator(self, d: int, like: torch.Tensor) -> linalg.CovarianceAccumulator:
        shards, rank, world = self.shard_plan
        return linalg.CovarianceAccumulator(
            d, like.device, defer_rows=linalg.default_defer_rows(d, like.element_size()),
            shards=shards, rank=rank, world=world)

    def _private_unit(self, mod) -> CovarianceUnit:
        kind, d = ("input", mod.in_features) if mod.input_side else ("output", mod.out_features)
        return CovarianceUnit(kind, self._accumulator(d, mod.weight), [mod])

    def finish_probe(self) -> None:
        """Turn the probe forward's records into units and fold the probe batch in."""
        if not self.probing:
            return
        self.probing = False
        share = os.environ.get("PTDECO_B200_SHARE_INPUTS", "1") != "0" and not self.multi_call
        groups: dict[int, list] = {}
        for rec in self.records:
            groups.setdefault(id(rec[1]) if share else id(rec[0]), []).append(rec)
        for recs in groups.values():
            mods = [r[0] for r in recs]
            in_f = mods[0].in_features
            shared = (share and len(recs) >= 2 and all(m.in_features == in_f for m in mods)
                      and in_f <= 2 * max(m.out_features for m in mods))
            if shared:  # one S = E[x x^T] for every module that read this tensor
                acc = self._accumulator(in_f, mods[0].weight)
                unit = CovarianceUnit("input", acc, mods)
                for m in mods:
                    m.unit = unit
                if not self.dry:
                    _update_Eyyt_in_place(acc, recs[0][2], step=self.probe_step)
                self.units.append(unit)
                continue
            for rec in recs:  # private accumulators (one record per call of the module)
                m = rec[0]
                if m.unit is None:
                    m.unit = self._private_unit(m)
                    self.units.append(m.unit)
                if not self.dry:
                    _update_Eyyt_in_place(m.unit.acc, rec[3] if m.unit.kind == "output" else rec[2],
                                          step=self.probe_step)
        self.records = []
        self.calls = 1


class CovarianceComputingLinearModule(torch.nn.Module):
    """D:166-208: stands in for a target Linear during the precompute pass; its forward IS the
    layer forward (y = x W^T on the tcgen05 GEMM engine) and folds the layer's activations into
    a covariance: the output y like the reference, or -- when in < out and only eigenvectors in
    range(W) are wanted, or when several targets read the same tensor (CovarianceUnits) -- the
    input x (C = W S W^T)."""

    def __init__(self, weight: torch.nn.Parameter, bias: Optional[torch.nn.Parameter],
                 decompose_in_float64: bool, num_vectors: Optional[int] = None,
                 units: Optional[CovarianceUnits] = None):
        super().__init__()
        self.weight = weight
        self.bias = bias
        self.in_features = weight.shape[1]
        self.out_features = weight.shape[0]
        self.input_side = linalg.use_input_side(self.in_features, self.out_features, num_vectors)
        self.units = units
        self.unit: Optional[CovarianceUnit] = None
        if units is None:  # stand-alone module: its own accumulator, decided now
            d = self.in_features if self.input_side else self.out_features
            self.unit = CovarianceUnit(
                "input" if self.input_side else "output",
                linalg.CovarianceAccumulator(
                    d, weight.device, defer_rows=linalg.default_defer_rows(d, weight.element_size())),
                [self])
        self.use_float64 = decompose_in_float64  # accepted; see falor's use_float64 note

    @property
    def acc(self) -> linalg.CovarianceAccumulator:
        if self.unit is None:
            self.units.finish_probe()
        return self.unit.acc

    @property
    def num_data_steps(self) -> int:
        return self.acc.steps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        rows = x.reshape(-1, self.in_features)
        y_rows = linalg.linear_nt(rows, self.weight.detach())
        if self.units is not None:
            self.units.on_forward(self, x, rows, y_rows)
        else:
            _update_Eyyt_in_place(self.unit.acc, rows if self.unit.kind == "input" else y_rows)
        y = y_rows.reshape(*x.shape[:-1], self.out_features)
        if self.bias is not None:
            y = y + self.bias
        return y

    def get_eigenvectors(self, num_vectors: Optional[int] = None, group=None) -> torch.Tensor:
        """Unlike D:206-208 the result stays on the GPU in fp32 (180 GB of HBM make the reference's
        round trip through host memory unnecessary); the rank search casts what it slices."""
        if self.unit is None:
            self.units.finish_probe()
        unit = self.unit
        if unit.kind == "output":
            return _get_eigenvectors(unit.acc, num_vectors, group)
        if group is not None:
            parallel.allreduce_accumulator(unit.acc, group)
        s_cov = unit.finalized_input_covariance()
        weight = self.weight.detach()
        if linalg.use_input_side(self.in_features, self.out_features, num_vectors):
            return linalg.eigvecs_from_input_covariance(s_cov, weight, num_vectors,
                                                        s_eig=unit.input_eigensystem())
        # out <= in (q/k/v/o of a shared tensor): the reference's own out x out problem, with
        # C = W S W^T formed by two tensor-core GEMMs instead of a SYRK over the activations
        cov = linalg.covariance_from_input(s_cov, weight, damp_factor=EIGEN_DAMPEN_FACTOR)
        _, u = linalg.eigh(cov, k=num_vectors)
        return u


def _compute_covariance_matrix_decomposition(
    *,
    root_module: torch.nn.Module,
    decomposed_submodule_name: str,
    data_iterator: collections.abc.Iterator[dict[str, torch.Tensor]],
    weight: torch.Tensor,
    num_data_steps: int,
    device: torch.device,
    decompose_in_float64: bool,
    num_vectors: Optional[int] = None,
) -> torch.Tensor:
    """D:211-244: per-layer calibration (num_data_steps full forwards) -> eigenvectors."""
    root_module.eval()
    wrapper = root_module.get_submodule(decomposed_submodule_name)
    assert isinstance(wrapper, WrappedDWAINModule)
    logger.info("Using float64 for decomposition" if decompose_in_float64
                else "Using float32 for decomposition")
    input_side = linalg.use_input_side(weight.shape[1], weight.shape[0], num_vectors)
    d = weight.shape[1] if input_side else weight.shape[0]
    acc = linalg.CovarianceAccumulator(
        d, device, defer_rows=linalg.default_defer_rows(d, weight.element_size()))
    # the hooked layer output stands in for y = x W^T only when it has one row per input position
    from_output = not input_side and wrapper.output_covers_input_positions()
    wrapper.capture_output = from_output
    try:
        for _ in range(num_data_steps):
            inputs = utils.to_device(next(data_iterator), device)
            # D:237 discards the model output: the forward stops right after the target layer
            _wrap.calibration_forward(root_module, inputs, wrapper)
            if input_side:
                _update_Eyyt_in_place(acc, wrapper.get_last_input())
            elif from_output:
                _update_Eyyt_in_place(acc, wrapper.get_last_output_rows(), sub=wrapper.get_bias())
            else:  # strided / padded 1x1 conv: the reference's y = x W^T over ALL input positions (D:239)
                _update_Eyyt_in_place(acc, linalg.linear_nt(wrapper.get_last_input(), weight))
    finally:
        wrapper.capture_output = False
        wrapper.output = None
    if input_side:
        return _get_eigenvectors_input_side(acc, weight, num_vectors)
    return _get_eigenvectors(acc, num_vectors)


def _get_eigenvectors_input_side(acc: linalg.CovarianceAccumulator, weight: torch.Tensor,
                                 num_vectors: int, group=None) -> torch.Tensor:
    parallel.allreduce_accumulator(acc, group)
    s_cov = acc.finalize(use_mean=False, damp_factor=0.0)
    return linalg.eigvecs_from_input_covariance(s_cov, weight, num_vectors)


def _get_eigenvectors(acc: linalg.CovarianceAccumulator, num_vectors: Optional[int] = None,
                      group=None) -> torch.Tensor:
    """D:155-163 on the accumulated covariance: /steps, damping 0.01*mean(diag), eigenvectors
    ascending (all, or the last `num_vectors`). With a process group the partial covariances are
    summed over ranks first."""
    parallel.allreduce_accumulator(acc, group)
    cov = acc.finalize(use_mean=False, damp_factor=EIGEN_DAMPEN_FACTOR)
    _, u = linalg.eigh(cov, k=num_vectors)
    return u


def _update_Eyyt_in_place(acc: linalg.CovarianceAccumulator, y_reshaped: torch.Tensor,
                          sub: Optional[torch.Tensor] = None, step: Optional[int] = None) -> None:
    """D:147-152: Eyyt += y^T y / N for one batch of rows."""
    acc.update(y_reshaped, sub=sub, step=step)


def _max_rank_consumed(dim_in: int, dim_out: int, reduction_factor: float) -> int:
    """Largest rank the descent of D:407-408 can ask for: int(full_rank * reduction_factor)."""
    full_rank = min(dim_in, dim_out)
    return max(1, min(full_rank, int(full_rank * reduction_factor)))


def _wrap_in_place(root_module: torch.nn.Module, decomposed_submodule_name: str) -> None:
    """D:281-304."""
    sub = root_module.get_submodule(decomposed_submodule_name)
    if isinstance(sub, torch.nn.Linear):
        wrapped: WrappedDWAINModule = WrappedDWAINLinear(sub, decomposed_submodule_name)
    elif is_decomposeable_module(sub):
        wrapped = WrappedDWAINConv2d1x1(sub, decomposed_submodule_name)
    else:
        raise ValueError(f"Cannot decompose {decomposed_submodule_name}={sub}")
    utils.replace_submodule_in_place(root_module, decomposed_submodule_name, wrapped)


def _unwrap_in_place(root_module: torch.nn.Module, decomposed_submodule_name: str) -> None:
    """D:307-316."""
    sub = root_module.get_submodule(decomposed_submodule_name)
    assert isinstance(sub, WrappedDWAINModule)
    utils.replace_submodule_in_place(root_module, decomposed_submodule_name, sub.get_orig_module())


def _get_decomposeable_submodule_names(module: torch.nn.Module,
                                       blacklisted_module_names: list[str]) -> list[str]:
    """D:549-559."""
    res = []
    for name, mod in module.named_modules():
        if is_decomposeable_module(mod):
            if name in blacklisted_module_names:
                logger.info(f"Skipping blacklisted module {name}")
            else:
                res.append(name)
    return res


def _add_meta_to_module_config(module_config: dict[str, Any], module_deco_results: dict[str, Any]) -> None:
    """D:562-566."""
    module_config[utils.MODCONFIG_META_KEY] = {
        k: v for k, v in module_deco_results.items() if k != "decomposed_module"}
