"""falor (Features Are LOw-Rank) on the sm_100a kernels.

Drop-in for the reference's src/ptdeco/falor/decomposition.py (cited as F:line): same public
entry points (`decompose_in_place`, `is_decomposeable_module`), same private primitives the
reference's tests import (`_wrap_in_place`, `_unwrap_in_place`,
`_compute_decompositon_of_covariance_matrix`), same iterator consumption order, same rank search
and quirks, same `decompose_config`. What changes is who does the arithmetic:

  F:159-161  y = x W^T ; Eyyt += y^T y / N ; Ey += mean(y)  -> hooked layer output fed to the
                                                               tcgen05 SYRK (ptdeco_syrk_accumulate)
  F:192-205  /steps, centring, damping                      -> ptdeco_cov_finalize
  F:207      torch.linalg.eigh                              -> ptdeco_eigh (top-k back-transform)
  F:347-348  U = W^T uk ; deco_weight = (U V)^T             -> ptdeco_gemm (tcgen05, bf16x3 split)
  F:228-232  NSR / symmetric KL                             -> ptdeco_nsr_metric / ptdeco_kl_metric

The model forward passes stay the user's torch code. CUDA only: there is no CPU path.
"""
from __future__ import annotations

import collections
import collections.abc
import logging
import time
from typing import Any, Optional

import torch

from .. import _native as nat
from .. import _wrap, linalg, parallel, utils

EIGEN_DAMPEN_FACTOR = 0.01  # F:22

logger = logging.getLogger("ptdeco.falor.decomposition")

__all__ = ["decompose_in_place", "is_decomposeable_module"]


class WrappedFALORModule(_wrap.WrappedModule):
    pass


class WrappedFALORLinear(_wrap.WrappedLinear, WrappedFALORModule):
    pass


class WrappedFALORConv2d1x1(_wrap.WrappedConv2d1x1, WrappedFALORModule):
    pass


is_decomposeable_module = _wrap.is_decomposeable_module
_is_num_params_reduced = _wrap.is_num_params_reduced


def decompose_in_place(
    *,
    module: torch.nn.Module,
    device: torch.device,
    data_iterator: collections.abc.Iterator[torch.Tensor],
    blacklisted_module_names: Optional[list[str]] = None,
    proportion_threshold: float,
    nsr_final_threshold: float,
    kl_final_threshold: float,
    num_data_steps: int,
    num_metric_steps: int,
    use_float64: bool,
    use_mean: bool,
    use_damping: bool,
    trace: Optional[list] = None,
    process_group: Any = None,
) -> dict[str, Any]:
    """F:424-511. Every target is analysed against the unmodified model; swaps happen afterwards.
    Returns the decompose_config (insertion order = forward module order).

    `process_group` (not in the reference) opts into the multi-GPU path: falor's layers are
    independent of each other (all are analysed against the ORIGINAL model, F:451-471), so with a
    group (or "world") layer number l is analysed by rank l mod world and the others only draw --
    and drop -- that layer's batches, whose count is data independent (`num_data_steps` plus
    `num_metric_steps` per bisection step), so every rank's iterator stays at the reference's
    position. The owners then broadcast their results and every rank installs identical modules.
    Requires identical replicas and iterators on all ranks (checked by a batch checksum)."""
    start_time = time.perf_counter()
    device = torch.device(device)
    if device.type != "cuda":
        raise nat.NativeError("ptdeco_b200.falor runs on CUDA (sm_100a) only; there is no CPU path")
    nat.lib()  # fail now, loudly, if the native library is missing
    results_all: dict[str, dict[str, Any]] = {}
    decompose_config: dict[str, Any] = {}
    if blacklisted_module_names is None:
        blacklisted_module_names = []

    names = _get_decomposeable_submodule_names(module)
    n = len(names)
    module.eval()
    pair_state = _wrap.PairState(module)
    group = parallel.resolve_group(process_group)
    my_rank, world = parallel.rank_and_world(group)
    owners: dict[str, int] = {}
    n_live = 0
    for i, name in enumerate(names, start=1):
        msg_prefix = f"Processing {name}: module {i} of {n}"
        if name in blacklisted_module_names:
            logger.info(f"{msg_prefix}, skipped as blacklisted")
            continue
        owners[name] = n_live % world
        n_live += 1
        if owners[name] != my_rank:
            # another rank analyses this layer: consume exactly the batches it consumes
            skipped = _batches_consumed_by_layer(module.get_submodule(name), num_data_steps, num_metric_steps)
            for _ in range(skipped):
                next(data_iterator)
            logger.info(f"{msg_prefix}, analysed by rank {owners[name]} ({skipped} batches skipped)")
            continue
        logger.info(msg_prefix)
        with torch.no_grad():
            results_all[name] = _process_module(
                root_module=module, decomposed_submodule_name=name, data_iterator=data_iterator,
                nsr_final_threshold=nsr_final_threshold, kl_final_threshold=kl_final_threshold,
                num_data_steps=num_data_steps, num_metric_steps=num_metric_steps, device=device,
                use_float64=use_float64, use_mean=use_mean, use_damping=use_damping, trace=trace,
                pair_state=pair_state)

    if group is not None:
        _exchange_layer_results(module, owners, results_all, my_rank, group, device, trace)

    counter: collections.Counter[str] = collections.Counter()
    for name in names:
        msg_prefix = f"Decomposing {name}:"
        if name in blacklisted_module_names:
            logger.info(f"{msg_prefix} SKIPPED blacklisted module {name}")
            continue
        result = results_all[name]
        new_module = result["decomposed_module"]
        proportion = result["proportion"]
        if new_module is None:
            logger.info(f"{msg_prefix} SKIPPED {proportion=:.4f} leads to num param increase")
            continue
        if proportion < proportion_threshold:
            old_type = utils.get_type_name(module.get_submodule(name))
            utils.replace_submodule_in_place(module, name, new_module)
            module_config = utils.get_module_config(new_module)
            _add_meta_to_module_config(module_config, result)
            decompose_config[name] = module_config
            counter[old_type] += 1
            logger.info(f"{msg_prefix} finished {proportion=:.3f}")
        else:
            logger.info(f"{msg_prefix} SKIPPED, {proportion=:.3f} above {proportion_threshold=:.3f}")

    for type_name, count in counter.items():
        logger.info(f"Decomposed {count} instances of {type_name}")
    logger.info(f"Total decomposable modules {n}")
    logger.info(f"Decomposition took {time.perf_counter() - start_time:.1f} seconds")
    return decompose_config


def _process_module(
    *,
    root_module: torch.nn.Module,
    decomposed_submodule_name: str,
    data_iterator: collections.abc.Iterator[torch.Tensor],
    nsr_final_threshold: float,
    kl_final_threshold: float,
    num_data_steps: int,
    num_metric_steps: int,
    device: torch.device,
    use_float64: bool,
    use_mean: bool,
    use_damping: bool,
    trace: Optional[list] = None,
    pair_state: Optional[_wrap.PairState] = None,
) -> dict[str, Any]:
    """F:284-399: covariance -> eigenvectors -> bisection on the rank. `trace` (not in the
    reference) collects one record per trial for the parity harness.

    A trial does not materialise the effective weight uk uk^T W (K5, F:348) nor copy it into the
    layer (F:222,226): the wrapper evaluates the layer as the two-factor op uk (W1 x) + b on the
    fused low-rank kernel (ptdeco_lowrank_forward), which is the same linear map."""
    decomposed_type = utils.get_type_name(root_module.get_submodule(decomposed_submodule_name))
    _wrap_in_place(root_module, decomposed_submodule_name)
    wrapper = root_module.get_submodule(decomposed_submodule_name)
    assert isinstance(wrapper, WrappedFALORModule)
    orig_weight = wrapper.get_weight_copy()
    nat.require_cuda(orig_weight, f"weight of {decomposed_submodule_name}")
    orig_device = orig_weight.device
    dim_out, dim_in = orig_weight.shape
    full_rank = min(dim_in, dim_out)
    msg_prefix = f"Processing {decomposed_submodule_name}:"

    if full_rank == 1:
        _unwrap_in_place(root_module, decomposed_submodule_name)
        logger.info(f"{msg_prefix} Module has rank 1, not decomposing")
        return {"proportion": 1.0, "nsr_final": 0.0, "kl_final": 0.0, "decomposed_module": None}

    logger.info(f"{msg_prefix} {decomposed_type} weight_shape={tuple(orig_weight.shape)}")
    logger.info(f"{msg_prefix} {nsr_final_threshold=:.6f} {kl_final_threshold=:.6f}")

    # The bisection asks for rank_best - rank_width; while every trial is rejected rank_best stays
    # full_rank and the width halves, so any rank up to full_rank - 1 can be requested (F:340-375).
    k_max = max(1, full_rank - 1)
    root_module.eval()
    if pair_state is not None:
        pair_state.begin_layer()
    u = _compute_decompositon_of_covariance_matrix(
        root_module=root_module, decomposed_submodule_name=decomposed_submodule_name,
        data_iterator=data_iterator, weight=orig_weight, num_data_steps=num_data_steps,
        device=device, use_float64=use_float64, use_mean=use_mean, use_damping=use_damping,
        num_vectors=k_max)

    w32 = orig_weight if orig_weight.dtype == torch.float32 else orig_weight.float()
    w1 = uk = None
    i = 1
    rank_best = full_rank
    rank_width = full_rank // 2
    nsr_best, kl_best = 0.0, 0.0
    nsr_new, kl_new = 0.0, 0.0
    while rank_width > 0:
        rank_new = rank_best - rank_width
        uk = u[:, u.shape[1] - rank_new:]  # [out, k], fp32 (F:346)
        w1 = linalg.factor_w1(w32, uk)  # = U^T, [k, in] (F:347)

        nsr_acc = torch.zeros((), dtype=torch.float64, device=orig_device)
        kl_acc = torch.zeros((), dtype=torch.float64, device=orig_device)
        for _ in range(num_metric_steps):
            x = next(data_iterator).to(device)
            nsr_sample, kl_sample = _compute_metrics(
                x=x, root_module=root_module, decomposed_submodule=wrapper,
                factors=(w1, uk), pair_state=pair_state)
            nsr_acc += nsr_sample.double()
            kl_acc += kl_sample.double()
        nsr_new, kl_new = (torch.stack([nsr_acc, kl_acc]) / num_metric_steps).tolist()  # one sync

        accepted = nsr_new < nsr_final_threshold and kl_new < kl_final_threshold
        if accepted:
            rank_best, nsr_best, kl_best = rank_new, nsr_new, kl_new
        if trace is not None:
            trace.append({"name": decomposed_submodule_name, "rank": rank_new, "nsr": nsr_new,
                          "kl": kl_new, "accepted": accepted})
        msg_iter = f"{i=} {rank_width=} {rank_new=} {nsr_new=:.6f} {kl_new=:.6f}"
        msg_cur = f"{rank_best=} {nsr_best=:.6f} {kl_best=:.6f}"
        logger.info(f"{msg_prefix} {msg_iter} {msg_cur}")
        rank_width //= 2
        i += 1
    assert w1 is not None and uk is not None

    proportion = rank_best / full_rank
    logger.info(f"{msg_prefix} iter=FINAL rank={rank_best} {proportion=:.4f} nsr={nsr_best:.6f} "
                f"kl={kl_new:.6f}")

    if full_rank != rank_best and _is_num_params_reduced(proportion, dim_in, dim_out):
        # Quirk kept from the reference: the module is built from the factors of the LAST tried
        # rank (F:346-348 run inside the loop, F:383-386 after it), while `proportion` reports
        # rank_best (F:379).
        new_module = wrapper.get_decomposed_module(u=w1, v=uk)
        # F:387 leaves the fp32 factors as they are; in a non-fp32 model that module would raise a
        # dtype mismatch at its first forward, so the factors follow the layer's dtype here
        new_module.to(device=orig_device, dtype=orig_weight.dtype)
    else:
        logger.info(f"{msg_prefix} {proportion=:.4f} leads to num param increase, not decomposing")
        new_module = None

    _unwrap_in_place(root_module, decomposed_submodule_name)
    return {"proportion": proportion, "nsr_final": nsr_new, "kl_final": kl_new,
            "decomposed_module": new_module}


def _batches_consumed_by_layer(target: torch.nn.Module, num_data_steps: int, num_metric_steps: int) -> int:
    """How many batches _process_module draws for this layer: data independent (F:188, F:340-375):
    num_data_steps for the covariance plus num_metric_steps per bisection step, and the number of
    steps depends only on full_rank (the width halves until it is 0); a rank-1 layer draws none."""
    w = target.weight
    full_rank = min(w.shape[0], w.shape[1])
    if full_rank == 1:
        return 0
    steps, width = 0, full_rank // 2
    while width > 0:
        steps += 1
        width //= 2
    return num_data_steps + steps * num_metric_steps


def _exchange_layer_results(module: torch.nn.Module, owners: dict[str, int], results_all: dict,
                            my_rank: int, group, device: torch.device, trace: Optional[list]) -> None:
    """Layer-sharded falor: every owner broadcasts its layer's outcome (scalars + the two factor
    matrices) so that all ranks build identical replacement modules; traces are merged in layer
    order on every rank."""
    import torch.distributed as dist
    per_layer_trace: dict[str, list] = {}
    if trace is not None:
        for t in trace:
            per_layer_trace.setdefault(t["name"], []).append(t)
        trace.clear()
    for name, owner in owners.items():
        src = dist.get_global_rank(group, owner)
        payload = [None]
        if owner == my_rank:
            r = results_all[name]
            new = r["decomposed_module"]
            payload[0] = {"proportion": r["proportion"], "nsr_final": r["nsr_final"], "kl_final": r["kl_final"],
                          "rank": None if new is None else int(new[0].weight.shape[0]),
                          "dtype": None if new is None else str(new[0].weight.dtype).replace("torch.", ""),
                          "trace": per_layer_trace.get(name, [])}
        dist.broadcast_object_list(payload, src=src, group=group)
        meta = payload[0]
        if trace is not None:
            trace.extend(meta["trace"])
        if owner == my_rank:
            new = results_all[name]["decomposed_module"]
        else:
            new = None
            if meta["rank"] is not None:
                _wrap_in_place(module, name)
                wrapper = module.get_submodule(name)
                w = wrapper.get_weight_copy()
                dt = getattr(torch, meta["dtype"])
                new = wrapper.get_decomposed_module(
                    u=torch.empty((meta["rank"], w.shape[1]), dtype=dt, device=w.device),
                    v=torch.empty((w.shape[0], meta["rank"]), dtype=dt, device=w.device))
                _unwrap_in_place(module, name)
            results_all[name] = {"proportion": meta["proportion"], "nsr_final": meta["nsr_final"],
                                 "kl_final": meta["kl_final"], "decomposed_module": new}
        if new is not None:
            for prm in (new[0].weight, new[1].weight):
                buf = prm.data.contiguous()
                dist.broadcast(buf, src=src, group=group)
                if owner != my_rank:
                    prm.data = buf


def _compute_metrics(
    *,
    x: torch.Tensor,
    root_module: torch.nn.Module,
    decomposed_submodule: torch.nn.Module,
    factors: tuple[torch.Tensor, torch.Tensor],
    pair_state: Optional[_wrap.PairState] = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """F:211-233: two full forwards (layer evaluated through the trial factors W1 [k, in],
    W2 = uk [out, k], then the original layer) -- or, once `pair_state` has verified it on this
    layer, one forward of the doubled batch (see _wrap.PairState) -- then NSR over the batch dim and
    symmetric-max KL of the logits; both returned as 0-dim device tensors (no host sync)."""
    assert isinstance(decomposed_submodule, WrappedFALORModule)
    root_module.eval()
    if pair_state is not None:
        y_deco, y_orig = pair_state.forward_pair(root_module, decomposed_submodule, x, factors)
    else:
        decomposed_submodule.set_trial(*factors)
        try:
            y_deco = root_module(x)
        finally:
            decomposed_submodule.clear_trial()
        y_orig = root_module(x)
    nsr_final = utils.calc_per_channel_noise_to_signal_ratio(y=y_orig, x=y_deco, non_channel_dim=(0,))
    kl_final = utils.calc_kl_loss(y_deco, y_orig)
    return nsr_final, kl_final


def _compute_decompositon_of_covariance_matrix(
    *,
    root_module: torch.nn.Module,
    decomposed_submodule_name: str,
    data_iterator: collections.abc.Iterator[torch.Tensor],
    weight: torch.Tensor,
    num_data_steps: int,
    device: torch.device,
    use_float64: bool,
    use_mean: bool,
    use_damping: bool,
    num_vectors: Optional[int] = None,
) -> torch.Tensor:
    """F:165-208. Returns eigenvectors in columns, ascending by eigenvalue: all `d` of them like
    torch.linalg.eigh, or only the last `num_vectors` (what the rank search consumes; slicing
    `u[:, u.shape[1] - rank:]` works on either).

    `use_float64`: the reference forms each per-batch product in the model dtype and only *adds*
    it into an fp64 accumulator (F:160,181-183), which measures 5.6e-8 vs 7.3e-8 against an fp64
    truth. Here products are exact-bf16 / bf16x3-split with fp32 accumulation and
    the tridiagonal eigenproblem is solved in fp64 either way, so the flag is accepted and has no
    further effect. The falor damping quirk is kept: damping reaches `cov` only when
    use_mean=False (F:196-205)."""
    root_module.eval()
    wrapper = root_module.get_submodule(decomposed_submodule_name)
    assert isinstance(wrapper, WrappedFALORModule)
    n_out, n_in = weight.shape
    input_side = linalg.use_input_side(n_in, n_out, num_vectors)
    d = n_in if input_side else n_out
    acc = linalg.CovarianceAccumulator(d, device, with_mean=use_mean,
                                       defer_rows=linalg.default_defer_rows(d, weight.element_size()))
    # the hooked layer output stands in for y = x W^T only when it has one row per input position
    from_output = not input_side and wrapper.output_covers_input_positions()
    wrapper.capture_output = from_output
    try:
        for _ in range(num_data_steps):
            inputs = next(data_iterator).to(device)
            # F:189 discards the model output: the forward stops right after the target layer
            _wrap.calibration_forward(root_module, inputs, wrapper)
            if input_side:  # C = W S W^T: accumulate S = E[x x^T] (and E[x]) instead of C
                acc.update(wrapper.get_last_input())
            elif from_output:
                _accumulate_Ey_and_Eyyt(acc, wrapper)
            else:  # strided / padded 1x1 conv: the reference's y = x W^T over ALL input positions (F:159)
                acc.update(linalg.linear_nt(wrapper.get_last_input(), weight))
    finally:
        wrapper.capture_output = False
        wrapper.output = None
    logger.info("Using mean for covariance" if use_mean else "Not using mean for covariance")
    if use_damping:
        logger.info("Using damping")
    if input_side:
        # centring carries over (E[y] = W E[x]); damping is a multiple of I on C and moves no
        # eigenvector, so it has no input-side counterpart
        s_cov = acc.finalize(use_mean=use_mean, damp_factor=0.0)
        return linalg.eigvecs_from_input_covariance(s_cov, weight, num_vectors)
    damp = EIGEN_DAMPEN_FACTOR if (use_damping and not use_mean) else 0.0
    cov = acc.finalize(use_mean=use_mean, damp_factor=damp)
    _, u = linalg.eigh(cov, k=num_vectors)
    return u


def _accumulate_Ey_and_Eyyt(acc: linalg.CovarianceAccumulator, wrapper: WrappedFALORModule) -> None:
    """F:156-162 for one calibration batch. The layer output captured by the wrapper is
    y + bias; the kernel subtracts the bias while staging, so what is accumulated is exactly the
    reference's y = x W^T: Eyyt += y^T y / N and (when tracked) Ey += mean(y)."""
    acc.update(wrapper.get_last_output_rows(), sub=wrapper.get_bias())


def _get_decomposeable_submodule_names(module: torch.nn.Module) -> list[str]:
    """F:411-414: forward (named_modules) order."""
    return [name for name, mod in module.named_modules() if is_decomposeable_module(mod)]


def _add_meta_to_module_config(module_config: dict[str, Any], module_deco_results: dict[str, Any]) -> None:
    """F:417-421."""
    module_config[utils.MODCONFIG_META_KEY] = {
        k: v for k, v in module_deco_results.items() if k != "decomposed_module"}


def _wrap_in_place(root_module: torch.nn.Module, decomposed_submodule_name: str) -> None:
    """F:236-259."""
    sub = root_module.get_submodule(decomposed_submodule_name)
    if isinstance(sub, torch.nn.Linear):
        wrapped: WrappedFALORModule = WrappedFALORLinear(sub, decomposed_submodule_name)
    elif is_decomposeable_module(sub):
        wrapped = WrappedFALORConv2d1x1(sub, decomposed_submodule_name)
    else:
        raise ValueError(f"Cannot decompose {decomposed_submodule_name}={sub}")
    utils.replace_submodule_in_place(root_module, decomposed_submodule_name, wrapped)


def _unwrap_in_place(root_module: torch.nn.Module, decomposed_submodule_name: str) -> None:
    """F:262-270."""
    sub = root_module.get_submodule(decomposed_submodule_name)
    assert isinstance(sub, WrappedFALORModule)
    utils.replace_submodule_in_place(root_module, decomposed_submodule_name, sub.get_orig_module())
