"""`task: decompose_dwain` of the LLM example: build the model, decompose it with
ptdeco_b200.dwain.decompose_in_place, write `decompose_config.json`, `decompose_state_dict.pt` and
`summary.json` (reference: examples/trainer_llm/run_decompose_dwain.py:136-305, same steps and
file names). Offline: synthetic token data (datasets_synth), no lm_eval, no FLOP counter."""
from __future__ import annotations

import collections.abc
import json
import logging
import pathlib
import random
import time
from typing import Any

import torch

import ptdeco_b200 as ptdeco
import ptdeco_b200.dwain

import builder
import configurator
import datasets_synth
import dwain_wrapper_module
import metrics
import utils

PPL_N_BATCHES = 8
LOADER_MERGER_SEED = 314159

logger = logging.getLogger(__name__)


def make_infinite_iterator_single(dl: collections.abc.Iterable[Any]):
    while True:
        yield from dl


def make_infinite_iterator_multi(dls: collections.abc.Sequence[collections.abc.Iterable[Any]], seed: int):
    """One whole pass over a randomly chosen loader at a time (reference :36-48)."""
    rng = random.Random(seed)
    its = [make_infinite_iterator_single(dl) for dl in dls]
    while True:
        yield from its[rng.randrange(len(its))]


def make_dataloaders(config: configurator.DecomposeDWAINConfig, vocab_size: int):
    names = config.decomposition_data_name
    names = [names] if isinstance(names, str) else list(names)
    need = config.num_data_steps + 64  # more than one calibration pass draws before wrapping around
    decomposition_dls = [
        datasets_synth.SyntheticTokenBatches(n, vocab_size, config.decomposition_data_max_length,
                                             config.decomposition_data_batch_size, need) for n in names]
    perplexity_dl = datasets_synth.SyntheticTokenBatches(
        config.perplexity_data_name, vocab_size, config.perplexity_data_max_length,
        config.perplexity_data_batch_size, PPL_N_BATCHES)
    return decomposition_dls, perplexity_dl


def make_finetune_fn(config: configurator.DecomposeDWAINConfig, ft_iterator):
    if config.finetuning_run and config.finetuning_use_lora:
        logger.info("Creating lora finetuning function")
        return lambda m, device, decomposed: dwain_wrapper_module.finetune_lora(
            model=m, device=device, decomposed_modules=decomposed, ft_iterator=ft_iterator,
            num_steps=config.finetuning_num_steps, lr=config.finetuning_lr,
            num_last_modules_to_finetune=config.finetuning_num_last_finetuned_modules,
            use_rank_pattern=config.finetuning_use_rank_pattern,
            min_rank_to_finetune=config.finetuning_lora_min_rank)
    if config.finetuning_run:
        logger.info("Creating full finetuning function")
        return lambda m, device, decomposed: dwain_wrapper_module.finetune_full(
            model=m, device=device, decomposed_modules=decomposed, ft_iterator=ft_iterator,
            num_steps=config.finetuning_num_steps, lr=config.finetuning_lr,
            num_last_modules_to_finetune=config.finetuning_num_last_finetuned_modules)
    logger.info("Creating empty finetuning function")
    return lambda m, device, decomposed: m


def main(config_raw: dict[str, Any], output_path: pathlib.Path, process_group=None) -> dict[str, Any]:
    start = time.perf_counter()
    config = configurator.DecomposeDWAINConfig(**config_raw)
    if config.lm_eval_initial or config.lm_eval_tasks:
        raise ValueError("lm_eval_tasks / lm_eval_initial need the `lm_eval` package, which is not installed here")
    dtype = utils.conv_str_to_dtype(config.decomposed_model_dtype)
    if not torch.cuda.is_available():
        raise RuntimeError("ptdeco_b200 has no CPU path: decompose_dwain needs a CUDA device")
    device = torch.device("cuda", torch.cuda.current_device())

    model, _ = builder.make_model_and_tokenizer(
        model_name=config.decomposed_model_name, model_revision=config.decomposed_model_revision,
        model_custom_builder_path=config.decomposed_model_custom_builder_path,
        model_custom_builder_config=config.decomposed_model_custom_builder_config,
        enable_gradient_checkpointing=config.decomposed_model_enable_gradient_checkpointing,
        dtype=dtype, log_linears=True)
    model.to(device)
    builder.validate_module_names(model, config.blacklisted_modules)

    decomposition_dls, perplexity_dl = make_dataloaders(config, model.config.vocab_size)

    with torch.no_grad():
        perplexity_initial = metrics.calc_perplexity(model, perplexity_dl, device, model.config.pad_token_id)
    params_initial = metrics.get_params(model) / 1.0e6
    logger.info(f"{perplexity_initial=} {params_initial=}")

    model_wrapped = dwain_wrapper_module.WrapperModule(model)
    model_wrapped.eval()
    if len(decomposition_dls) > 1:
        logger.info("Using multi-loader data iterator")
        decomposition_it = make_infinite_iterator_multi(decomposition_dls, LOADER_MERGER_SEED)
    else:
        logger.info("Using single-loader data iterator")
        decomposition_it = make_infinite_iterator_single(decomposition_dls[0])
    finetune_fn = make_finetune_fn(config, decomposition_it)

    t_deco = time.perf_counter()
    decompose_config = ptdeco.dwain.decompose_in_place(
        module=model_wrapped, device=device,
        blacklisted_module_names=dwain_wrapper_module.add_prefix(config.blacklisted_modules),
        data_iterator=decomposition_it, loss_fn=dwain_wrapper_module.ce_loss, finetune_fn=finetune_fn,
        metric_iterator=decomposition_it, nsr_final_threshold=config.nsr_final_threshold,
        num_data_steps=config.num_data_steps, num_metric_steps=config.num_metric_steps,
        min_rank=config.min_rank, trade_off_factor=config.trade_off_factor,
        reduction_factor=config.reduction_factor, max_accepted_ppl_diff=config.max_accepted_ppl_diff,
        decompose_in_float64=config.decompose_in_float64,
        precomputing_covariance_num_splits=config.precomputing_covariance_num_splits,
        **({"process_group": process_group} if process_group is not None else {}))
    torch.cuda.synchronize()
    time_decomposition = time.perf_counter() - t_deco

    dwain_wrapper_module.save_raw_model_decompose_config_and_state_dict(
        output_path, decompose_config, model_wrapped.raw_model.state_dict())

    with torch.no_grad():
        perplexity_final = metrics.calc_perplexity(model_wrapped.raw_model, perplexity_dl, device,
                                                   model.config.pad_token_id)
    params_final = metrics.get_params(model_wrapped.raw_model) / 1.0e6
    params_frac = params_final / params_initial * 100.0
    logger.info(f"{perplexity_initial=} -> {perplexity_final=}")
    logger.info(f"{params_initial=} -> {params_final=} {params_frac:.2f}")

    summary = {
        "perplexity_initial": perplexity_initial, "perplexity_final": perplexity_final,
        "mparams_initial": params_initial, "mparams_final": params_final, "mparams_frac": params_frac,
        "gflops_initial": None, "gflops_final": None, "gflops_frac": None,  # fvcore is not installed
        "modules_decomposed": len(decompose_config),
        "time_decomposition": time_decomposition,
        "time_decomposition_and_perplex_eval": time.perf_counter() - start,
        "time_lm_eval_initial": -1.0, "time_lm_eval_final": -1.0,
        "device": f"{device} @ {torch.cuda.get_device_name(device)}",
    }
    with open(output_path / "summary.json", "wt") as f:
        json.dump(summary, f)
    return summary
