"""`task: decompose_falor` of the vision example (reference:
examples/trainer_vision/run_decompose_falor.py:28-160, same steps and artifact names):
model -> initial stats -> ptdeco_b200.falor.decompose_in_place(use_mean=False, use_damping=True) ->
`decompose_config.json`, `decompose_state_dict.pt`, `summary.json`."""
from __future__ import annotations

import json
import logging
import pathlib
import time
from typing import Any

import torch

import ptdeco_b200 as ptdeco
import ptdeco_b200.falor

import builder
import configurator

logger = logging.getLogger(__name__)

NORMALIZATIONS = {"imagenet": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), "zero_to_one": ((0.0,) * 3, (1.0,) * 3),
                  "negative_one_to_one": ((0.5,) * 3, (0.5,) * 3)}


def make_image_iterator(batch_size: int, h_w: tuple[int, int], normalization: str, stream: int = 0):
    """Seeded synthetic NCHW batches, normalised like the reference's DALI pipeline output
    (batch i of stream s: torch.Generator().manual_seed(1314159 + 1000003 * s + i))."""
    if normalization not in NORMALIZATIONS:
        raise ValueError(f"Unknown normalization {normalization}")
    mean, std = (torch.tensor(x).view(1, 3, 1, 1) for x in NORMALIZATIONS[normalization])
    i = 0
    while True:
        g = torch.Generator().manual_seed(1314159 + 1000003 * stream + i)
        yield (torch.rand(batch_size, 3, *h_w, generator=g) - mean) / std
        i += 1


def agreement(model: torch.nn.Module, reference_logits: list[torch.Tensor], batches: list[torch.Tensor],
              device: torch.device) -> float:
    """Top-1 agreement (%) with given logits: stands in for ImageNet accuracy on synthetic images."""
    hit = tot = 0
    with torch.no_grad():
        for x, ref in zip(batches, reference_logits):
            pred = model(x.to(device)).argmax(-1)
            hit += int((pred == ref.to(device).argmax(-1)).sum())
            tot += pred.numel()
    return 100.0 * hit / max(tot, 1)


def main(config_raw: dict[str, Any], output_path: pathlib.Path) -> dict[str, Any]:
    config = configurator.DecomposeFALORConfig(**config_raw)
    if config.imagenet_root_dir != "synthetic":
        raise ValueError("imagenet_root_dir: only 'synthetic' is available offline (the reference reads ImageNet "
                         "through NVIDIA DALI)")
    if not torch.cuda.is_available():
        raise RuntimeError("ptdeco_b200 has no CPU path: decompose_falor needs a CUDA device")
    device = torch.device("cuda", torch.cuda.current_device())
    data_iterator = make_image_iterator(config.batch_size, tuple(config.input_h_w), config.normalization, 0)
    val_it = make_image_iterator(config.batch_size, tuple(config.input_h_w), config.normalization, 1)
    val_batches = [next(val_it) for _ in range(4)]

    model = builder.make_model(config.decompose_model_name, log_linears_and_conv1x1=True)
    builder.validate_module_names(model, config.blacklisted_modules)
    model.to(device)
    stats_initial = builder.get_model_stats(model)
    with torch.no_grad():
        logits_initial = [model(x.to(device)).float().cpu() for x in val_batches]

    t0 = time.perf_counter()
    decompose_config = ptdeco.falor.decompose_in_place(
        module=model, device=device, data_iterator=data_iterator,
        proportion_threshold=config.proportion_threshold, kl_final_threshold=config.kl_final_threshold,
        nsr_final_threshold=config.nsr_final_threshold, num_data_steps=config.num_data_steps,
        num_metric_steps=config.num_metric_steps, blacklisted_module_names=config.blacklisted_modules,
        use_float64=config.use_float64, use_mean=False, use_damping=True)
    torch.cuda.synchronize()
    t_decomposition = time.perf_counter() - t0

    stats_final = builder.get_model_stats(model)
    with open(output_path / "decompose_config.json", "wt") as f:
        json.dump(decompose_config, f)
    torch.save(model.state_dict(), output_path / "decompose_state_dict.pt")
    summary = {
        "top1_agreement_with_original": agreement(model, logits_initial, val_batches, device),
        "mparams_initial": stats_initial["mparams"],
        "mparams_initial_decomposeable": stats_initial["mparams_decomposeable"],
        "mparams_final": stats_final["mparams"],
        "mparams_frac": stats_final["mparams"] / stats_initial["mparams"] * 100.0,
        "modules_decomposed": len(decompose_config),
        "time_decomposition": t_decomposition,
        "device": f"{device} @ {torch.cuda.get_device_name(device)}",
    }
    with open(output_path / "summary.json", "wt") as f:
        json.dump(summary, f)
    logger.info(f"summary: {summary}")
    return summary
