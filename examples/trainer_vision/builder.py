"""Model construction / statistics / artifact loading of the vision example (reference:
examples/trainer_vision/builder.py). `<builder>.<model>` names: `torchvision.<name>` builds a
seeded random-init torchvision classifier; `timm.<name>` (the reference's only builder) needs the
timm package and pretrained weights, neither of which exists offline."""
from __future__ import annotations

import json
import logging
from typing import Optional

import torch

import ptdeco_b200 as ptdeco
import ptdeco_b200.falor
import ptdeco_b200.utils

logger = logging.getLogger(__name__)


def make_model(model_name: str, log_linears_and_conv1x1: bool = False, seed: int = 271828) -> torch.nn.Module:
    builder, _, name = model_name.partition(".")
    logger.info(f"Creating model: {builder} {name}")
    if builder == "torchvision":
        import torchvision
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            model = torchvision.models.get_model(name, weights=None)
    elif builder == "timm":
        raise ValueError("the timm builder needs the `timm` package and network access; use torchvision.<name>")
    else:
        raise ValueError(f"Unknown model builder {builder}")
    model.eval()
    if log_linears_and_conv1x1:
        lines = ["All decomposeable modules of the model:"]
        for i, (n, m) in enumerate(((n, m) for n, m in model.named_modules()
                                    if ptdeco.falor.is_decomposeable_module(m)), start=1):
            kind = "linear" if isinstance(m, torch.nn.Linear) else "conv1x1"
            bias = "+ bias" if m.bias is not None else "no bias"
            lines.append(f"  - {n} # ({i}) {kind} {bias} {tuple(m.weight.shape)}")
        logger.info("\n".join(lines))
    return model


def validate_module_names(model: torch.nn.Module, module_names: Optional[list[str]]) -> None:
    if module_names is None:
        return
    known = {name for name, _ in model.named_modules()}
    unknown = [name for name in module_names if name not in known]
    if unknown:
        raise ValueError(f"Unknown module names specified: {', '.join(unknown)}")


def get_model_stats(model: torch.nn.Module) -> dict[str, float]:
    """Parameter counts (the reference also reports fvcore FLOPs, not installed here)."""
    dec = sum(p.numel() for m in model.modules() if ptdeco.falor.is_decomposeable_module(m)
              for p in m.parameters(recurse=False))
    return {"mparams": ptdeco.utils.get_num_params(model) / 1e6, "mparams_decomposeable": dec / 1e6}


def apply_decompose_config_and_state_dict_in_place(*, model: torch.nn.Module, decompose_config_path: str,
                                                   state_dict_path: str, device: torch.device) -> None:
    with open(decompose_config_path, "rt") as f:
        decompose_config = json.load(f)
    ptdeco.utils.apply_decompose_config_in_place(model, decompose_config)
    model.to(device)
    model.load_state_dict(torch.load(state_dict_path, map_location=device))
    model.eval()
