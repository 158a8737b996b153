"""decompose_config schema <-> modules, semantics of src/ptdeco/utils/modconfig.py (U/m).

The JSON layout is the artifact contract (README.md:56-105 of the reference): configs written by
the reference load here and vice versa.
"""
from __future__ import annotations

import collections
import logging
from typing import Any

import torch

from . import common

__all__ = [
    "get_module_config",
    "build_module_from_config",
    "apply_decompose_config_in_place",
    "MODCONFIG_META_KEY",
]

logger = logging.getLogger("ptdeco.utils.modconfig")

MODCONFIG_META_KEY = "__meta__"  # U/m:18

_CONV_KEYS = ("in_channels", "out_channels", "kernel_size", "groups", "bias", "stride", "padding",
              "padding_mode", "dilation")


def get_module_config(m: torch.nn.Module) -> dict[str, Any]:
    """U/m:21-61. Key order matches the reference so dumped JSON is byte-identical."""
    if isinstance(m, torch.nn.Sequential):
        return {"type": "Sequential",
                "modules": {name: get_module_config(child) for name, child in m.named_children()}}
    if isinstance(m, torch.nn.Conv2d):
        return {
            "type": "Conv2d",
            "in_channels": m.in_channels,
            "out_channels": m.out_channels,
            "kernel_size": m.kernel_size,
            "bias": m.bias is not None,
            "groups": m.groups,
            "padding": m.padding,
            "padding_mode": m.padding_mode,
            "stride": m.stride,
            "dilation": m.dilation,
        }
    if isinstance(m, torch.nn.Linear):
        return {"type": "Linear", "in_features": m.in_features, "out_features": m.out_features,
                "bias": m.bias is not None}
    raise ValueError(f"get_module_config not implemented for {type(m)}")


def build_module_from_config(config: dict[str, Any]) -> torch.nn.Module:
    """U/m:64-111. Weights are freshly initialised; callers load a state dict afterwards."""
    kind = config.get("type")
    if kind == "Sequential":
        children = config["modules"]
        built = collections.OrderedDict((k, build_module_from_config(v)) for k, v in children.items())
        if next(iter(children.keys())) == "0":
            return torch.nn.Sequential(*built.values())  # positional, like the reference (U/m:92-94)
        return torch.nn.Sequential(built)
    if kind == "Conv2d":
        return torch.nn.Conv2d(**{k: config[k] for k in _CONV_KEYS})
    if kind == "Linear":
        return torch.nn.Linear(in_features=config["in_features"],
                               out_features=config["out_features"], bias=config["bias"])
    raise ValueError(f"type={kind!r} not supported")


def apply_decompose_config_in_place(module: torch.nn.Module, decompose_config: dict[str, Any]) -> None:
    """U/m:114-130: swap every configured submodule for a freshly built one on the old one's device."""
    counter: collections.Counter[str] = collections.Counter()
    for name, cfg in decompose_config.items():
        old = module.get_submodule(name)
        new = build_module_from_config(cfg)
        new.to(common.get_default_device(old))
        common.replace_submodule_in_place(module, name, new)
        counter[common.get_type_name(old)] += 1
    for type_name, count in counter.items():
        logger.info(f"Decomposed {count} instances of {type_name}")
