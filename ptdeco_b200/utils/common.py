"""Glue helpers with the semantics of the reference's src/ptdeco/utils/common.py (U/c)."""
from __future__ import annotations

import gc
import logging
from typing import Any, TypeVar

import torch

__all__ = [
    "replace_submodule_in_place",
    "split_module_parent_child_name",
    "get_type_name",
    "is_compound_module",
    "get_num_params",
    "get_default_device",
    "to_device",
    "relieve_gpu_memory_pressure",
    "free_gpu_reserved_memory",
    "get_gpu_reserved_memory_gb",
]

logger = logging.getLogger("ptdeco.utils.common")

T = TypeVar("T", torch.Tensor, dict)


def replace_submodule_in_place(root_module: torch.nn.Module, submodule_name: str,
                               new_submodule: torch.nn.Module) -> None:
    """U/c:88-93: setattr on the parent (AttributeError from get_submodule on bad names)."""
    parent_name, child_name = split_module_parent_child_name(submodule_name)
    setattr(root_module.get_submodule(parent_name), child_name, new_submodule)


def split_module_parent_child_name(target: str) -> tuple[str, str]:
    parent, _, child = target.rpartition(".")
    return parent, child


def get_type_name(o: Any) -> str:
    t = type(o)
    return f"{t.__module__}.{t.__name__}"


def is_compound_module(m: torch.nn.Module) -> bool:
    return next(m.children(), None) is not None


def get_num_params(m: torch.nn.Module, only_trainable: bool = False) -> int:
    """U/c:58-63: parameters de-duplicated by storage pointer (tied weights count once)."""
    seen: dict[int, torch.nn.Parameter] = {}
    for p in m.parameters():
        if only_trainable and not p.requires_grad:
            continue
        seen[p.data_ptr()] = p
    return sum(p.numel() for p in seen.values())


def get_default_device(module: torch.nn.Module) -> torch.device:
    """U/c:75-80: device of the first parameter, cpu for parameter-less modules."""
    for p in module.parameters():
        return p.device
    return torch.device("cpu")


def to_device(o: T, device: torch.device) -> T:
    """Moves a tensor, or the tensor values of a dict (other values pass through), to `device`.
    Same contract as the reference helper (U/c:25-36): any other type raises ValueError."""
    if isinstance(o, dict):
        moved = dict(o)
        for key, value in o.items():
            if isinstance(value, torch.Tensor):
                moved[key] = value.to(device)
        return moved
    if not isinstance(o, torch.Tensor):
        raise ValueError(f"Unsupported type {type(o)}")
    return o.to(device)


def relieve_gpu_memory_pressure(threshold: float = 0.6) -> bool:
    """The reference calls free_gpu_reserved_memory() (gc.collect + empty_cache) between layers
    (D:737,787,795) so that a fragmented caching allocator does not run a small GPU out of memory.
    Each call costs a Python GC pass plus re-cudaMalloc of every activation on the next forward
    (measured ~0.1 s per call on the 8B-shape model: tens of seconds over 224 layers). Here the
    cache is dropped only when the current device's reserved memory exceeds `threshold` of its
    capacity. Returns whether it was dropped."""
    if not torch.cuda.is_available():
        return False
    dev = torch.cuda.current_device()
    total = torch.cuda.get_device_properties(dev).total_memory
    if torch.cuda.memory_reserved(dev) <= threshold * total:
        return False
    free_gpu_reserved_memory()
    return True


def free_gpu_reserved_memory() -> None:
    """Python GC pass + torch.cuda.empty_cache(), logging the reserved memory around it (U/c:46-55)."""
    if torch.cuda.is_available():
        gib_before = get_gpu_reserved_memory_gb()
        gc.collect()
        torch.cuda.empty_cache()
        gib_after = get_gpu_reserved_memory_gb()
        logger.info(f"GPU memory: {gib_before:.2f} -> {gib_after:.2f} GB ({(gib_after - gib_before):.2f} GB)")


def get_gpu_reserved_memory_gb() -> float:
    """Reserved bytes of the caching allocator over every visible device, in GiB (U/c:39-43)."""
    reserved = sum(torch.cuda.memory_reserved(device=i) for i in range(torch.cuda.device_count()))
    return reserved / float(1 << 30)
