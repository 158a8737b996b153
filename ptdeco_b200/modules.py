"""Fast-path module for decomposed layers.

`LowRankSequential` IS an `nn.Sequential(first, second)` -- same children, same state-dict keys
(`0.weight`, `1.weight`, `1.bias`), `utils.get_module_config` still reports `"Sequential"` -- whose
forward runs the fused two-GEMM kernel (`ptdeco_lowrank_forward`, rank-k intermediate kept on chip)
when it can: CUDA tensors, bf16 or fp32, no autograd. Anything else takes the reference's path
(`nn.Sequential.forward`, i.e. two F.linear / F.conv2d calls), which is the user's torch code, not a
fallback of a kernel."""
from __future__ import annotations

import torch

from . import linalg


class LowRankSequential(torch.nn.Sequential):
    """Two-factor Linear->Linear or Conv1x1->Conv1x1 pair built by get_decomposed_module
    (F:84-95, F:136-153, D:74-85, D:126-144)."""

    def _fusable(self, x: torch.Tensor) -> bool:
        if len(self) != 2 or not x.is_cuda or (torch.is_grad_enabled() and (
                x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            return False
        a, b = self[0], self[1]
        if isinstance(a, torch.nn.Linear) and isinstance(b, torch.nn.Linear):
            ok = a.bias is None
        elif isinstance(a, torch.nn.Conv2d) and isinstance(b, torch.nn.Conv2d):
            ok = a.bias is None and all(
                c.kernel_size == (1, 1) and c.stride == (1, 1) and c.padding == (0, 0)
                and c.dilation == (1, 1) and c.groups == 1 for c in (a, b))
        else:
            return False
        return ok and x.dtype == a.weight.dtype == b.weight.dtype and x.dtype in (
            torch.float32, torch.bfloat16)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self._fusable(x):
            return super().forward(x)
        a, b = self[0], self[1]
        if isinstance(a, torch.nn.Linear):
            rows = x.reshape(-1, a.in_features)
            y = linalg.lowrank_forward(rows, a.weight, b.weight, b.bias)
            return y.reshape(*x.shape[:-1], b.out_features)
        n, c, h, w = x.shape
        rows = x.permute(0, 2, 3, 1).reshape(-1, c)
        y = linalg.lowrank_forward(rows, a.weight[:, :, 0, 0], b.weight[:, :, 0, 0], b.bias)
        return y.reshape(n, h, w, b.out_channels).permute(0, 3, 1, 2)


def fuse_decomposed_modules_in_place(module: torch.nn.Module) -> int:
    """Swap every two-factor nn.Sequential (as built by the decomposition or by
    utils.apply_decompose_config_in_place) for a LowRankSequential sharing the same children.
    Returns how many were swapped. Config and state dict are unchanged by the swap."""
    swapped = 0
    for name, sub in list(module.named_modules()):
        if type(sub) is not torch.nn.Sequential or len(sub) != 2 or not name:
            continue
        a, b = sub[0], sub[1]
        pair = (isinstance(a, torch.nn.Linear) and isinstance(b, torch.nn.Linear)) or (
            isinstance(a, torch.nn.Conv2d) and isinstance(b, torch.nn.Conv2d)
            and a.kernel_size == (1, 1) and b.kernel_size == (1, 1))
        if not pair or a.bias is not None:
            continue
        if (a.out_features if isinstance(a, torch.nn.Linear) else a.out_channels) != (
                b.in_features if isinstance(b, torch.nn.Linear) else b.in_channels):
            continue
        fused = LowRankSequential(a, b)
        parent_name, _, child = name.rpartition(".")
        setattr(module.get_submodule(parent_name), child, fused)
        swapped += 1
    return swapped
