"""Rank-search metrics with the semantics of src/ptdeco/utils/losses_primitives.py (U/l), computed
by the sm_100a reduction kernels (ptdeco_nsr_metric / ptdeco_kl_metric). CUDA tensors only."""
from __future__ import annotations

import torch

from .. import _native as nat

__all__ = [
    "calc_per_channel_noise_to_signal_ratio",
    "calc_kl_divergence",
    "calc_kl_loss",
]


def _rows_by_channels(t: torch.Tensor, non_channel_dim: tuple[int, ...]) -> torch.Tensor:
    nd = t.dim()
    red = sorted(d % nd for d in non_channel_dim)
    keep = [d for d in range(nd) if d not in red]
    rows = 1
    for d in red:
        rows *= t.shape[d]
    return t.permute(*red, *keep).reshape(rows, -1).contiguous()


def calc_per_channel_noise_to_signal_ratio(
    x: torch.Tensor,
    y: torch.Tensor,
    non_channel_dim: tuple[int, ...] = (0, 2, 3),
    epsilon: float = 1e-3,
    mode: str = "mean",
) -> torch.Tensor:
    """U/l:10-22: mean over channels of mean((x-y)^2) / (unbiased var(y) + eps); `mode` is unused
    in the reference as well. Returns a 0-dim fp32 tensor on x's device (no host sync)."""
    nat.require_cuda(x, "x")
    nat.require_cuda(y, "y")
    if x.shape != y.shape:
        raise ValueError(f"shape mismatch {tuple(x.shape)} vs {tuple(y.shape)}")
    if x.dtype != y.dtype:
        y = y.to(x.dtype)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x, y = x.float(), y.float()
    xr = _rows_by_channels(x, tuple(non_channel_dim))
    yr = _rows_by_channels(y, tuple(non_channel_dim))
    rows, ch = xr.shape
    L = nat.lib()
    out = torch.empty((), dtype=torch.float32, device=x.device)
    with nat.device_of(x):
        ws = nat.WORKSPACE.get(x.device, L.ptdeco_nsr_workspace_bytes(ch))
        nat.check(L.ptdeco_nsr_metric(xr.data_ptr(), yr.data_ptr(), nat.dtype_code(xr), rows, ch,
                                      float(epsilon), ws.data_ptr(), ws.numel(), out.data_ptr(),
                                      nat.stream_ptr(x.device)), "ptdeco_nsr_metric")
    return out


def calc_kl_divergence(q_logits: torch.Tensor, p_logits: torch.Tensor) -> torch.Tensor:
    """U/l:48-54 per-row KL(p || q) of softmax(logits). Only the [rows, classes] layout used by
    calc_kl_loss is kernel-backed; the per-row vector is not on the hot path, so it is derived
    from log-softmax with torch ops on the same device."""
    nat.require_cuda(q_logits, "q_logits")
    lq = torch.log_softmax(q_logits.float(), dim=-1)
    lp = torch.log_softmax(p_logits.float(), dim=-1)
    return (lp.exp() * (lp - lq)).sum(dim=1)


def calc_kl_loss(student_logits: torch.Tensor, teacher_logits: torch.Tensor) -> torch.Tensor:
    """U/l:57-63: mean over rows of max(KL(t||s), KL(s||t)). 0-dim fp32 tensor, no host sync."""
    nat.require_cuda(student_logits, "student_logits")
    nat.require_cuda(teacher_logits, "teacher_logits")
    if student_logits.dim() != 2 or student_logits.shape != teacher_logits.shape:
        raise ValueError("calc_kl_loss expects two [rows, classes] logit tensors of equal shape")
    s, t = student_logits.contiguous(), teacher_logits.contiguous()
    if s.dtype != t.dtype or s.dtype not in (torch.float32, torch.bfloat16):
        s, t = s.float(), t.float()
    out = torch.empty((), dtype=torch.float32, device=s.device)
    with nat.device_of(s):
        nat.check(nat.lib().ptdeco_kl_metric(s.data_ptr(), t.data_ptr(), nat.dtype_code(s), s.shape[0],
                                             s.shape[1], out.data_ptr(), nat.stream_ptr(s.device)),
                  "ptdeco_kl_metric")
    return out
