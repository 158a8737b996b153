"""ORACLE (test infrastructure, NOT the product): the reference's per-layer arithmetic restated with
the reference's own library, torch on CPU (MKL GEMM / LAPACK syevd) -- the closest thing to running
the reference's call sites on the GPU box's host cores, where /root/reference does not exist.
Used by bench.py's cpu_baseline / --impl reference legs and by tests only.
"""
from __future__ import annotations

import torch

EIGEN_DAMPEN_FACTOR = 0.01  # F:22, D:14


def update_Eyyt_in_place(Eyyt: torch.Tensor, y_reshaped: torch.Tensor) -> None:
    """D:147-152 (and F:160): Eyyt += einsum("bp,bq->pq", y, y) / N."""
    Eyyt += torch.einsum("bp,bq->pq", y_reshaped, y_reshaped) / y_reshaped.shape[0]


def get_eigenvectors(Eyyt: torch.Tensor) -> torch.Tensor:
    """D:155-163: damping 0.01 * mean(diag) then `_, u = torch.linalg.eigh(Eyyt)`."""
    damp = EIGEN_DAMPEN_FACTOR * torch.mean(torch.diag(Eyyt))
    idx = torch.arange(Eyyt.shape[-1])
    Eyyt[idx, idx] = Eyyt[idx, idx] + damp
    _, u = torch.linalg.eigh(Eyyt)
    return u


def two_factor_forward(x: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor, bias) -> torch.Tensor:
    """Forward of the nn.Sequential(Linear, Linear) built at F:84-95 / D:74-85."""
    return torch.nn.functional.linear(torch.nn.functional.linear(x, w1), w2, bias)
