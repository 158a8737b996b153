"""Deterministic, index-addressable synthetic streams.

batch i of stream s is generated from torch.Generator().manual_seed(1314159 + 1000003*s + i)
(1314159 is the reference tests' data seed, tests/test_deco_primitives_falor.py:19-20).
"""
from __future__ import annotations

from typing import Callable, Iterator

import torch

DATA_SEED = 1314159
MODEL_SEED = 271828


def _gen(stream: int, index: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(DATA_SEED + 1000003 * stream + index)
    return g


def image_batch(stream: int, index: int, bs: int, ch: int = 3, size: int = 224) -> torch.Tensor:
    return torch.rand(bs, ch, size, size, generator=_gen(stream, index))


def lowrank_image_batch(stream: int, index: int, bs: int, ch: int, size: int, latent: int,
                        noise: float = 0.02) -> torch.Tensor:
    """Images that live near a `latent`-dimensional subspace (fixed seeded basis + small noise) so
    layer activations have a decaying spectrum and rank decisions have margin."""
    gb = torch.Generator()
    gb.manual_seed(DATA_SEED - 7)
    basis = torch.randn(latent, ch * size * size, generator=gb) / latent ** 0.5
    g = _gen(stream, index)
    z = torch.randn(bs, latent, generator=g)
    scale = torch.logspace(0, -2, latent)
    x = (z * scale) @ basis + noise * torch.randn(bs, ch * size * size, generator=g)
    return x.reshape(bs, ch, size, size)


def token_batch(stream: int, index: int, bs: int, seq: int, vocab: int) -> dict[str, torch.Tensor]:
    ids = torch.randint(0, vocab, (bs, seq), generator=_gen(stream, index))
    return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids.clone()}


def step_spectrum_activations(n: int, d: int, seed: int = 0, dtype=torch.float32,
                              device="cpu") -> torch.Tensor:
    """Y = Z diag(s) Q with singular scales 1 / 0.3 / 0.1 / 0.01 on index blocks
    [0,d/8) [d/8,d/4) [d/4,d/2) [d/2,d): every tested k in {d/8, d/4, d/2} sits on a >= x9
    eigenvalue step. Q is a product of seeded Householder reflectors (cheap for
    large d, exactly orthogonal)."""
    g = torch.Generator(device=device)
    g.manual_seed(DATA_SEED + 17 * seed + d)
    z = torch.randn(n, d, generator=g, device=device, dtype=torch.float32)
    s = torch.empty(d, device=device)
    s[: d // 8] = 1.0
    s[d // 8: d // 4] = 0.3
    s[d // 4: d // 2] = 0.1
    s[d // 2:] = 0.01
    y = z * s
    gq = torch.Generator(device=device)
    gq.manual_seed(DATA_SEED + 99 + d)  # Q depends on d only, not on the batch seed
    for _ in range(4):
        v = torch.randn(d, generator=gq, device=device)
        v = v / v.norm()
        y = y - 2.0 * (y @ v)[:, None] * v[None, :]
    return y.to(dtype)


class IndexedStream:
    """Iterator over make(i) for i = start, start+1, ...; `position` is observable so tests can
    check that the drop-in consumes batches in the reference's order."""

    def __init__(self, make: Callable[[int], object], start: int = 0):
        self.make = make
        self.position = start

    def __iter__(self) -> Iterator:
        return self

    def __next__(self):
        item = self.make(self.position)
        self.position += 1
        return item
