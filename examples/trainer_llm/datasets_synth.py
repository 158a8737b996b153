"""Offline stand-in for the reference's datasets_hf.py (which tokenises Hugging Face datasets):
index-addressable synthetic token batches in the same dict layout the reference's loaders yield
(`input_ids`, `attention_mask`, `labels`, all [batch, max_length] int64).

Names: "synthetic.tokens" or "synthetic.tokens:<stream id>". Batch i of stream s is drawn from
torch.Generator().manual_seed(1314159 + 1000003 * s + i) (the repo's synthetic-input convention)."""
from __future__ import annotations

import torch

PREFIX = "synthetic.tokens"


def is_synthetic(name: str) -> bool:
    return name == PREFIX or name.startswith(PREFIX + ":")


class SyntheticTokenBatches:
    """A finite, re-iterable "dataloader" of `num_batches` token batches."""

    def __init__(self, name: str, vocab_size: int, max_length: int, batch_size: int, num_batches: int):
        if not is_synthetic(name):
            raise ValueError(
                f"dataset {name!r}: only '{PREFIX}[:<stream>]' is available offline (the reference's "
                "datasets_hf loaders need the `datasets` package and network access)")
        self.stream = int(name.split(":", 1)[1]) if ":" in name else 0
        self.vocab_size, self.max_length = int(vocab_size), int(max_length)
        self.batch_size, self.num_batches = int(batch_size), int(num_batches)

    def __len__(self) -> int:
        return self.num_batches

    def batch(self, i: int) -> dict[str, torch.Tensor]:
        g = torch.Generator().manual_seed(1314159 + 1000003 * self.stream + i)
        ids = torch.randint(0, self.vocab_size, (self.batch_size, self.max_length), generator=g)
        return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids.clone()}

    def __iter__(self):
        for i in range(self.num_batches):
            yield self.batch(i)
