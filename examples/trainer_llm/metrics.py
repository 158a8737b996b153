"""Model statistics of the LLM example: perplexity over a loader and parameter counts (reference:
examples/trainer_llm/metrics.py). FLOP counting (fvcore) and lm_eval are not available offline."""
from __future__ import annotations

import logging
import time

import torch

import ptdeco_b200.utils

logger = logging.getLogger(__name__)


def calc_perplexity(model: torch.nn.Module, testloader, device: torch.device, pad_token_id: int) -> float:
    """exp(mean next-token NLL over all non-pad positions of the loader) (reference :41-84)."""
    start = time.perf_counter()
    model.eval()
    nll = torch.zeros((), dtype=torch.float64, device=device)
    count = torch.zeros((), dtype=torch.float64, device=device)
    for batch in testloader:
        batch = ptdeco_b200.utils.to_device(batch, device)
        logits = model(input_ids=batch["input_ids"]).logits[:, :-1, :].float()
        labels = batch["input_ids"][:, 1:]
        keep = batch["attention_mask"][:, 1:].bool()
        losses = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1),
                                                   reduction="none").reshape(labels.shape)
        nll += (losses * keep).sum().double()
        count += keep.sum().double()
    ppl = float(torch.exp(nll / count.clamp_min(1.0)).item())
    logger.info(f"Perplexity evaluation took {time.perf_counter() - start:.2f} s, {ppl=:.4f}")
    return ppl


def get_params(model: torch.nn.Module) -> int:
    return ptdeco_b200.utils.get_num_params(model)
