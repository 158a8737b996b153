"""Model wrapper, loss, artifact writer and fine-tuning hooks of the dwain LLM example
(reference: examples/trainer_llm/dwain_wrapper_module.py). dwain.decompose_in_place drives a
module that maps a batch dict to logits; Hugging Face causal LMs are wrapped so that they do, and
the `raw_model.` prefix the wrapper adds to module names is stripped again on save, so the
artifacts load into the bare model (builder.apply_decompose_config_and_state_dict_in_place)."""
from __future__ import annotations

import collections
import json
import logging
import pathlib
import time
from typing import Any

import torch

import ptdeco_b200 as ptdeco
import ptdeco_b200.utils

PREFIX = "raw_model."

logger = logging.getLogger(__name__)


class WrapperModule(torch.nn.Module):
    """dict -> logits. Only `input_ids` reaches the model (reference :27-30)."""

    def __init__(self, model: torch.nn.Module):
        super().__init__()
        self.raw_model = model
        self.config = getattr(model, "config", None)

    def forward(self, x: dict[str, torch.Tensor], **kwargs: Any) -> torch.Tensor:
        return self.raw_model(input_ids=x["input_ids"], **kwargs).logits


def ce_loss(input_dict: dict[str, torch.Tensor], output: torch.Tensor) -> torch.Tensor:
    """Next-token cross entropy with padded positions' logits zeroed (reference :33-46)."""
    labels = input_dict["labels"][..., 1:].contiguous()
    mask = input_dict["attention_mask"][..., :-1]
    logits = output[..., :-1, :].contiguous() * mask.unsqueeze(-1)
    return torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1))


def add_prefix(module_names: list[str]) -> list[str]:
    return [PREFIX + name for name in module_names]


def _strip(name: str) -> str:
    return name[len(PREFIX):] if name.startswith(PREFIX) else name


def strip_prefix_list(module_names: list[str]) -> list[str]:
    return [_strip(name) for name in module_names]


def strip_prefix_dict(d: dict[str, Any]) -> dict[str, Any]:
    out: dict[str, Any] = collections.OrderedDict() if isinstance(d, collections.OrderedDict) else {}
    for key, value in d.items():
        out[_strip(key)] = value
    return out


def save_raw_model_decompose_config_and_state_dict(
        output_path: pathlib.Path, decompose_config: dict[str, Any],
        state_dict: dict[str, torch.Tensor]) -> None:
    """`decompose_config.json` + `decompose_state_dict.pt` with bare-model names (reference :78-89)."""
    with open(output_path / "decompose_config.json", "wt") as f:
        json.dump(strip_prefix_dict(decompose_config), f)
    torch.save(strip_prefix_dict(state_dict), output_path / "decompose_state_dict.pt")


def _select_trainable(model: torch.nn.Module, names: list[str], what: str) -> None:
    for pname, param in model.named_parameters():
        if any(n in pname for n in names):
            logger.info(f"{what} - enabling grad for {pname}, {param.requires_grad=}")
        else:
            param.requires_grad = False


def _linear_warmup(optimizer: torch.optim.Optimizer, warmup: int, total: int):
    def factor(step: int) -> float:
        if step < warmup:
            return step / max(1, warmup)
        return max(0.0, (total - step) / max(1, total - warmup))
    return torch.optim.lr_scheduler.LambdaLR(optimizer, factor)


def finetune_full(*, model: torch.nn.Module, device: torch.device, ft_iterator, decomposed_modules: list[str],
                  num_last_modules_to_finetune: int = 8, num_steps: int = 100,
                  lr: float = 0.0001) -> torch.nn.Module:
    """finetune_fn of dwain.decompose_in_place: AdamW on the parameters of the last few decomposed
    modules, everything else frozen, linear warm-up (10 steps) then linear decay (reference :92-147).
    The decomposition runs under no_grad, so gradients are switched on here; decomposed layers are
    ptdeco_b200.modules.LowRankSequential, which takes the autograd path while grads are on."""
    if len(decomposed_modules) == 0:
        logger.info("Skipping full fine-tuning - empty list of decomposed modules")
        return model
    start = time.perf_counter()
    _select_trainable(model, decomposed_modules[-num_last_modules_to_finetune:], "full fine-tuning")
    params = [p for p in model.parameters() if p.requires_grad]
    if not params:
        return model
    optimizer = torch.optim.AdamW(params, lr=lr)
    scheduler = _linear_warmup(optimizer, 10, num_steps)
    model.train()
    total = 0.0
    with torch.enable_grad():
        for step in range(num_steps):
            batch = ptdeco.utils.to_device(next(ft_iterator), device)
            optimizer.zero_grad()
            loss = ce_loss(batch, model(batch))
            loss.backward()
            optimizer.step()
            scheduler.step()
            total += loss.item()
            if step % 10 == 0:
                logger.info(f"Step: {step}/{num_steps}, loss: {total / (step + 1)}")
    model.eval()
    logger.info(f"Full fine-tuning took {time.perf_counter() - start:.2f} seconds")
    return model


def finetune_lora(**kwargs: Any) -> torch.nn.Module:
    """The reference's LoRA hook (:150-265) is a thin driver around `peft`, which this image does not
    have; fail loudly rather than silently skipping the fine-tuning a config asked for."""
    try:
        import peft  # noqa: F401
    except ImportError as exc:
        raise RuntimeError("finetuning_use_lora needs the `peft` package; set finetuning_use_lora: false "
                           "(full fine-tuning of the last decomposed modules) or finetuning_run: false") from exc
    raise NotImplementedError("LoRA fine-tuning is not part of this example")
