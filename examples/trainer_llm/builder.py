"""Model construction and artifact loading for the LLM example (reference:
examples/trainer_llm/builder.py). Three ways to name a model:

* `decomposed_model_custom_builder_path`: a Python file exposing
  `make_model_and_tokenizer(*, model_name, model_revision, dtype, model_builder_config)` -- the
  reference's custom-builder contract (:66-91);
* `decomposed_model_name: "random-init:<model_type>"` with the Hugging Face config fields in
  `decomposed_model_custom_builder_config`: random-init weights of that architecture, no network;
* anything else goes to `transformers.AutoModelForCausalLM.from_pretrained` (local cache only here)."""
from __future__ import annotations

import importlib.util
import json
import logging
import sys
from typing import Any, Optional

import torch
import transformers

import ptdeco_b200 as ptdeco
import ptdeco_b200.utils

RANDOM_INIT = "random-init:"

logger = logging.getLogger(__name__)


def _log_linear_submodules(m: torch.nn.Module) -> None:
    lines = ["All linear modules of the model:"]
    for i, (name, mod) in enumerate(
            ((n, x) for n, x in m.named_modules() if isinstance(x, torch.nn.Linear)), start=1):
        bias = "+ bias" if mod.bias is not None else "no bias"
        lines.append(f"  - {name} # ({i}) {bias} {tuple(mod.weight.shape)}")
    logger.info("\n".join(lines))


def _load_custom_builder(path: str):
    spec = importlib.util.spec_from_file_location("builder_custom", path)
    if spec is None or spec.loader is None:
        raise ValueError(f"Error loading custom builder {path}")
    module = importlib.util.module_from_spec(spec)
    sys.modules["builder_custom"] = module
    spec.loader.exec_module(module)
    return module


def make_model_and_tokenizer(*, model_name: str, model_revision: str, model_custom_builder_path: Optional[str],
                             model_custom_builder_config: Optional[dict[str, Any]],
                             enable_gradient_checkpointing: bool, dtype: torch.dtype,
                             log_linears: bool = False):
    tokenizer = None
    if model_custom_builder_path is not None:
        logger.info(f"Custom builder {model_custom_builder_path} - {model_name} revision={model_revision} "
                    f"with {dtype=} grad_checkpointing={enable_gradient_checkpointing}")
        model, tokenizer = _load_custom_builder(model_custom_builder_path).make_model_and_tokenizer(
            model_name=model_name, model_revision=model_revision, dtype=dtype,
            model_builder_config=model_custom_builder_config)
    elif model_name.startswith(RANDOM_INIT):
        model_type = model_name[len(RANDOM_INIT):]
        hf_config = transformers.AutoConfig.for_model(model_type, **(model_custom_builder_config or {}))
        seed = int((model_custom_builder_config or {}).get("seed", 271828))
        logger.info(f"Random-init builder - {model_type} seed={seed} with {dtype=}")
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            model = transformers.AutoModelForCausalLM.from_config(hf_config)
    else:
        logger.info(f"Standard builder - {model_name} revision={model_revision} with {dtype=}")
        tokenizer = transformers.AutoTokenizer.from_pretrained(model_name, revision=model_revision,
                                                               local_files_only=True)
        model = transformers.AutoModelForCausalLM.from_pretrained(model_name, revision=model_revision,
                                                                  torch_dtype=dtype, local_files_only=True)
    if enable_gradient_checkpointing:
        model.gradient_checkpointing_enable()
    if getattr(model.config, "pad_token_id", None) is None:
        model.config.pad_token_id = getattr(model.config, "eos_token_id", None) or 0
    if log_linears:
        _log_linear_submodules(model)
    model.to(dtype)
    model.eval()
    return model, tokenizer


def apply_decompose_config_and_state_dict_in_place(*, model: torch.nn.Module, decompose_config_path: str,
                                                   state_dict_path: str, device: torch.device,
                                                   dtype: torch.dtype, log_linears: bool = False) -> None:
    """Rebuild the decomposed architecture from `decompose_config.json`, then load the weights
    (reference :119-145)."""
    with open(decompose_config_path, "rt") as f:
        decompose_config = json.load(f)
    ptdeco.utils.apply_decompose_config_in_place(model, decompose_config)
    model.to(device)
    model.to(dtype)
    ptdeco.utils.free_gpu_reserved_memory()
    logger.info(f"Applied decompose config {decompose_config_path}")
    model.load_state_dict(torch.load(state_dict_path, map_location=device))
    logger.info(f"Loaded state dict {state_dict_path}")
    model.eval()
    if log_linears:
        _log_linear_submodules(model)


def validate_module_names(model: torch.nn.Module, module_names: Optional[list[str]]) -> None:
    if module_names is None:
        return
    known = {name for name, _ in model.named_modules()}
    unknown = [name for name in module_names if name not in known]
    if unknown:
        raise ValueError(f"Unknown module names specified: {', '.join(unknown)}")
