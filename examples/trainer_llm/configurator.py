"""Config schemas of the LLM trainer example. Field names are the reference's
(examples/trainer_llm/configurator.py:74-125), so its YAML files load unchanged; what this offline
image cannot provide (hub downloads, `datasets`, `lm_eval`, `peft`) is rejected when the run starts,
with a message saying which field asked for it."""
from __future__ import annotations

from typing import Any, Literal, Optional, Union

import pydantic

DTYPES = ("torch.float32", "torch.bfloat16", "torch.float16")


class _Base(pydantic.BaseModel):
    model_config = pydantic.ConfigDict(extra="forbid")

    ptdeco_trainer_llm_version: Optional[str] = None
    ptdeco_version: Optional[str] = None


class DecomposeDWAINConfig(_Base):
    task: Literal["decompose_dwain"]

    # model
    decomposed_model_name: str
    decomposed_model_revision: str = "main"
    decomposed_model_custom_builder_path: Optional[str] = None
    decomposed_model_custom_builder_config: Optional[dict[str, Any]] = None
    decomposed_model_dtype: str
    decomposed_model_enable_gradient_checkpointing: bool = False

    # data
    decomposition_data_name: Union[str, list[str]]
    decomposition_data_separator: str = "\n\n"
    decomposition_data_max_length: int
    decomposition_data_batch_size: int
    perplexity_data_name: str
    perplexity_data_separator: str = "\n\n"
    perplexity_data_max_length: int
    perplexity_data_batch_size: int

    # decomposition (ptdeco.dwain.decompose_in_place keyword arguments)
    num_data_steps: int
    num_metric_steps: int
    trade_off_factor: float
    reduction_factor: float
    max_accepted_ppl_diff: float
    nsr_final_threshold: float
    min_rank: int
    decompose_in_float64: bool
    precomputing_covariance_num_splits: Optional[int] = None
    blacklisted_modules: list[str]

    # fine-tuning between layers (finetune_fn)
    finetuning_run: bool = False
    finetuning_use_lora: bool = False
    finetuning_lora_min_rank: int = 32
    finetuning_lr: float = 0.0001
    finetuning_num_steps: int = 0
    finetuning_num_last_finetuned_modules: int = 8
    finetuning_use_rank_pattern: bool = False

    # lm_eval (not available offline)
    lm_eval_initial: bool = False
    lm_eval_tasks: Optional[list[str]] = None

    @pydantic.field_validator("decomposed_model_dtype")
    @classmethod
    def _known_dtype(cls, v: str) -> str:
        if v not in DTYPES:
            raise ValueError(f"decomposed_model_dtype must be one of {DTYPES}, got {v!r}")
        return v
