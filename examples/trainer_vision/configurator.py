"""Config schema of the vision trainer's `decompose_falor` task; field names are the reference's
(examples/trainer_vision/configurator.py:51-61 plus its data fields), so its YAML loads unchanged.
ImageNet through DALI and timm are not available offline: `imagenet_root_dir: synthetic` selects
seeded random images, `decompose_model_name: torchvision.<name>` a random-init torchvision model."""
from __future__ import annotations

from typing import Literal, Optional

import pydantic


class DecomposeFALORConfig(pydantic.BaseModel):
    model_config = pydantic.ConfigDict(extra="forbid")

    ptdeco_trainer_version: Optional[str] = None
    ptdeco_version: Optional[str] = None
    task: Literal["decompose_falor"]

    # data
    imagenet_root_dir: str
    trn_imagenet_classes_fname: Optional[str] = None
    val_imagenet_classes_fname: Optional[str] = None
    batch_size: int
    normalization: str = "imagenet"
    input_h_w: tuple[int, int]

    # model
    decompose_model_name: str

    # ptdeco.falor.decompose_in_place keyword arguments
    proportion_threshold: float
    blacklisted_modules: list[str]
    kl_final_threshold: float
    nsr_final_threshold: float
    num_data_steps: int
    num_metric_steps: int
    use_float64: bool = False
