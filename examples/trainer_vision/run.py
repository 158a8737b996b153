#!/usr/bin/env python3
"""Task dispatcher of the vision example (reference: examples/trainer_vision/run.py).

    python examples/trainer_vision/run.py --config examples/trainer_vision/examples_config/decompose_falor_convnext.yaml \\
        --output-path /tmp/out
"""
from __future__ import annotations

import argparse
import logging
import pathlib
import shutil
import sys
from typing import Any

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import yaml


def dispatch(config: dict[str, Any], output_path: pathlib.Path) -> Any:
    task = config.get("task")
    if task == "decompose_falor":
        import run_decompose_falor
        return run_decompose_falor.main(config_raw=config, output_path=output_path)
    if task in ("decompose_dwain", "decompose_lockd", "finetune"):
        raise ValueError(f"task {task!r} of the reference's vision trainer is not part of this example "
                         "(see examples/trainer_llm for dwain)")
    raise ValueError("config.task unspecified" if task is None else f"Unknown config.task={task}")


def main(argv=None) -> None:
    p = argparse.ArgumentParser()
    p.add_argument("--config", type=pathlib.Path, required=True)
    p.add_argument("--output-path", type=pathlib.Path, required=True)
    args = p.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    args.output_path.mkdir(exist_ok=True, parents=True)
    shutil.copy2(args.config, args.output_path / "config.yaml")
    with open(args.config, "rt") as f:
        dispatch(yaml.safe_load(f), args.output_path)


if __name__ == "__main__":
    main()
