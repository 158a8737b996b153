#!/usr/bin/env python3
"""Task dispatcher of the LLM example (reference: examples/trainer_llm/run.py:168-205).

    python examples/trainer_llm/run.py --config examples/trainer_llm/examples_config/decompose_dwain_llama_random.yaml \\
        --output-path /tmp/out

Writes the config copy, `decompose_config.json`, `decompose_state_dict.pt` and `summary.json` into the
output directory. Under torchrun (one rank per GPU) the decomposition runs sharded and rank 0 writes."""
from __future__ import annotations

import argparse
import logging
import os
import pathlib
import shutil
import sys
import time
from typing import Any

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import torch
import yaml

import ptdeco_b200 as ptdeco

import version

logger = logging.getLogger(__name__)


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--config", type=pathlib.Path)
    p.add_argument("--output-path", type=pathlib.Path)
    p.add_argument("--version", action="store_true")
    args = p.parse_args(argv)
    if not args.version and (args.config is None or args.output_path is None):
        p.error("--config and --output-path are required")
    return args


def read_config(path: pathlib.Path) -> dict[str, Any]:
    with open(path, "rt") as f:
        return yaml.safe_load(f)


def dispatch(config: dict[str, Any], output_path: pathlib.Path, process_group=None) -> Any:
    task = config.get("task")
    if task == "decompose_dwain":
        import run_decompose_dwain
        return run_decompose_dwain.main(config_raw=config, output_path=output_path, process_group=process_group)
    if task == "finetune":
        raise ValueError("task 'finetune' (the reference's run_finetune.py) needs `datasets` and `peft`, "
                         "which are not installed here; only 'decompose_dwain' is available")
    raise ValueError("config.task unspecified" if task is None else f"Unknown config.task={task}")


def main(args: argparse.Namespace) -> None:
    if args.version:
        print(f"Using ptdeco_b200 LLM trainer {version.__version__}")
        print(f"Using ptdeco_b200 {getattr(ptdeco, '__version__', 'dev')}")
        print(f"Using torch {torch.__version__}")
        return
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    start = time.perf_counter()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    group = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        group = "world"
    out = args.output_path if rank == 0 else args.output_path / f"rank{rank}"
    out.mkdir(exist_ok=True, parents=True)
    shutil.copy2(args.config, out / "config.yaml")
    dispatch(read_config(args.config), out, process_group=group)
    logger.info(f"Run took: {time.perf_counter() - start:.1f} s")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(parse_args())
