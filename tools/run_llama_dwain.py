"""End-to-end dwain.decompose_in_place on a Llama-3-8B-SHAPE decoder (random init, bf16,
synthetic tokens) -- BASELINE.json configs[2] as a whole-model wall-time. Works single process or
under torchrun (calibration sharded over ranks, eigensolves round-robin).

    python tools/run_llama_dwain.py --layers 32 --data-steps 8 --metric-steps 1 --splits 1
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import ptdeco_b200.dwain as dwain
from synth import models, streams


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--data-steps", type=int, default=8)
    ap.add_argument("--metric-steps", type=int, default=1)
    ap.add_argument("--splits", type=int, default=1)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--min-rank", type=int, default=32)
    ap.add_argument("--nsr", type=float, default=0.05)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    with torch.device(dev):
        model = models.LlamaLikeDecoder(layers=args.layers, init=False).to(torch.bfloat16)
    models.fast_init_(model, 271828)
    model.eval()
    build_s = time.time() - t0
    data = streams.IndexedStream(lambda i: streams.token_batch(2, i, 1, args.seq, 128256))
    metric = streams.IndexedStream(lambda i: streams.token_batch(3, i, 1, args.seq, 128256))
    trace = []
    torch.cuda.synchronize()
    t0 = time.time()
    cfg = dwain.decompose_in_place(
        module=model, device=dev, data_iterator=data, metric_iterator=metric,
        loss_fn=models.llama_ce_loss, finetune_fn=lambda m, d, n: m, num_data_steps=args.data_steps,
        num_metric_steps=args.metric_steps, blacklisted_module_names=["lm_head"],
        nsr_final_threshold=args.nsr, min_rank=args.min_rank, decompose_in_float64=True,
        precomputing_covariance_num_splits=args.splits, trace=trace,
        process_group="world" if world > 1 else None)
    torch.cuda.synchronize()
    wall = time.time() - t0
    if rank == 0:
        ranks = {n: c["modules"]["0"]["out_features"] for n, c in cfg.items()}
        hist = {}
        for r in ranks.values():
            hist[r] = hist.get(r, 0) + 1
        print("LLAMA_DWAIN " + json.dumps({
            "world": world, "layers": args.layers, "targets": 7 * args.layers, "decomposed": len(cfg),
            "trials": len(trace), "wall_s": wall, "model_build_s": build_s, "data_steps": args.data_steps,
            "metric_steps": args.metric_steps, "rank_histogram": hist,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
