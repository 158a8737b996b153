"""BASELINE.json configs[3]: the per-layer decomposition pipeline at Llama-3-70B layer shapes
(q/o 8192x8192, k/v 1024x8192, gate/up 28672x8192, down 8192x28672; random-init bf16 weights,
synthetic step-spectrum activations), target layers distributed round-robin over the ranks. No
whole-model forward: 141 GB of bf16 weights leave no room for it next to the accumulators
(SURVEY.md 8e), so every target is processed on its own, as the sharded run would after calibration:

  covariance (tcgen05 SYRK; input-side route for in < out)  ->  eigensolve (top-k)  ->  geometric
  rank descent k = full/2, full/4, ... >= min_rank: factors W1 = Uk^T W, two-factor forward (K7),
  per-channel NSR of the layer output against the full-rank layer; smallest rank under the
  threshold wins (the layer-local part of D:333-537; the perplexity gates need the whole model).

    python tools/run_llama70b_layers.py [--blocks 1] [--tokens 16384] [--reference-formulation]
    torchrun --nproc-per-node 8 tools/run_llama70b_layers.py --blocks 8

Prints one JSON line with per-target timings and the projection to the 80-block model.
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

from ptdeco_b200 import linalg, parallel, utils
from synth import streams

TARGETS = [("q_proj", 8192, 8192), ("k_proj", 1024, 8192), ("v_proj", 1024, 8192), ("o_proj", 8192, 8192),
           ("gate_proj", 28672, 8192), ("up_proj", 28672, 8192), ("down_proj", 8192, 28672)]  # (name, out, in)
N_BLOCKS_70B = 80


def process_target(name, out_f, in_f, tokens, batch, min_rank, nsr_thr, seed, dev):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    g = torch.Generator(device=dev).manual_seed(seed)
    w = (torch.randn(out_f, in_f, generator=g, device=dev) / in_f ** 0.5).to(torch.bfloat16)
    full_rank = min(in_f, out_f)
    k_max = max(1, full_rank // 2)
    input_side = linalg.use_input_side(in_f, out_f, k_max)
    d = in_f if input_side else out_f
    acc = linalg.CovarianceAccumulator(d, dev, defer_rows=linalg.default_defer_rows(d, 2))
    xs = []
    for i in range(tokens // batch):  # activations with a decaying spectrum (rank decisions have margin)
        xs.append(streams.step_spectrum_activations(batch, in_f, seed=seed + i, device=dev).to(torch.bfloat16))
    torch.cuda.synchronize()
    ev[0].record()
    for x in xs:  # calibration: layer forward on the engine + covariance update
        if input_side:
            acc.update(x)
        else:
            acc.update(linalg.linear_nt(x, w))
    cov = acc.finalize(use_mean=False, damp_factor=0.0 if input_side else 0.01)
    ev[1].record()
    if input_side:
        u = linalg.eigvecs_from_input_covariance(cov, w, k_max)
    else:
        _, u = linalg.eigh(cov, k=k_max)
    ev[2].record()
    del acc, cov
    x = xs[0]
    y_orig = linalg.linear_nt(x, w)
    trials = []
    rank, best = full_rank, None
    while rank > min_rank:
        rank = int(rank * 0.5)
        uk = u[:, u.shape[1] - rank:].to(torch.bfloat16).contiguous()
        w1 = linalg.factor_w1(w, uk)
        y_deco = linalg.lowrank_forward(x, w1, uk, None)
        nsr = utils.calc_per_channel_noise_to_signal_ratio(y=y_orig, x=y_deco, non_channel_dim=(0,))
        trials.append((rank, nsr))
    ev[3].record()
    measured = [(r, float(v)) for r, v in trials]  # one host sync for the layer
    for r, v in measured:
        if v < nsr_thr:
            best = r
    fwd_ms = None
    if best is not None:  # decomposed-layer forward at the chosen rank, prefill batch
        uk = u[:, u.shape[1] - best:].to(torch.bfloat16).contiguous()
        w1 = linalg.factor_w1(w, uk)
        for _ in range(2):
            linalg.lowrank_forward(x, w1, uk, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            linalg.lowrank_forward(x, w1, uk, None)
        e1.record()
        torch.cuda.synchronize()
        fwd_ms = e0.elapsed_time(e1) / 5
    ev[4].record()
    torch.cuda.synchronize()
    rec = {"name": name, "out": out_f, "in": in_f, "covariance_d": d, "input_side": input_side,
           "calib_ms": ev[0].elapsed_time(ev[1]), "eig_ms": ev[1].elapsed_time(ev[2]),
           "rank_search_ms": ev[2].elapsed_time(ev[3]), "trials": measured, "rank": best,
           "lowrank_forward_ms": fwd_ms, "dense_forward_tokens": batch}
    del u, w, xs
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1, help="decoder blocks to process (70B has 80)")
    ap.add_argument("--tokens", type=int, default=16384, help="calibration tokens per target")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--min-rank", type=int, default=32)
    ap.add_argument("--nsr", type=float, default=0.05)
    ap.add_argument("--reference-formulation", action="store_true",
                    help="output-side covariance everywhere (d = 28672 for gate/up), as the reference")
    args = ap.parse_args()
    if args.reference_formulation:
        os.environ["PTDECO_B200_INPUT_SIDE"] = "0"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    todo = [(b, t) for b in range(args.blocks) for t in range(len(TARGETS))]
    mine = [(i, bt) for i, bt in enumerate(todo) if parallel.owner_of(i, world) == rank]
    torch.cuda.synchronize()
    t0 = time.time()
    recs = []
    for i, (b, t) in mine:
        name, out_f, in_f = TARGETS[t]
        rec = process_target(f"layers.{b}.{name}", out_f, in_f, args.tokens, args.batch, args.min_rank,
                             args.nsr, 1000 * b + t, dev)
        rec["owner"] = rank
        recs.append(rec)
    torch.cuda.synchronize()
    wall = time.time() - t0
    wall_max = parallel.max_over_ranks(wall, dev) if world > 1 else wall
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, recs)
        recs = [r for part in gathered for r in part]
    if rank == 0:
        per_block = sum(r["calib_ms"] + r["eig_ms"] + r["rank_search_ms"] for r in recs) / max(1, args.blocks)
        print("LLAMA70B_LAYERS " + json.dumps({
            "world": world, "blocks": args.blocks, "targets": len(todo), "tokens_per_target": args.tokens,
            "formulation": "reference (output-side)" if args.reference_formulation else "input-side for in < out",
            "wall_s": wall_max, "gpu_ms_per_block": per_block,
            "projected_s_80_blocks_on_this_world": per_block * N_BLOCKS_70B / world / 1e3,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "targets_detail": recs}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
