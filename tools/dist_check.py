"""Multi-GPU parity of the sharded calibration path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_check.py

Every rank builds the same seeded tiny Llama; the covariance precompute runs sharded over the
ranks (steps i = rank mod world, NCCL reduce to the owner, owner eigensolve, broadcast) and is
compared with the unsharded computation done redundantly on every rank: same subspaces (principal
angle cosines >= 0.9999 at the ranks the search consumes) and bit-identical results across ranks.
Then dwain.decompose_in_place with precompute splits must return the golden ranks on every rank.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import ptdeco_b200.dwain as dwain
import ptdeco_b200.dwain.decomposition as D
from ptdeco_b200 import parallel
from synth import cases


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    group = parallel.default_group()
    out = {"world": world}

    model, stream, _, kw = cases.dwain_case("llama_tiny")
    model.to(dev)
    names = D._get_decomposeable_submodule_names(model, ["lm_head"])
    sharded = D._precompute_covariance_matrix_decompositions(
        module=model, submodule_names=names, num_data_steps=8, data_iterator=stream, device=dev,
        decompose_in_float64=True, reduction_factor=0.5, group=group)
    model2, stream2, _, _ = cases.dwain_case("llama_tiny")
    model2.to(dev)
    single = D._precompute_covariance_matrix_decompositions(
        module=model2, submodule_names=names, num_data_steps=8, data_iterator=stream2, device=dev,
        decompose_in_float64=True, reduction_factor=0.5, group=None)
    assert stream.position == stream2.position == 8
    worst = 1.0
    for n in names:
        a, b = sharded[n].double(), single[n].double()
        k = a.shape[1]
        for kk in {k, max(1, k // 2)}:
            s = torch.linalg.svdvals(a[:, k - kk:].T @ b[:, k - kk:])
            worst = min(worst, s.min().item())
        if world > 1:
            ref = sharded[n].clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(ref, sharded[n]), f"rank {rank} holds a different U for {n}"
    out["min_cosine_sharded_vs_single"] = worst
    assert worst >= 0.9999, worst

    # deterministic mode (PTDECO_FLAG_DETERMINISTIC on every kernel call): bit-identical results
    # from run to run AND against the single-GPU computation: calibration step i is folded into
    # canonical shard i mod 8 whatever GPU ran it, and the owner adds the shards in index order
    from ptdeco_b200 import _native as nat
    nat.set_deterministic(True)
    reruns = []
    for g in (group, group, None):
        m_, s_, _, _ = cases.dwain_case("llama_tiny")
        m_.to(dev)
        reruns.append(D._precompute_covariance_matrix_decompositions(
            module=m_, submodule_names=names, num_data_steps=8, data_iterator=s_, device=dev,
            decompose_in_float64=True, reduction_factor=0.5, group=g))
    out["deterministic_reruns_bit_identical"] = all(torch.equal(reruns[0][n], reruns[1][n]) for n in names)
    out["deterministic_sharded_equals_single_gpu_bits"] = all(
        torch.equal(reruns[0][n], reruns[2][n]) for n in names)
    assert out["deterministic_reruns_bit_identical"]
    assert out["deterministic_sharded_equals_single_gpu_bits"]
    # ... and so is the whole decomposition: every trial's measured metrics, not just the ranks
    traces = []
    for g in ("world" if world > 1 else None, None):
        m_, s_, mm_, kw_ = cases.dwain_case("llama_tiny_splits")
        m_.to(dev)
        tr = []
        dwain.decompose_in_place(module=m_, device=dev, data_iterator=s_, metric_iterator=mm_,
                                 loss_fn=cases.dwain_loss_fn("llama_tiny_splits"),
                                 finetune_fn=lambda m, d, nn: m, process_group=g, trace=tr, **kw_)
        traces.append([(t["name"], t["rank"], t["nsr"], t["ppl_diff"], t["ppl_deco"], t["accepted"]) for t in tr])
    nat.set_deterministic(False)
    out["deterministic_trial_metrics_equal_single_gpu_bits"] = traces[0] == traces[1]
    out["deterministic_trials_compared"] = len(traces[0])
    assert traces[0] == traces[1], [x for x, y in zip(*traces) if x != y][:3]

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "dwain_llama_tiny_splits.json")))
    model3, s3, m3, kw3 = cases.dwain_case("llama_tiny_splits")
    model3.to(dev)
    cfg = dwain.decompose_in_place(module=model3, device=dev, data_iterator=s3, metric_iterator=m3,
                                   loss_fn=cases.dwain_loss_fn("llama_tiny_splits"),
                                   finetune_fn=lambda m, d, nn: m,
                                   process_group="world" if world > 1 else None, **kw3)
    ranks = {n: c["modules"]["0"]["out_features"] for n, c in cfg.items()}
    granks = {n: c["modules"]["0"]["out_features"] for n, c in gold["decompose_config"].items()}
    out["ranks_equal_golden"] = ranks == granks
    # the bf16 golden case (the headline dtype) through the sharded path as well
    gold16 = json.load(open(os.path.join(ROOT, "tests", "golden", "dwain_llama_tiny_bf16_splits.json")))
    model4, s4, m4, kw4 = cases.dwain_case("llama_tiny_bf16_splits")
    model4.to(dev)
    cfg16 = dwain.decompose_in_place(module=model4, device=dev, data_iterator=s4, metric_iterator=m4,
                                     loss_fn=cases.dwain_loss_fn("llama_tiny_bf16_splits"),
                                     finetune_fn=lambda m, d, nn: m,
                                     process_group="world" if world > 1 else None, **kw4)
    r16 = {n: c["modules"]["0"]["out_features"] for n, c in cfg16.items()}
    g16 = {n: c["modules"]["0"]["out_features"] for n, c in gold16["decompose_config"].items()}
    out["bf16_ranks_equal_golden"] = r16 == g16
    assert r16 == g16, (r16, g16)
    out["positions"] = [s3.position, m3.position, gold["stream_position"], gold["metric_stream_position"]]
    assert ranks == granks, (ranks, granks)
    # falor: layers are independent -> layer l is analysed by rank l mod world, the others skip its
    # batches; every rank must end with the golden config, the golden iterator position, and
    # bit-identical replacement modules
    import time

    import ptdeco_b200.falor as falor
    for fname in ("convmlp", "deit_small"):
        goldf = json.load(open(os.path.join(ROOT, "tests", "golden", f"falor_{fname}.json")))
        fm, fs, fkw = cases.falor_case(fname)
        fm.to(dev)
        ftrace = []
        t0 = time.perf_counter()
        fcfg = falor.decompose_in_place(module=fm, device=dev, data_iterator=fs, trace=ftrace,
                                        process_group="world" if world > 1 else None, **fkw)
        torch.cuda.synchronize()
        out[f"falor_{fname}_s"] = time.perf_counter() - t0
        fr = {n: c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels")) for n, c in fcfg.items()}
        gr = {n: c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels"))
              for n, c in goldf["decompose_config"].items()}
        assert fr == gr, (fr, gr)
        assert fs.position == goldf["stream_position"], (fs.position, goldf["stream_position"])
        assert [(t["name"], t["rank"]) for t in ftrace] == [(t["name"], t["rank"]) for t in goldf["trace"]]
        if world > 1:
            for k_, v_ in fm.state_dict().items():
                ref = v_.clone()
                dist.broadcast(ref, src=0)
                assert torch.equal(ref, v_), f"rank {rank}: {k_} differs from rank 0"
        out[f"falor_{fname}_ranks_equal_golden"] = True
    if world > 1:
        dist.barrier()
    if rank == 0:
        print("DIST_CHECK " + json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
