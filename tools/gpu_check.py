"""GPU bring-up / diagnostics for the eigensolver, low-rank forward and the drivers (run under
gpurun). Each case runs in its own subprocess under a timeout; one JSON line per case, all results
in gpurun_out/gpu_check.json.

    python tools/gpu_check.py [eigh] [lowrank] [falor] [dwain]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def spectrum_cov(d, seed=0, kind="step"):
    import torch
    from synth import streams
    g = torch.Generator().manual_seed(seed)
    if kind == "step":
        y = streams.step_spectrum_activations(4 * d, d, seed=seed)
    elif kind == "lowrank":  # rank-deficient covariance + damping: a large near-null cluster
        r = max(2, d // 4)
        y = torch.randn(4 * d, r, generator=g) @ torch.randn(r, d, generator=g)
    else:
        y = torch.randn(4 * d, d, generator=g) * torch.logspace(0, -3, d)
    c = (y.double().T @ y.double()) / y.shape[0]
    c = c + 0.01 * c.diagonal().mean() * torch.eye(d, dtype=torch.float64)
    return c


def run_eigh(cfg):
    import torch
    from ptdeco_b200 import linalg
    d, k = cfg["d"], cfg.get("k") or cfg["d"]
    dev = torch.device("cuda:0")
    c64 = spectrum_cov(d, cfg.get("seed", 0), cfg.get("kind", "step"))
    c32 = c64.float().to(dev)
    t0 = time.time()
    ev, u = linalg.eigh(c32, k=k)
    torch.cuda.synchronize()
    out = {"first_call_s": time.time() - t0}
    ref_ev, ref_u = torch.linalg.eigh(c32.double())
    out["nan"] = bool(torch.isnan(u).any().item() or torch.isnan(ev).any().item())
    out["eval_err_normwise"] = ((ev.double() - ref_ev).abs().max() / ref_ev.abs().max()).item()
    ud = u.double()
    out["orth_err"] = (ud.T @ ud - torch.eye(k, dtype=torch.float64, device=dev)).abs().max().item()
    lam_k = ref_ev[d - k:]
    out["resid_over_norm"] = ((c32.double() @ ud - ud * lam_k).abs().max() / ref_ev.abs().max()).item()
    for kk in sorted({max(1, d // 8), max(1, d // 4), max(1, d // 2)}):
        if kk <= k:
            s = torch.linalg.svdvals(ref_u[:, d - kk:].T @ ud[:, k - kk:])
            out[f"min_cos_k{kk}"] = s.min().item()
    if cfg.get("time"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        linalg.eigh(c32, k=k)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(cfg.get("iters", 3)):
            linalg.eigh(c32, k=k)
        e1.record()
        torch.cuda.synchronize()
        out["ms"] = e0.elapsed_time(e1) / cfg.get("iters", 3)
        if cfg.get("cusolver"):
            torch.linalg.eigh(c32)
            torch.cuda.synchronize()
            e0.record()
            torch.linalg.eigh(c32)
            e1.record()
            torch.cuda.synchronize()
            out["torch_cuda_eigh_ms"] = e0.elapsed_time(e1)
    return out


def run_lowrank(cfg):
    import torch
    from ptdeco_b200 import linalg
    dev = torch.device("cuda:0")
    n, fin, k, fout = cfg["n"], cfg["in"], cfg["k"], cfg["out"]
    dt = torch.bfloat16 if cfg.get("dtype", "bf16") == "bf16" else torch.float32
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, fin, generator=g).to(dt).to(dev)
    w1 = (torch.randn(k, fin, generator=g) / fin ** 0.5).to(dt).to(dev)
    w2 = (torch.randn(fout, k, generator=g) / k ** 0.5).to(dt).to(dev)
    b = torch.randn(fout, generator=g).to(dev) if cfg.get("bias", True) else None
    y = linalg.lowrank_forward(x, w1, w2, b)
    torch.cuda.synchronize()
    h = x.double() @ w1.double().T
    if dt == torch.bfloat16:
        h = h.to(torch.bfloat16).double()  # the intermediate is bf16 in both implementations
    ref = h @ w2.double().T + (b.double() if b is not None else 0)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    out = {"rel_err": err, "nan": bool(torch.isnan(y).any().item())}
    if cfg.get("time"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            linalg.lowrank_forward(x, w1, w2, b)
        e0.record()
        for _ in range(10):
            linalg.lowrank_forward(x, w1, w2, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        es = 2 if dt == torch.bfloat16 else 4
        out["ms"] = ms
        out["alg_gbs"] = (es * n * (fin + fout) + es * k * (fin + fout)) / ms / 1e6
        out["tflops"] = 2.0 * n * k * (fin + fout) / ms / 1e9
        seq = torch.nn.Sequential(torch.nn.Linear(fin, k, bias=False), torch.nn.Linear(k, fout)).to(dev).to(dt)
        for _ in range(3):
            seq(x)
        e0.record()
        for _ in range(10):
            seq(x)
        e1.record()
        torch.cuda.synchronize()
        out["torch_sequential_ms"] = e0.elapsed_time(e1) / 10
    return out


def run_falor(cfg):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import ptdeco_b200.falor as falor
    from synth import cases
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", f"falor_{cfg['name']}.json")))
    dev = torch.device("cuda:0")
    model, stream, kw = cases.falor_case(cfg["name"])
    trace = []
    if cfg.get("pregenerate", True):
        # the synthetic stream draws every batch with the CPU RNG (2.5 ms + a pageable H2D copy each);
        # that is the caller's data pipeline, not the library: materialise the batches on the device
        # first and time the decomposition on its own
        t0 = time.time()
        batches = [next(stream).to(dev) for _ in range(gold["stream_position"])]
        torch.cuda.synchronize()
        gen_s = time.time() - t0
        from synth.streams import IndexedStream
        stream = IndexedStream(lambda i: batches[i])
    else:
        gen_s = 0.0
    model.to(dev)
    torch.cuda.synchronize()
    t0 = time.time()
    dc = falor.decompose_in_place(module=model, device=dev, data_iterator=stream, trace=trace, **kw)
    torch.cuda.synchronize()
    out = {"seconds": time.time() - t0, "data_generation_seconds": gen_s, "stream_position": stream.position,
           "gold_position": gold["stream_position"], "n_trials": len(trace),
           "gold_trials": len(gold["trace"])}
    mism = []
    for t, g in zip(trace, gold["trace"]):
        if (t["name"], t["rank"]) != (g["name"], g["rank"]):
            mism.append({"mine": [t["name"], t["rank"]], "gold": [g["name"], g["rank"]]})
            break
    out["first_mismatch"] = mism
    diffs = []
    for idx, (t, g) in enumerate(zip(trace, gold["trace"])):
        if (t["name"], t["rank"]) != (g["name"], g["rank"]):
            break
        diffs.append((abs(t["nsr"] - g["nsr"]) / max(g["nsr"], 1e-9), idx, t["name"], t["rank"], t["nsr"], g["nsr"]))
    diffs.sort(reverse=True)
    out["worst_trials"] = [list(d[1:]) for d in diffs[:8]]
    out["n_matching_prefix"] = len(diffs)
    out["rel_diff_quantiles"] = [sorted(d[0] for d in diffs)[int(q * (len(diffs) - 1))] for q in (0.5, 0.9, 0.99)] if diffs else []
    out["max_rel_nsr_diff"] = max((abs(t["nsr"] - g["nsr"]) / max(g["nsr"], 1e-9) for t, g in zip(trace, gold["trace"])
                                   if (t["name"], t["rank"]) == (g["name"], g["rank"])), default=None)
    ranks = {n: c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels")) for n, c in dc.items()}
    granks = {n: c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels"))
              for n, c in gold["decompose_config"].items()}
    out["ranks_equal"] = ranks == granks
    out["n_decomposed"] = len(ranks)
    out["rank_diffs"] = {n: [ranks.get(n), granks.get(n)] for n in set(ranks) | set(granks) if ranks.get(n) != granks.get(n)}
    return out


def run_dwain(cfg):
    import torch
    import ptdeco_b200.dwain as dwain
    from synth import cases
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", f"dwain_{cfg['name']}.json")))
    dev = torch.device("cuda:0")
    model, stream, mstream, kw = cases.dwain_case(cfg["name"])
    trace = []
    t0 = time.time()
    dc = dwain.decompose_in_place(module=model.to(dev), device=dev, data_iterator=stream,
                                  metric_iterator=mstream, loss_fn=cases.dwain_loss_fn(cfg["name"]),
                                  finetune_fn=lambda m, d, n: m, trace=trace, **kw)
    torch.cuda.synchronize()
    out = {"seconds": time.time() - t0, "pos": [stream.position, mstream.position],
           "gold_pos": [gold["stream_position"], gold["metric_stream_position"]]}
    out["trace_equal"] = [(t["name"], t["rank"]) for t in trace] == [(g["name"], g["rank"]) for g in gold["trace"]]
    out["max_rel_nsr_diff"] = max((abs(t["nsr"] - g["nsr"]) / max(g["nsr"], 1e-9) for t, g in zip(trace, gold["trace"])), default=None)
    ranks = {n: c["modules"]["0"]["out_features"] for n, c in dc.items()}
    granks = {n: c["modules"]["0"]["out_features"] for n, c in gold["decompose_config"].items()}
    out["ranks_equal"] = ranks == granks
    out["ranks"] = ranks
    return out


def run_gemmacc(cfg):
    """fp32-split GEMM accuracy / SYRK speed as a function of the TMEM accumulation chunk."""
    import torch
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    L = nat.lib()
    dev = torch.device("cuda:0")
    L.ptdeco_debug_set(6, cfg["chunk"])
    g = torch.Generator().manual_seed(1)
    M, N, K = cfg.get("M", 512), cfg.get("N", 512), cfg["K"]
    a = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, K, generator=g).to(dev)
    c = linalg.gemm(a, False, b, False, M, N, K)
    ref = a.double() @ b.double().T
    out = {"rel_err_max": ((c.double() - ref).abs().max() / ref.abs().max()).item(),
           "rel_fro": ((c.double() - ref).norm() / ref.norm()).item(),
           "bias": ((c.double() - ref).mean() / ref.abs().mean()).item()}
    ab, bb = a.to(torch.bfloat16), b.to(torch.bfloat16)
    c = linalg.gemm(ab, False, bb, False, M, N, K)
    ref = ab.double() @ bb.double().T
    out["bf16_rel_fro"] = ((c.double() - ref).norm() / ref.norm()).item()
    n, d = 8192, 4096
    y = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev)
    acc = linalg.CovarianceAccumulator(d, dev)
    for _ in range(3):
        acc.update(y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        acc.update(y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["syrk_d4096_ms"] = ms
    out["syrk_tflops_alg"] = n * d * (d + 1) / ms / 1e9
    y32 = y.float()
    acc = linalg.CovarianceAccumulator(d, dev)
    for _ in range(2):
        acc.update(y32)
    e0.record()
    for _ in range(5):
        acc.update(y32)
    e1.record()
    torch.cuda.synchronize()
    out["syrk_f32_d4096_ms"] = e0.elapsed_time(e1) / 5
    return out


RUNNERS = {"gemmacc": run_gemmacc, "eigh": run_eigh, "lowrank": run_lowrank, "falor": run_falor, "dwain": run_dwain}


def sub(kind, cfg, timeout=300):
    cmd = [sys.executable, os.path.abspath(__file__), "--run", kind, json.dumps(cfg)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"error": "timeout"}
    for line in p.stdout.splitlines()[::-1]:
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    return {"error": "crash", "rc": p.returncode, "stderr": p.stderr[-1500:]}


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--run":
        print("RESULT " + json.dumps(RUNNERS[sys.argv[2]](json.loads(sys.argv[3]))))
        return
    what = sys.argv[1:] or ["eigh", "lowrank", "falor", "dwain"]
    results = []

    def rec(name, kind, cfg, timeout=300):
        r = sub(kind, cfg, timeout)
        results.append({"name": name, "cfg": cfg, "res": r})
        print(name, json.dumps(r), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
            json.dump(results, f, indent=1)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if "gemmacc" in what:
        for chunk in (16, 8, 4, 2, 1):
            rec(f"gemmacc_chunk{chunk}_K4096", "gemmacc", dict(chunk=chunk, K=4096))
    if "eigh" in what:
        rec("eigh_d2", "eigh", dict(d=2))
        rec("eigh_d10", "eigh", dict(d=10, kind="decay"))
        rec("eigh_d32", "eigh", dict(d=32))
        rec("eigh_d96", "eigh", dict(d=96))
        rec("eigh_d97", "eigh", dict(d=97, kind="decay"))
        rec("eigh_d128", "eigh", dict(d=128))
        rec("eigh_d192", "eigh", dict(d=192))
        rec("eigh_d200_k50", "eigh", dict(d=200, k=50))
        rec("eigh_d576_lowrank", "eigh", dict(d=576, kind="lowrank"))
        rec("eigh_d768", "eigh", dict(d=768, time=True))
        rec("eigh_d1000_k500", "eigh", dict(d=1000, k=500, kind="decay", time=True))
        rec("eigh_d2048", "eigh", dict(d=2048, time=True, cusolver=True))
        rec("eigh_d4096", "eigh", dict(d=4096, time=True, cusolver=True), timeout=600)
        rec("eigh_d4096_k2048", "eigh", dict(d=4096, k=2048, time=True), timeout=600)
    if "eighbig" in what:
        rec("eigh_d8192_k2048", "eigh", dict(d=8192, k=2048, time=True, iters=1), timeout=900)
        rec("eigh_d14336_k2048", "eigh", dict(d=14336, k=2048, time=True, iters=1, kind="lowrank"), timeout=1200)
    if "lowrank" in what:
        rec("lowrank_bf16_small", "lowrank", dict(n=256, **{"in": 192, "k": 48, "out": 160}))
        rec("lowrank_f32_small", "lowrank", dict(n=300, **{"in": 200, "k": 50, "out": 168}, dtype="f32"))
        rec("lowrank_bf16_4096_k128_n8192", "lowrank", dict(n=8192, **{"in": 4096, "k": 128, "out": 4096}, time=True))
        rec("lowrank_bf16_4096_k256_n8192", "lowrank", dict(n=8192, **{"in": 4096, "k": 256, "out": 4096}, time=True))
        rec("lowrank_bf16_4096_k1024_n8192", "lowrank", dict(n=8192, **{"in": 4096, "k": 1024, "out": 4096}, time=True))
        rec("lowrank_bf16_4096_k256_n128", "lowrank", dict(n=128, **{"in": 4096, "k": 256, "out": 4096}, time=True))
        rec("lowrank_bf16_ragged", "lowrank", dict(n=333, **{"in": 200, "k": 72, "out": 520}))
        rec("lowrank_bf16_4096_k128_n32768", "lowrank", dict(n=32768, **{"in": 4096, "k": 128, "out": 4096}, time=True))
        rec("lowrank_bf16_4096_k64_n32768", "lowrank", dict(n=32768, **{"in": 4096, "k": 64, "out": 4096}, time=True))
        rec("lowrank_bf16_4096_k128_n37888", "lowrank", dict(n=37888, **{"in": 4096, "k": 128, "out": 4096}, time=True, bias=False))
        rec("lowrank_bf16_4096_k128_n75776", "lowrank", dict(n=75776, **{"in": 4096, "k": 128, "out": 4096}, time=True, bias=False))
        rec("lowrank_bf16_14336_k128_n32768", "lowrank", dict(n=32768, **{"in": 4096, "k": 128, "out": 14336}, time=True, bias=False))
        rec("lowrank_bf16_14336x4096_k256_n8192", "lowrank", dict(n=8192, **{"in": 4096, "k": 256, "out": 14336}, time=True))
        rec("lowrank_bf16_4096_k128_n16", "lowrank", dict(n=16, **{"in": 4096, "k": 128, "out": 4096}, time=True))
    if "falor" in what:
        for n in ("mlp", "convmlp", "deit_small"):
            rec(f"falor_{n}", "falor", dict(name=n))
    if "falorbig" in what:
        rec("falor_deit_tiny", "falor", dict(name="deit_tiny"), timeout=1200)
        rec("falor_convnext_tiny", "falor", dict(name="convnext_tiny"), timeout=1200)
    if "dwain" in what:
        for n in ("llama_tiny", "llama_tiny_splits"):
            rec(f"dwain_{n}", "dwain", dict(name=n))


if __name__ == "__main__":
    main()
