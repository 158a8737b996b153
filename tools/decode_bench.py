"""Decode-size (N <= 128) low-rank forward: parity against torch and timing of the single-launch
weight-streaming kernel next to the generic path and to nn.Sequential(Linear, Linear) on cuBLAS.

    python tools/decode_bench.py [--json gpurun_out/decode_bench.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ptdeco_b200 import linalg


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timed_graph(fn, iters=20):
    """GPU time per call with the host out of the picture: `iters` calls captured into one CUDA
    graph and replayed (decode loops run under CUDA graphs in serving). None if capture fails."""
    try:
        fn()
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(iters):
                    fn()
        torch.cuda.synchronize()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (3 * iters)
    except Exception as exc:  # noqa: BLE001
        print("graph capture failed:", repr(exc)[:200], flush=True)
        torch.cuda.synchronize()
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "decode_bench.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6550.0
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    out = []
    # ---- parity on ragged shapes (n, in, k, out, bias)
    shapes = [(1, 256, 32, 256, False), (5, 320, 96, 1000, True), (16, 4096, 512, 4096, True),
              (33, 768, 200, 3072, True), (128, 4096, 1024, 4096, False), (100, 1024, 1000, 520, True),
              (16, 14336, 1792, 4096, True)]
    worst = 0.0
    for (n, fin, k, fout, has_b) in shapes:
        x = torch.randn(n, fin, generator=g, device=dev).to(torch.bfloat16)
        w1 = (torch.randn(k, fin, generator=g, device=dev) / fin ** 0.5).to(torch.bfloat16)
        w2 = (torch.randn(fout, k, generator=g, device=dev) / k ** 0.5).to(torch.bfloat16)
        b = torch.randn(fout, generator=g, device=dev) if has_b else None
        y = linalg.lowrank_forward(x, w1, w2, b)
        torch.cuda.synchronize()
        h = (x.double() @ w1.double().T).to(torch.bfloat16).double()  # bf16 intermediate like the module pair
        ref = h @ w2.double().T + (b.double() if has_b else 0.0)
        err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
        worst = max(worst, err)
        rec = {"check": [n, fin, k, fout, has_b], "max_rel_err": err}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    print(json.dumps({"worst_rel_err": worst, "ok": worst < 1.5e-2}), flush=True)
    # ---- timing
    dims = [(4096, 512), (4096, 1024), (4096, 2048), (8192, 1024), (8192, 4096), (14336, 1792),
            (14336, 7168), (768, 96), (768, 384)]
    if not args.quick:
        dims += [(28672, 3584), (28672, 14336)]
    for (d, k) in dims:
        for n in (1, 16, 128):
            x = torch.randn(n, d, generator=g, device=dev).to(torch.bfloat16)
            w1 = (torch.randn(k, d, generator=g, device=dev) / d ** 0.5).to(torch.bfloat16)
            w2 = (torch.randn(d, k, generator=g, device=dev) / k ** 0.5).to(torch.bfloat16)
            os.environ.pop("PTDECO_B200_NO_DECODE", None)
            ms = timed(lambda: linalg.lowrank_forward(x, w1, w2, None))
            gms = timed_graph(lambda: linalg.lowrank_forward(x, w1, w2, None))
            os.environ["PTDECO_B200_NO_DECODE"] = "1"
            ms_generic = timed(lambda: linalg.lowrank_forward(x, w1, w2, None))
            gms_generic = timed_graph(lambda: linalg.lowrank_forward(x, w1, w2, None))
            os.environ.pop("PTDECO_B200_NO_DECODE", None)
            seq = torch.nn.Sequential(torch.nn.Linear(d, k, bias=False), torch.nn.Linear(k, d, bias=False)
                                      ).to(dev).to(torch.bfloat16)
            with torch.no_grad():
                ms_torch = timed(lambda: seq(x))
                gms_torch = timed_graph(lambda: seq(x))
            byt = 2.0 * n * 2 * d + 2.0 * k * 2 * d
            rec = {"op": "lowrank_forward_decode", "d": d, "k": k, "N": n, "us": 1e3 * ms,
                   "us_generic_path": 1e3 * ms_generic, "us_torch_sequential": 1e3 * ms_torch,
                   "graph_us": None if gms is None else 1e3 * gms,
                   "graph_us_generic_path": None if gms_generic is None else 1e3 * gms_generic,
                   "graph_us_torch_sequential": None if gms_torch is None else 1e3 * gms_torch,
                   "gbs_alg": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / hbm}
            if gms:
                rec["graph_frac_hbm"] = byt / gms / 1e6 / hbm
            out.append(rec)
            print(json.dumps(rec), flush=True)
            del x, w1, w2, seq
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    with open(args.json, "w") as f:
        json.dump({"hbm_gbs": hbm, "points": out}, f, indent=1)


if __name__ == "__main__":
    main()
