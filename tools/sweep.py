"""Microbench sweep of BASELINE.json configs[4]: SYRK, eigh (top-k) and low-rank forward at
d = 768 / 4096 / 8192 / 14336 / 28672, k = d/8 ... d/2, with the roofline each point sits on.
Writes gpurun_out/sweep.json (copy to profiles/).

    python tools/sweep.py [--max-d 28672]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ptdeco_b200 import linalg
from synth import streams


def timed(fn, iters=5, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-d", type=int, default=28672)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm, tf = float(peaks["hbm_gbs"]), float(peaks["bf16_tflops"])
    dev = torch.device("cuda:0")
    out = []
    for d in (768, 4096, 8192, 14336, 28672):
        if d > args.max_d:
            continue
        n = 8192
        g = torch.Generator(device=dev).manual_seed(d)
        y = torch.randn(n, d, generator=g, device=dev).to(torch.bfloat16)
        acc = linalg.CovarianceAccumulator(d, dev)
        ms = timed(lambda: acc.update(y), iters=5 if d < 20000 else 2)
        out.append({"op": "syrk_bf16", "d": d, "N": n, "ms": ms, "tflops_alg": n * d * (d + 1) / ms / 1e9,
                    "frac_tensor_burst": n * d * (d + 1) / ms / 1e9 / tf, "bound": "tensor"})
        print(json.dumps(out[-1]), flush=True)
        y32 = y.float() if d <= 8192 else None
        if y32 is not None:
            acc32 = linalg.CovarianceAccumulator(d, dev)
            ms = timed(lambda: acc32.update(y32), iters=3)
            out.append({"op": "syrk_fp32_bf16x3", "d": d, "N": n, "ms": ms,
                        "tflops_alg": n * d * (d + 1) / ms / 1e9, "bound": "tensor (6 passes + staging)"})
            print(json.dumps(out[-1]), flush=True)
            del acc32, y32
        # eigh on a step-spectrum covariance (>= 4d tokens accumulated)
        acc = linalg.CovarianceAccumulator(d, dev)
        for i in range(max(1, 4 * d // n)):
            acc.update(streams.step_spectrum_activations(n, d, seed=i, device="cuda").to(torch.bfloat16))
        cov = acc.finalize(False, 0.01).clone()
        del acc, y
        ks = (d // 8, d // 4, d // 2) if d <= 14336 else (d // 8,)
        for k in ks:
            t0 = time.time()
            ev, u = linalg.eigh(cov, k=k)
            torch.cuda.synchronize()
            first = time.time() - t0
            ms = timed(lambda: linalg.eigh(cov, k=k), iters=1 if d >= 8192 else 3, warm=0)
            ud = u[:, -min(k, 512):].double()
            orth = (ud.T @ ud - torch.eye(ud.shape[1], dtype=torch.float64, device=dev)).abs().max().item()
            flop = (4.0 / 3 + 2.0 * k / d + 4.0 / 3 * k / d) * d ** 3
            out.append({"op": "eigh_topk", "d": d, "k": k, "ms": ms, "first_call_ms": 1e3 * first,
                        "gflops_convention": flop / ms / 1e6, "orth_err_last512": orth,
                        "sytrd_bytes_lower_bound_ms": (4.0 / 3) * d ** 3 / hbm / 1e6,
                        "bound": "L2/latency (d<=4096), HBM symv (d>=8192)"})
            print(json.dumps(out[-1]), flush=True)
            del ev, u, ud
        del cov
        torch.cuda.empty_cache()
        # low-rank forward, square layer in = out = d
        for k in (d // 8, d // 4, d // 2):
            for nn in (16, 128, 8192):
                if d >= 28672 and nn == 8192:
                    continue
                x = torch.randn(nn, d, generator=g, device=dev).to(torch.bfloat16)
                w1 = (torch.randn(k, d, generator=g, device=dev) / d ** 0.5).to(torch.bfloat16)
                w2 = (torch.randn(d, k, generator=g, device=dev) / k ** 0.5).to(torch.bfloat16)
                ms = timed(lambda: linalg.lowrank_forward(x, w1, w2, None), iters=5)
                byt = 2.0 * nn * 2 * d + 2.0 * k * 2 * d
                flop = 2.0 * nn * k * 2 * d
                t_mem, t_tc = byt / hbm / 1e6, flop / tf / 1e9
                out.append({"op": "lowrank_forward_bf16", "d": d, "k": k, "N": nn, "ms": ms,
                            "fused": k <= 256, "gbs_alg": byt / ms / 1e6, "tflops": flop / ms / 1e9,
                            "bound": "hbm" if t_mem >= t_tc else "tensor",
                            "frac_of_bound": max(t_mem, t_tc) / ms})
                print(json.dumps(out[-1]), flush=True)
                del x, w1, w2
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
        json.dump({"peaks": peaks, "points": out}, f, indent=1)


if __name__ == "__main__":
    main()
