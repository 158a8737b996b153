"""Fused low-rank forward with and without the k-split clusters (PTDECO_B200_FUSED_KSPLIT /
ptdeco_debug_set(208, ...): 0 auto, 1 off, 2 whenever possible): error against an fp32 torch
product of the bf16 operands, and CUDA-event time next to nn.Sequential on cuBLAS."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ptdeco_b200 import _native as nat
from ptdeco_b200 import linalg


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # inputs are smaller than the L2: start every launch cold
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    dev = torch.device("cuda:0")
    L = nat.lib()
    shapes = [(8192, 4096, 128, 4096), (8192, 4096, 64, 4096), (8192, 4096, 128, 14336), (4096, 4096, 128, 4096),
              (2048, 4096, 128, 4096), (2048, 4096, 32, 4096), (1000, 768, 96, 3072), (8192, 14336, 128, 4096),
              (16384, 4096, 128, 4096), (5000, 4096, 100, 1024), (8192, 128, 64, 4096)]
    out = []
    for (n, in_f, k, out_f) in shapes:
        g = torch.Generator(device=dev).manual_seed(n + k)
        x = torch.randn(n, in_f, generator=g, device=dev).to(torch.bfloat16)
        w1 = (torch.randn(k, in_f, generator=g, device=dev) / in_f ** 0.5).to(torch.bfloat16)
        w2 = (torch.randn(out_f, k, generator=g, device=dev) / k ** 0.5).to(torch.bfloat16)
        b = torch.randn(out_f, generator=g, device=dev)
        h = (x.float() @ w1.float().T).to(torch.bfloat16).float()
        ref = h @ w2.float().T + b
        rec = {"shape": [n, in_f, k, out_f]}
        for name, knob in (("off", 1), ("auto", 0), ("forced", 2)):
            L.ptdeco_debug_set(208, knob)
            y = linalg.lowrank_forward(x, w1, w2, b)
            torch.cuda.synchronize()
            rec[name + "_rel_err"] = ((y.float() - ref).norm() / ref.norm()).item()
            rec[name + "_ms"] = timed(lambda: linalg.lowrank_forward(x, w1, w2, b))
        L.ptdeco_debug_set(208, 0)
        seq = torch.nn.Sequential(torch.nn.Linear(in_f, k, bias=False), torch.nn.Linear(k, out_f)).to(dev).to(torch.bfloat16)
        with torch.no_grad():
            seq[0].weight.copy_(w1); seq[1].weight.copy_(w2); seq[1].bias.copy_(b)
            rec["torch_ms"] = timed(lambda: seq(x))
        out.append(rec)
        print(json.dumps(rec), flush=True)
        assert max(rec["off_rel_err"], rec["auto_rel_err"], rec["forced_rel_err"]) < 1e-2, rec
    with open(os.path.join(ROOT, "gpurun_out", "lowrank_ksplit_check.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
