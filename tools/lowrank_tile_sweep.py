"""Fused low-rank forward: time against the token rows per tile (ptdeco_debug_set key 207) to
calibrate the host's tile-shape cost model, next to nn.Sequential on cuBLAS.

    python tools/lowrank_tile_sweep.py > gpurun_out/lowrank_tile_sweep.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ptdeco_b200 import _native as nat
from ptdeco_b200 import linalg


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    L = nat.lib()
    g = torch.Generator(device=dev).manual_seed(1)
    out = []
    for n, fin, k, fout in ((8192, 4096, 128, 4096), (8192, 4096, 256, 4096), (16384, 4096, 128, 4096),
                            (32768, 4096, 128, 4096), (32768, 4096, 64, 4096), (37888, 4096, 128, 4096),
                            (8192, 4096, 128, 14336), (2048, 4096, 128, 4096), (4096, 768, 96, 3072)):
        x = torch.randn(n, fin, generator=g, device=dev).to(torch.bfloat16)
        w1 = (torch.randn(k, fin, generator=g, device=dev) / fin ** 0.5).to(torch.bfloat16)
        w2 = (torch.randn(fout, k, generator=g, device=dev) / k ** 0.5).to(torch.bfloat16)
        seq = torch.nn.Sequential(torch.nn.Linear(fin, k, bias=False), torch.nn.Linear(k, fout, bias=False)).to(dev).to(torch.bfloat16)
        with torch.no_grad():
            t_torch = timed(lambda: seq(x))
        row = {"shape": [n, fin, k, fout], "torch_ms": t_torch, "by_tile_m": {}}
        for tm in (0, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128):
            L.ptdeco_debug_set(207, tm)
            row["by_tile_m"][tm] = timed(lambda: linalg.lowrank_forward(x, w1, w2, None))
        L.ptdeco_debug_set(207, 0)
        best = min((v, t) for t, v in row["by_tile_m"].items() if t)
        row["best"] = {"tile_m": best[1], "ms": best[0], "model_ms": row["by_tile_m"][0]}
        print(json.dumps(row), flush=True)
        out.append(row)


if __name__ == "__main__":
    main()
