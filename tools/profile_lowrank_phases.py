import sys, os
sys.path.insert(0, "/root/repo")
import torch
from ptdeco_b200 import _native as nat, linalg
L = nat.lib(); dev = torch.device("cuda:0")
names = ["setup", "GEMM 1 (h_full)", "H hand-over (rest)", "first Y tile", "remaining Y tiles", "store drain", "wait peer GEMM 1", "DSMEM stores"]
for (n, in_f, k, out_f) in [(8192, 4096, 128, 4096), (2048, 4096, 128, 4096), (16384, 4096, 128, 4096)]:
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, in_f, generator=g, device=dev).to(torch.bfloat16)
    w1 = torch.randn(k, in_f, generator=g, device=dev).to(torch.bfloat16)
    w2 = torch.randn(out_f, k, generator=g, device=dev).to(torch.bfloat16)
    for mode in (1, 0):
        L.ptdeco_debug_set(208, mode)
        for _ in range(3): linalg.lowrank_forward(x, w1, w2)
        torch.cuda.synchronize()
        L.ptdeco_debug_set(209, 1)
        reps = 10
        for _ in range(reps): linalg.lowrank_forward(x, w1, w2)
        torch.cuda.synchronize()
        cyc = [L.ptdeco_debug_get(210 + i) / reps for i in range(8)]
        L.ptdeco_debug_set(209, 0)
        print((n, in_f, k, out_f), "ksplit off" if mode else "ksplit auto", " | ".join(f"{a} {c:.0f}" for a, c in zip(names, cyc)), "total", sum(cyc))
