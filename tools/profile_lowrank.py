"""A few calls of the fused low-rank forward at one shape (run plain, then under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ptdeco_b200 import linalg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
k = int(sys.argv[2]) if len(sys.argv) > 2 else 128
fin = fout = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, fin, generator=g, device="cuda").to(torch.bfloat16)
w1 = (torch.randn(k, fin, generator=g, device="cuda") / 64).to(torch.bfloat16)
w2 = (torch.randn(fout, k, generator=g, device="cuda") / 8).to(torch.bfloat16)
for _ in range(3):
    y = linalg.lowrank_forward(x, w1, w2, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = linalg.lowrank_forward(x, w1, w2, None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"lowrank n={n} k={k}: {ms:.4f} ms, {(2 * n * (fin + fout) + 2 * k * (fin + fout)) / ms / 1e6:.0f} GB/s algorithmic")
