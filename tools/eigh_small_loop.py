import sys; sys.path.insert(0, "/root/repo")
import torch
from ptdeco_b200 import linalg
from tools.gpu_check import spectrum_cov
for d in (64, 96, 128):
    cov = spectrum_cov(d).float().cuda()
    for _ in range(3): linalg.eigh(cov)
    # keep the GPU busy so the clocks are up: a big GEMM right before, then 50 back-to-back solves
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    for _ in range(20): a @ a
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20): a @ a
    e0.record()
    for _ in range(50): linalg.eigh(cov)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 50
    e0.record()
    for _ in range(50): torch.linalg.eigh(cov)
    e1.record(); torch.cuda.synchronize()
    print(f"d={d}: {t:.3f} ms per eigh (back-to-back, warm clocks); cuSOLVER {e0.elapsed_time(e1)/50:.3f} ms")
