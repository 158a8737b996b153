"""One launch each of the engine's K-major x K-major GEMM (CTA-pair kernel) and torch.matmul (cuBLAS)
at 8192^3 bf16, plus one SYRK launch at d = 14336 / N = 16384, for an ncu side-by-side:

    ncu --set full --clock-control none -o gpurun_out/gemm_ab python tools/micro/gemm_vs_cublas_once.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from ptdeco_b200 import linalg

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(8192, 8192, generator=g, device=dev).to(torch.bfloat16)
w = torch.randn(8192, 8192, generator=g, device=dev).to(torch.bfloat16)
a = linalg.linear_nt(x, w)
b = x @ w.T
torch.cuda.synchronize()
if "--syrk" in sys.argv:
    y = torch.randn(16384, 14336, generator=g, device=dev).to(torch.bfloat16)
    acc = linalg.CovarianceAccumulator(14336, dev)
    acc._syrk(y, None, 1.0 / 16384)
    torch.cuda.synchronize()
print("rel", ((a.float() - b.float()).norm() / b.float().norm()).item())
