"""One warm + one measured eigh call at a given size (run plain, then under ncu's launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ptdeco_b200 import linalg
from tools.gpu_check import spectrum_cov

d = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else d
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cov = spectrum_cov(d).float().cuda()
linalg.eigh(cov, k=k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    linalg.eigh(cov, k=k)
e1.record()
torch.cuda.synchronize()
print(f"eigh d={d} k={k}: {e0.elapsed_time(e1) / reps:.3f} ms")
