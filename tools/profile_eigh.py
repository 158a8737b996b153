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
if os.environ.get("PANEL_PROF"):
    from ptdeco_b200 import _native as nat
    L = nat.lib()
    L.ptdeco_debug_set(100, 1)
    linalg.eigh(cov, k=k)
    torch.cuda.synchronize()
    cyc = [L.ptdeco_debug_get(100 + i) for i in range(10)]
    L.ptdeco_debug_set(100, 0)
    names = ["P1 column update", "barrier A", "v fill + V writes", "partial dots + atomics", "barrier B",
             "P3 w rows", "column gather + scalars", "symv row items", "symv segment sums", "P3 gather"]
    tot = sum(cyc)
    for n_, c in zip(names, cyc):
        print(f"  panel phase {n_:22s} {c / 1e6:9.2f} Mcycles  {100.0 * c / max(tot, 1):5.1f} %  {c / d / 1.0:8.0f} cycles/column")
