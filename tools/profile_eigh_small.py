import sys; sys.path.insert(0, "/root/repo")
import torch
from ptdeco_b200 import linalg, _native as nat
from tools.gpu_check import spectrum_cov
L = nat.lib()
cov = spectrum_cov(96).float().cuda()
for _ in range(3): linalg.eigh(cov)
torch.cuda.synchronize()
L.ptdeco_debug_set(100, 1)
linalg.eigh(cov); torch.cuda.synchronize()
cyc = [L.ptdeco_debug_get(100 + i) for i in range(5)]
L.ptdeco_debug_set(100, 0)
print(dict(zip(["warp-0 reflector (+sync)", "symv", "sync", "pv + update", "sync"], [c / 95 for c in cyc])), sum(cyc) / 95)
