"""Where does falor's wall time go on a small vision model? (host vs device, per component)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptdeco_b200 import _graphs, linalg, utils
from synth import cases

torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "deit_tiny"
model, stream, kw = cases.falor_case(name)
model.to(dev).eval()

def wall(fn, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n

print("batch generation (host)      ms", wall(lambda: next(stream), 10))
xb = next(stream)
print("H2D .to(device)              ms", wall(lambda: xb.to(dev), 10))
x = xb.to(dev)
with torch.no_grad():
    print("eager forward                ms", wall(lambda: model(x)))
    gf = _graphs.GraphedForward(model)
    for _ in range(3): gf(x)
    print("graph enabled/entries        ", gf.enabled, len(gf._entries))
    print("graph replay forward         ms", wall(lambda: gf(x)))
    y0 = model(x); y1 = y0 + 0.01 * torch.randn_like(y0)
    print("nsr + kl metric kernels      ms", wall(lambda: (utils.calc_per_channel_noise_to_signal_ratio(x=y1, y=y0, non_channel_dim=(0,)), utils.calc_kl_loss(y1, y0))))
    w = torch.randn(768, 192, device=dev); u = torch.linalg.qr(torch.randn(768, 768, device=dev))[0]
    def factors():
        uk = u[:, 768 - 96:]
        w1 = linalg.factor_w1(w, uk)
        return linalg.deco_weight(uk, w1)
    print("factor_w1 + deco_weight      ms", wall(factors))
    lin = torch.nn.Linear(192, 768).to(dev)
    print("set_weight x2                ms", wall(lambda: (lin.weight.copy_(w), lin.weight.copy_(w))))
    acc = linalg.CovarianceAccumulator(768, dev)
    yy = torch.randn(985, 768, device=dev)
    print("syrk update (fp32, 985x768)  ms", wall(lambda: acc.update(yy, sub=lin.bias)))
    cov = acc.finalize(False, 0.01).clone()
    print("eigh d=768 k=191             ms", wall(lambda: linalg.eigh(cov, k=191), 5))
    print(".tolist() sync               ms", wall(lambda: torch.stack([y0.sum(), y1.sum()]).tolist()))
