"""Accuracy and timing of the eigensolver with the lower-triangle (sym) panel symv on and off.

    python tools/eigh_sym_check.py [--big]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ptdeco_b200 import _native as nat
from ptdeco_b200 import linalg
from tools.gpu_check import spectrum_cov


def check(cov, k):
    ev, u = linalg.eigh(cov, k=k)
    torch.cuda.synchronize()
    c64 = cov.double()
    ref = torch.linalg.eigvalsh(c64)
    lam = ev.double()
    ev_err = float((lam - ref).abs().max() / ref.abs().max())
    ud = u.double()
    lk = lam[-k:]
    res = float((c64 @ ud - ud * lk).abs().max() / ref.abs().max())
    orth = float((ud.T @ ud - torch.eye(k, dtype=torch.float64, device=cov.device)).abs().max())
    return ev_err, res, orth


def timed(cov, k, reps):
    linalg.eigh(cov, k=k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        linalg.eigh(cov, k=k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    L = nat.lib()
    big = "--big" in sys.argv
    for d, k in ((200, 200), (1000, 300), (2048, 2048), (4096, 1024)):
        cov = spectrum_cov(d).float().cuda()
        for name, thr in (("full-row symv", 0), ("lower-triangle symv", 128)):
            L.ptdeco_debug_set(101, thr)
            ev_err, res, orth = check(cov, k)
            print(json.dumps({"d": d, "k": k, "mode": name, "eval_err": ev_err, "residual": res, "orth": orth}), flush=True)
    for d, k, reps in ((4096, 4096, 2), (8192, 2048, 1)) + (((14336, 2048, 1),) if big else ()):
        cov = spectrum_cov(d).float().cuda()
        for name, thr in (("full-row symv", 0), ("lower-triangle symv", 128), ("default (m >= 5120)", 5120)):
            L.ptdeco_debug_set(101, thr)
            print(json.dumps({"d": d, "k": k, "mode": name, "ms": timed(cov, k, reps)}), flush=True)
    L.ptdeco_debug_set(101, 5120)


if __name__ == "__main__":
    main()
