"""Eigensolver accuracy + timing: resident (shared-memory) tridiagonalisation against the blocked
panel kernel of round 1, per-phase cycle counters, and torch.linalg.eigh on the same GPU.

    python tools/eigh_bench.py [--sizes 96,768,...] [--rows 2,4,8] [--out gpurun_out/eigh_bench.json]
Each size runs in its own subprocess under a timeout (a hung cooperative kernel must not take the
box down).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(d: int, k: int, rows: list[int], cusolver: bool) -> dict:
    import torch

    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    from tools.gpu_check import spectrum_cov

    dev = torch.device("cuda:0")
    L = nat.lib()
    c64 = spectrum_cov(d, 0, "step")
    c32 = c64.float().to(dev)
    ref = torch.linalg.eigvalsh(c32.double())

    def acc():
        ev, u = linalg.eigh(c32, k=k)
        torch.cuda.synchronize()
        ud = u.double()
        lam = ev.double()
        return {
            "nan": bool(torch.isnan(u).any().item()),
            "eval_err": float((lam - ref).abs().max() / ref.abs().max()),
            "orth": float((ud.T @ ud - torch.eye(k, dtype=torch.float64, device=dev)).abs().max()),
            "resid": float((c32.double() @ ud - ud * lam[d - k:]).abs().max() / ref.abs().max()),
        }

    def timed(reps=3):
        linalg.eigh(c32, k=k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            linalg.eigh(c32, k=k)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"d": d, "k": k}
    L.ptdeco_debug_set(104, 0)  # no Jacobi: the general path at every size
    L.ptdeco_debug_set(102, 0)
    out["blocked"] = acc()
    out["blocked"]["ms"] = timed()
    for r in rows:
        L.ptdeco_debug_set(103, r)  # enables the resident kernel with this rows-per-CTA target
        key = f"resident_rows{r}"
        out[key] = acc()
        out[key]["ms"] = timed()
        L.ptdeco_debug_set(100, 1)
        linalg.eigh(c32, k=k)
        torch.cuda.synchronize()
        out[key]["phase_cycles"] = [int(L.ptdeco_debug_get(100 + i)) for i in range(10)]
        L.ptdeco_debug_set(100, 0)
    if d <= 96:
        L.ptdeco_debug_set(104, 96)
        out["jacobi"] = acc()
        out["jacobi"]["ms"] = timed()
        L.ptdeco_debug_set(104, 0)
    L.ptdeco_debug_set(106, 1)  # division-based (LAPACK-form) Sturm count instead of the product form
    out["sturm_ratio_form"] = acc()
    out["sturm_ratio_form"]["ms"] = timed()
    L.ptdeco_debug_set(106, 0)
    if d >= 1024:  # 4 lanes per eigenvalue in the multisection instead of 16
        L.ptdeco_debug_set(105, 1)
        out["bisect_4_lanes"] = {"ms": timed()}
        L.ptdeco_debug_set(105, 0)
    if cusolver:
        torch.linalg.eigh(c32)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.linalg.eigh(c32)
        e1.record()
        torch.cuda.synchronize()
        out["torch_cuda_eigh_ms"] = e0.elapsed_time(e1)
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="32,96,192,768,1024,2048,2560,4096")
    ap.add_argument("--rows", default="4")
    ap.add_argument("--kfrac", type=float, default=1.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "eigh_bench.json"))
    ap.add_argument("--one", type=int, default=0)
    ap.add_argument("--no-cusolver", action="store_true")
    args = ap.parse_args()
    rows = [int(x) for x in args.rows.split(",")]
    if args.one:
        d = args.one
        print(json.dumps(one(d, max(1, int(d * args.kfrac)), rows, not args.no_cusolver)))
        return
    results = []
    for d in [int(x) for x in args.sizes.split(",")]:
        cmd = [sys.executable, os.path.abspath(__file__), "--one", str(d), "--rows", args.rows,
               "--kfrac", str(args.kfrac)] + (["--no-cusolver"] if args.no_cusolver else [])
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            res = json.loads(line[-1]) if line else {"d": d, "error": (p.stderr or p.stdout)[-800:]}
        except subprocess.TimeoutExpired:
            res = {"d": d, "error": "timeout"}
        print(json.dumps(res), flush=True)
        results.append(res)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
