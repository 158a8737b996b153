"""Focused check + timing of the tcgen05 engine's SYRK instance (the bench's dominant kernel).

    python tools/syrk_bench.py [--json gpurun_out/syrk_bench.json]

For each d: one accuracy check of C += Y^T Y / N against an fp64 torch product (relative Frobenius
error over the lower triangle), then CUDA-event timing of back-to-back launches at N = 8192 staged
tokens. Also times a K-major x K-major GEMM (the layer-forward instance) next to torch.matmul.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ptdeco_b200 import linalg


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


class NvmlSampler:
    """SM clock / board power sampled through NVML every few ms from a thread (nvidia-smi as a
    subprocess is far too slow to see inside a sub-second timed region)."""

    def __init__(self, index=0, period_s=0.005):
        import threading
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.period = period_s
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1e3))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join()

    def summary(self, skip_frac=0.3):
        rows = self.rows[int(len(self.rows) * skip_frac):]  # the power cap takes ~0.3 s to bite
        if not rows:
            return {"samples": 0}
        mhz = sorted(r[0] for r in rows)
        w = sorted(r[1] for r in rows)
        return {"samples": len(rows), "sm_mhz_median": mhz[len(mhz) // 2], "sm_mhz_min": mhz[0],
                "watts_median": w[len(w) // 2], "watts_max": w[-1]}


def sustained(fn, flop, seconds=2.0):
    """Back-to-back launches for ~`seconds` with NVML clock / power samples taken meanwhile:
    separates tensor-pipe utilisation from the clock the power cap allows."""
    ms1 = timed(fn, iters=5, warm=2)
    iters = max(10, int(seconds * 1e3 / ms1))
    with NvmlSampler() as smp:
        ms = timed(fn, iters=iters, warm=0)
    tf = flop / ms / 1e9
    rec = {"ms": ms, "tflops": tf, "iters": iters}
    rec.update(smp.summary())
    if rec.get("sm_mhz_median"):
        rec["pipe_util_at_clock"] = tf / (148 * 8192 * rec["sm_mhz_median"] * 1e-6)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "syrk_bench.json"))
    ap.add_argument("--dims", default="1024,4096,8192,14336")
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--sustained", action="store_true")
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--pair", type=int, default=-1, help="1: CTA-pair kernel (default path), 0: single-CTA kernel")
    args = ap.parse_args()
    if args.pair == 0:
        from ptdeco_b200 import _native as nat
        nat.lib().ptdeco_debug_set(7, 1)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    dev = torch.device("cuda:0")
    out = []
    n = args.tokens
    for d in [int(x) for x in args.dims.split(",")]:
        g = torch.Generator(device=dev).manual_seed(d)
        y = torch.randn(n, d, generator=g, device=dev).to(torch.bfloat16)
        acc = linalg.CovarianceAccumulator(d, dev)
        acc.update(y)
        acc.flush()
        torch.cuda.synchronize()
        got = torch.tril(acc.C.double())
        rows = min(n, 2048 if d > 8192 else n)  # bound the fp64 reference
        if rows < n:
            acc2 = linalg.CovarianceAccumulator(d, dev)
            acc2.update(y[:rows].contiguous())
            acc2.flush()
            got = torch.tril(acc2.C.double())
            del acc2
        yd = y[:rows].double()
        ref = torch.tril(yd.T @ yd / rows)
        rel = ((got - ref).norm() / ref.norm()).item()
        del yd, ref, got
        ms = timed(lambda: acc._syrk(y, None, 1.0 / n))
        rec = {"op": "syrk_bf16", "d": d, "N": n, "ms": ms, "rel_err": rel,
               "tflops_alg": n * d * (d + 1) / ms / 1e9}
        if "bf16_tflops" in peaks:
            rec["frac_tensor_burst"] = rec["tflops_alg"] / peaks["bf16_tflops"]
        out.append(rec)
        print(json.dumps(rec), flush=True)
        del acc, y
        torch.cuda.empty_cache()
    if args.sustained:
        from ptdeco_b200 import _native as nat
        L = nat.lib()

        def knobs(pair=1, sleep=0, band=0):
            L.ptdeco_debug_set(7, 0 if pair else 1)
            L.ptdeco_debug_set(8, sleep)
            L.ptdeco_debug_set(9, band)

        d = 14336
        g = torch.Generator(device=dev).manual_seed(d)
        y = torch.randn(2 * n, d, generator=g, device=dev).to(torch.bfloat16)
        acc = linalg.CovarianceAccumulator(d, dev)
        for cfg in ({"pair": 0}, {"pair": 1, "band": 8}, {"pair": 1, "band": 8, "sleep": 300},
                    {"pair": 1, "band": 8, "sleep": 300, "chunk": 64},
                    {"pair": 1, "band": 8, "sleep": 300, "tokens": 2 * n}):
            cfg = dict(cfg)
            nt = cfg.pop("tokens", n)
            L.ptdeco_debug_set(6, cfg.pop("chunk", 0))
            knobs(**cfg)
            yy = y[:nt]
            rec = {"op": "syrk_bf16_sustained", "d": d, "N": nt, "cfg": cfg}
            rec.update(sustained(lambda: acc._syrk(yy, None, 1.0 / nt), nt * d * (d + 1.0), args.seconds))
            out.append(rec)
            print(json.dumps(rec), flush=True)
        L.ptdeco_debug_set(6, 0)
        del acc, y
        torch.cuda.empty_cache()
        x = torch.randn(8192, 8192, generator=g, device=dev).to(torch.bfloat16)
        w = torch.randn(8192, 8192, generator=g, device=dev).to(torch.bfloat16)
        for cfg in ({"pair": 0}, {"pair": 1, "band": 8}, {"pair": 1, "band": 8, "sleep": 300},
                    {"pair": 1, "band": 16, "sleep": 300}):
            knobs(**cfg)
            rec = {"op": "gemm_nt_bf16_sustained_ours", "cfg": cfg}
            rec.update(sustained(lambda: linalg.linear_nt(x, w), 2.0 * 8192 ** 3, args.seconds))
            out.append(rec)
            print(json.dumps(rec), flush=True)
        knobs()
        rec = {"op": "gemm_nt_bf16_sustained_torch"}
        rec.update(sustained(lambda: x @ w.T, 2.0 * 8192 ** 3, args.seconds))
        out.append(rec)
        print(json.dumps(rec), flush=True)
        del x, w
    # plain GEMM instance (X W^T) next to cuBLAS
    for (m, nn, k) in ((8192, 8192, 8192), (8192, 4096, 4096)):
        g = torch.Generator(device=dev).manual_seed(m + nn + k)
        x = torch.randn(m, k, generator=g, device=dev).to(torch.bfloat16)
        w = torch.randn(nn, k, generator=g, device=dev).to(torch.bfloat16)
        ours = linalg.linear_nt(x, w)
        ref = x @ w.T
        rel = ((ours.float() - ref.float()).norm() / ref.float().norm()).item()
        ms = timed(lambda: linalg.linear_nt(x, w))
        ms_t = timed(lambda: x @ w.T)
        rec = {"op": "gemm_nt_bf16", "M": m, "N": nn, "K": k, "ms": ms, "tflops": 2.0 * m * nn * k / ms / 1e9,
               "torch_ms": ms_t, "torch_tflops": 2.0 * m * nn * k / ms_t / 1e9, "rel_vs_torch": rel}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    with open(args.json, "w") as f:
        json.dump({"peaks": peaks, "points": out}, f, indent=1)


if __name__ == "__main__":
    main()
