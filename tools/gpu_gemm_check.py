"""GPU bring-up check of the tcgen05 GEMM engine through the C-ABI (run under gpurun).

Each case runs in its own subprocess under a timeout so a trap / hang in one descriptor variant
cannot poison the others. Prints one line per case; writes gpurun_out/gemm_check.json.

    python tools/gpu_gemm_check.py            # default cases (+ descriptor sweep on failure)
    python tools/gpu_gemm_check.py --one '{"a_mn":1,...}'
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(cfg: dict) -> dict:
    import torch

    from ptdeco_b200 import _native as nat

    L = nat.lib()
    for k, v in cfg.get("dbg", {}).items():
        L.ptdeco_debug_set(int(k), int(v))
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(cfg.get("seed", 1))
    M, N, K = cfg["M"], cfg["N"], cfg["K"]
    a_mn, b_mn = cfg["a_mn"], cfg["b_mn"]
    dt = torch.bfloat16 if cfg.get("dtype", "bf16") == "bf16" else torch.float32

    def make(rows, cols):
        if cfg.get("ints", True) and dt == torch.bfloat16:
            t = torch.randint(-3, 4, (rows, cols), generator=g).to(torch.float32)
        else:
            t = torch.randn(rows, cols, generator=g)
        return t.to(dt).to(dev)

    A = make(K, M) if a_mn else make(M, K)
    B = make(K, N) if b_mn else make(N, K)
    Af = (A.double().T if a_mn else A.double())  # [M,K]
    Bf = (B.double().T if b_mn else B.double())  # [N,K]
    ref = Af @ Bf.T
    C = torch.zeros(M, N, dtype=torch.float32, device=dev)
    acc = int(cfg.get("accumulate", 0))
    if acc:
        C.fill_(1.0)
        ref = ref + 1.0
    ws_bytes = L.ptdeco_gemm_workspace_bytes(nat.dtype_code(A), nat.dtype_code(B), M, N, K)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.ptdeco_gemm(A.data_ptr(), nat.dtype_code(A), a_mn, A.stride(0), B.data_ptr(),
                       nat.dtype_code(B), b_mn, B.stride(0), M, N, K, 1.0, None, C.data_ptr(),
                       nat.F32, C.stride(0), acc, ws.data_ptr(), ws.numel(), st)
    torch.cuda.synchronize()
    out = {"rc": rc}
    if rc == 0:
        err = (C.double() - ref).abs()
        scale = ref.abs().max().item() + 1e-30
        out["max_abs_err"] = err.max().item()
        out["rel_err"] = err.max().item() / scale
        out["frac_exact"] = (err < 1e-3 * scale).double().mean().item()
        # which 64x64 blocks of the output are right (helps decode a wrong descriptor)
        bm = []
        for i in range(0, min(M, 256), 64):
            row = []
            for j in range(0, min(N, 256), 64):
                row.append(int((err[i:i + 64, j:j + 64] < 1e-3 * scale).double().mean().item() * 100))
            bm.append(row)
        out["block_ok_pct"] = bm
        out["info"] = [L.ptdeco_debug_get(i) for i in range(5)]
    return out


def run_syrk(cfg: dict) -> dict:
    import torch

    from ptdeco_b200 import _native as nat

    L = nat.lib()
    for k, v in cfg.get("dbg", {}).items():
        L.ptdeco_debug_set(int(k), int(v))
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(cfg.get("seed", 1))
    n, d = cfg["n"], cfg["d"]
    dt = torch.bfloat16 if cfg.get("dtype", "bf16") == "bf16" else torch.float32
    Y = torch.randn(n, d, generator=g).to(dt).to(dev)
    C = torch.zeros(d, d, dtype=torch.float32, device=dev)
    cs = torch.zeros(d, dtype=torch.float32, device=dev)
    ws_bytes = L.ptdeco_syrk_workspace_bytes(nat.dtype_code(Y), n, d)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    steps = cfg.get("steps", 2)
    for _ in range(steps):
        rc = L.ptdeco_syrk_accumulate(Y.data_ptr(), nat.dtype_code(Y), n, d, Y.stride(0), None,
                                      C.data_ptr(), C.stride(0), cs.data_ptr(), 1.0 / n,
                                      ws.data_ptr(), ws.numel(), st)
        if rc:
            return {"rc": rc}
    rc = L.ptdeco_cov_finalize(C.data_ptr(), C.stride(0), d, cs.data_ptr(), steps, 0, 0.01, None, st)
    torch.cuda.synchronize()
    Yd = Y.double()
    ref = (Yd.T @ Yd) / n
    ref = ref + 0.01 * ref.diagonal().mean() * torch.eye(d, device=dev, dtype=torch.float64)
    err = (C.double() - ref)
    out = {"rc": rc, "rel_fro": (err.norm() / ref.norm()).item(),
           "max_abs_over_max": (err.abs().max() / ref.abs().max()).item(),
           "sym_err": (C - C.T).abs().max().item(),
           "colsum_err": (cs.double() / steps - Yd.mean(0)).abs().max().item(),
           "info": [L.ptdeco_debug_get(i) for i in range(5)]}
    if cfg.get("time", False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            L.ptdeco_syrk_accumulate(Y.data_ptr(), nat.dtype_code(Y), n, d, Y.stride(0), None,
                                     C.data_ptr(), C.stride(0), None, 1.0 / n, ws.data_ptr(),
                                     ws.numel(), st)
        iters = 10
        e0.record()
        for _ in range(iters):
            L.ptdeco_syrk_accumulate(Y.data_ptr(), nat.dtype_code(Y), n, d, Y.stride(0), None,
                                     C.data_ptr(), C.stride(0), None, 1.0 / n, ws.data_ptr(),
                                     ws.numel(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out["ms"] = ms
        out["tflops_alg"] = n * d * (d + 1) / ms / 1e9
    return out


def sub(kind: str, cfg: dict, timeout: int = 120) -> dict:
    cmd = [sys.executable, os.path.abspath(__file__), "--" + kind, json.dumps(cfg)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"error": "timeout"}
    for line in p.stdout.splitlines()[::-1]:
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    return {"error": "crash", "rc": p.returncode, "stderr": p.stderr[-400:]}


def main() -> None:
    if len(sys.argv) >= 3 and sys.argv[1] in ("--one", "--syrk"):
        cfg = json.loads(sys.argv[2])
        res = run_one(cfg) if sys.argv[1] == "--one" else run_syrk(cfg)
        print("RESULT " + json.dumps(res))
        return

    results = []

    def rec(name, kind, cfg):
        r = sub(kind, cfg)
        results.append({"name": name, "cfg": cfg, "res": r})
        print(name, json.dumps(r), flush=True)
        return r

    def ok(r):
        return r.get("rc") == 0 and r.get("rel_err", 1.0) < 1e-5

    base = dict(M=128, N=128, K=64, a_mn=0, b_mn=0)
    r_kk = rec("kk_128x128x64", "one", base)
    r_mm = rec("mnmn_128x128x64", "one", dict(base, a_mn=1, b_mn=1))
    if not ok(r_mm):
        # descriptor sweep for MN-major: dbg keys 2 = LBO, 3 = SBO (bytes)
        for lbo, sbo in [(1024, 8192), (8192, 128), (128, 8192), (1024, 1024), (8192, 8192),
                         (16, 1024), (1024, 16), (2048, 1024), (1024, 2048)]:
            rec(f"mnmn_sweep_lbo{lbo}_sbo{sbo}", "one",
                dict(base, a_mn=1, b_mn=1, dbg={"2": lbo, "3": sbo}))
    if not ok(r_kk):
        for lbo, sbo in [(1024, 1024), (0, 1024), (16, 128), (128, 1024), (1, 1024)]:
            rec(f"kk_sweep_lbo{lbo}_sbo{sbo}", "one", dict(base, dbg={"4": lbo, "5": sbo}))
    rec("kk_256x256x256", "one", dict(M=256, N=256, K=256, a_mn=0, b_mn=0))
    rec("mnmn_256x256x256", "one", dict(M=256, N=256, K=256, a_mn=1, b_mn=1))
    rec("kmn_256x256x256", "one", dict(M=256, N=256, K=256, a_mn=0, b_mn=1))
    rec("mnk_256x256x256", "one", dict(M=256, N=256, K=256, a_mn=1, b_mn=0))
    rec("mnmn_ragged_200x328x1000", "one", dict(M=200, N=328, K=1000, a_mn=1, b_mn=1))
    rec("kk_ragged_200x328x1000", "one", dict(M=200, N=328, K=1000, a_mn=0, b_mn=0))
    rec("kk_tile128_forced", "one", dict(M=384, N=384, K=512, a_mn=0, b_mn=0, dbg={"0": 128}))
    rec("mnmn_acc_splitk", "one", dict(M=256, N=256, K=4096, a_mn=1, b_mn=1, accumulate=1))
    rec("mnmn_big_persistent", "one", dict(M=2048, N=2048, K=512, a_mn=1, b_mn=1))
    rec("mnmn_fp32_split", "one", dict(M=256, N=320, K=512, a_mn=1, b_mn=1, dtype="f32"))
    rec("kk_fp32_split", "one", dict(M=256, N=320, K=512, a_mn=0, b_mn=0, dtype="f32"))
    rec("syrk_bf16_d32", "syrk", dict(n=2048, d=32))
    rec("syrk_bf16_d768", "syrk", dict(n=1576, d=768))
    rec("syrk_f32_d768", "syrk", dict(n=1576, d=768, dtype="f32"))
    rec("syrk_f32_d1000", "syrk", dict(n=3136, d=1000, dtype="f32"))
    rec("syrk_bf16_d4096", "syrk", dict(n=8192, d=4096, time=True))
    rec("syrk_bf16_d4096_t128", "syrk", dict(n=8192, d=4096, time=True, dbg={"0": 128}))
    rec("syrk_f32_d4096", "syrk", dict(n=8192, d=4096, dtype="f32", time=True))
    rec("syrk_bf16_d14336", "syrk", dict(n=8192, d=14336, time=True, steps=1))
    rec("syrk_bf16_d4096_n2048", "syrk", dict(n=2048, d=4096, time=True))
    rec("syrk_bf16_d1024_n2048", "syrk", dict(n=2048, d=1024, time=True))
    rec("syrk_bf16_d8192", "syrk", dict(n=8192, d=8192, time=True, steps=1))
    rec("syrk_bf16_d14336_chunk1e6", "syrk", dict(n=8192, d=14336, time=True, steps=1, dbg={"6": 1000000}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_check.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
