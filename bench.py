"""Benchmark of the calibration hot path (BASELINE.json: "calib tokens/s", configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-e2e] [--no-extra]

Workload (config.workload): dwain calibration of a Llama-3-8B-shape decoder -- per step one
2048-token bf16 sequence per GPU is folded into the fp32 output covariances of all 224 target
Linears (32 x {q 4096, k 1024, v 1024, o 4096, gate 14336, up 14336, down 4096}); tokens are
sharded over GPUs (weak scaling), the only exchange is the final d x d reduction, timed separately.

  value   tokens/s with the layer outputs already resident in HBM; batches are staged and folded
          in up to 16384 tokens per tcgen05 SYRK launch (the accumulator RMW costs 8 d^2 bytes per
          launch whatever N is); the final flush is inside the timed region
  e2e     tokens/s through the public API path (ptdeco_b200.dwain covariance-computing modules
          installed in a random-init Llama-3-8B-shape model): pinned host token ids -> H2D ->
          full model forward (layer forwards on the tcgen05 GEMM engine, one SYRK per accumulator:
          q/k/v and gate/up share theirs) -> D2H of a per-step checksum; the deferred SYRK staging
          is flushed inside the timed region
  extra   eigh at d = 768 / 2048 / 4096 next to torch.linalg.eigh on the same GPU (cuSOLVER) and
          on the host CPU; the reference's einsum on device="cuda" (cuBLAS) for the calibration
          step; fused low-rank forward next to nn.Sequential on cuBLAS; whole-model
          decompose_in_place wall times (falor DeiT-tiny / ConvNeXt-tiny, dwain Llama-3-8B shape)
  strong_scaling_e2e   (N >= 1) fixed 512 x 2048 tokens of calibration through the public API:
          sharded forwards -> lower-triangle NCCL reduce -> round-robin eigensolves -> broadcast
  roofline    tensor-bound: algorithmic N*d*(d+1) FLOP / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference   the reference's _update_Eyyt_in_place arithmetic (oracle port,
          numpy fp32 on all host cores) on a bounded sample: the 7 targets of ONE decoder layer,
          scaled by 1/32 to the whole model
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYER_DIMS = [("q_proj", 4096), ("k_proj", 1024), ("v_proj", 1024), ("o_proj", 4096),
              ("gate_proj", 14336), ("up_proj", 14336), ("down_proj", 4096)]
N_LAYERS = 32
SEQ = 2048
METRIC = "calib tokens/s"
UNIT = "tokens/s"


def alg_flops_per_token() -> float:
    return float(N_LAYERS * sum(d * (d + 1) for _, d in LAYER_DIMS))


def load_peaks() -> tuple[dict, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML every 5 ms while the timed region
    runs (nvidia-smi as a subprocess returns 2-3 samples per second: too slow for a sub-second
    region); falls back to nvidia-smi when pynvml is unavailable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index = index
        self.rows: list[tuple[float, float, int]] = []  # (sm MHz, watts, reason bits)
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nv = pynvml
        except Exception:
            self._nv = None

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                return index
        return index

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.rows.append((float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)),
                                      nv.nvmlDeviceGetPowerUsage(self._h) / 1e3,
                                      int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))))
                    self._stop.wait(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                c = [x.strip() for x in out.split(",")]
                bits = 0
                for (name, bit), col in zip(self.REASONS, (3, 5, 4, 6)):
                    if len(c) > col and c[col].lower().startswith("active"):
                        bits |= bit
                self.rows.append((float(c[0]), float(c[2]), bits))
                self.max_mhz = float(c[1])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        # samples "under load": board power above half of the maximum seen
        wmax = max((r[1] for r in self.rows), default=0.0)
        load = [r for r in self.rows if r[1] >= 0.5 * wmax] or self.rows
        sm = sorted(r[0] for r in load)
        bits = 0
        for r in self.rows:
            bits |= r[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [name for name, bit in self.REASONS if bits & bit],
                "samples": len(self.rows), "watts_max": wmax or None,
                "source": "nvml" if self._nv is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_sample(threads: int) -> tuple[float, str]:
    """One bounded sample of the reference arithmetic (D:147-152, oracle port) on the host:
    fp32 y^T y / N for the 7 targets of one decoder layer, N = 2048. Returns (tokens/s scaled to the
    32-layer model, description)."""
    import torch

    from oracle import torch_cpu as R

    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(1314159)
    t = 0.0
    for _, d in LAYER_DIMS:
        y = torch.randn(SEQ, d, generator=gen)
        acc = torch.zeros(d, d)
        t0 = time.perf_counter()
        R.update_Eyyt_in_place(acc, y)
        t += time.perf_counter() - t0
    return SEQ / (N_LAYERS * t), ("oracle port of _update_Eyyt_in_place (torch CPU fp32 einsum, all host threads) on "
                                  "the 7 targets of 1 of 32 decoder layers, 2048 tokens; tokens/s scaled by 1/32")


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_sample(cores)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sample = cpu_sample(cores)
        vals.append(v)
    dt = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(n_gpus: int) -> dict:
    return {"workload": "BASELINE configs[2]: dwain calibration, Llama-3-8B-shape decoder (random init, bf16), "
                        "2048-token synthetic sequences, all 224 target Linears",
            "tokens_per_step_per_gpu": SEQ, "global_tokens_per_step": SEQ * n_gpus,
            "tokens_per_syrk_launch": "up to 16384 (8 steps staged), see DESIGN.md",
            "parallelism": f"dp{n_gpus} (tokens sharded, d x d fp32 reduction at the end)",
            "l2": "inputs (176 MB) + accumulators (59 GB) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------ extras
def _timed_ms(fn, reps: int = 3, warm: int = 1) -> float:
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def extra_eigh(dev, gen, world: int) -> dict:
    """metric (iii): eigh ms per size; torch.linalg.eigh(device="cuda") is the library-GPU
    comparator of BASELINE.md section 3, the host CPU is the reference's own path."""
    import torch

    from ptdeco_b200 import linalg
    out = {}
    for d in (768, 2048, 4096):
        y = torch.randn(4 * d, d, generator=gen, device=dev) * torch.logspace(0, -2, d, device=dev)
        acc = linalg.CovarianceAccumulator(d, dev)
        acc.update(y)
        cov = acc.finalize(False, 0.01).clone()
        row = {"ms_full": _timed_ms(lambda: linalg.eigh(cov)),
               "ms_top_half": _timed_ms(lambda: linalg.eigh(cov, k=d // 2)),
               "torch_cuda_cusolver_ms": _timed_ms(lambda: torch.linalg.eigh(cov), reps=2)}
        if world == 1 and d == 4096:
            c = cov.cpu()
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = time.perf_counter()
            torch.linalg.eigh(c)
            row["host_cpu_fp32_ms"] = 1e3 * (time.perf_counter() - t0)
            row["host_cores"] = os.cpu_count()
        if world == 1 and d == 2048:  # the reference's decompose_in_float64 path (D:155-163 on fp64)
            c = cov.double().cpu()
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = time.perf_counter()
            torch.linalg.eigh(c)
            row["host_cpu_fp64_ms"] = 1e3 * (time.perf_counter() - t0)
            row["host_cores"] = os.cpu_count()
        out[f"d{d}"] = row
        del cov, acc, y
    return out


def extra_reference_gpu_syrk(dev, acts) -> dict:
    """The reference's _update_Eyyt_in_place (D:147-152: Eyyt += einsum(y, y) / N) with its tensors
    on device="cuda", i.e. torch / cuBLAS: one decoder layer's 7 targets, scaled to 32 layers."""
    import torch
    accs = {name: torch.zeros(d, d, device=dev) for name, d in LAYER_DIMS}

    def step():
        for name, _ in LAYER_DIMS:
            y = acts[name]
            accs[name] += torch.einsum("bp,bq->pq", y, y) / y.shape[0]

    ms = _timed_ms(step, reps=5, warm=2)
    return {"tokens_per_s": SEQ / (N_LAYERS * ms * 1e-3), "ms_per_decoder_layer": ms,
            "what": "reference arithmetic on device=cuda (torch einsum -> cuBLAS bf16 GEMM, full square, "
                    "bf16-rounded per-step product), 7 targets of 1 of 32 layers, scaled by 1/32"}


def extra_lowrank(dev, gen, peaks) -> dict:
    import torch

    from ptdeco_b200 import linalg
    out = {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf = float(peaks.get("bf16_tflops", 1590.0))
    for n, fin, k, fout in ((37888, 4096, 128, 4096), (32768, 4096, 128, 4096), (8192, 4096, 128, 4096),
                            (8192, 4096, 512, 4096), (8192, 4096, 1024, 4096), (8192, 4096, 2048, 14336)):
        x = torch.randn(n, fin, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
        w1 = (torch.randn(k, fin, generator=gen, device=dev) / fin ** 0.5).to(torch.bfloat16)
        w2 = (torch.randn(fout, k, generator=gen, device=dev) / k ** 0.5).to(torch.bfloat16)
        ms = _timed_ms(lambda: linalg.lowrank_forward(x, w1, w2, None), reps=10, warm=3)
        seq = torch.nn.Sequential(torch.nn.Linear(fin, k, bias=False),
                                  torch.nn.Linear(k, fout, bias=False)).to(dev).to(torch.bfloat16)
        with torch.no_grad():
            ms_t = _timed_ms(lambda: seq(x), reps=10, warm=3)
        alg_bytes = 2.0 * n * (fin + fout) + 2.0 * k * (fin + fout)
        flops = 2.0 * n * k * (fin + fout)
        t_hbm, t_tc = alg_bytes / (hbm * 1e9), flops / (tf * 1e12)
        out[f"N{n}_in{fin}_k{k}_out{fout}"] = {
            "ms": ms, "torch_sequential_cublas_ms": ms_t, "bound": "hbm" if t_hbm >= t_tc else "tensor",
            "frac_of_bound": max(t_hbm, t_tc) / (ms * 1e-3), "achieved_gbs": alg_bytes / ms / 1e6,
            "achieved_tflops": flops / ms / 1e9}
        if (n, k) == (8192, 128):  # the reference's decomposed module on the host (fp32, all cores)
            torch.set_num_threads(os.cpu_count() or 1)
            seq_cpu = seq.float().cpu()
            x_cpu = x[:2048].float().cpu()
            with torch.no_grad():
                seq_cpu(x_cpu)
                t0 = time.perf_counter()
                seq_cpu(x_cpu)
                dt = time.perf_counter() - t0
            out[f"N{n}_in{fin}_k{k}_out{fout}"]["host_cpu_fp32_ms_scaled_from_2048_rows"] = 1e3 * dt * n / 2048
            out[f"N{n}_in{fin}_k{k}_out{fout}"]["host_cores"] = os.cpu_count()
        del x, w1, w2, seq
    return out


class _DeviceStream:
    """Pre-generated batches resident on the device (the synthetic CPU RNG stream costs more than
    the decomposition itself); `position` like synth.streams.IndexedStream."""

    def __init__(self, make, count: int, dev):
        self.items = [make(i).to(dev) for i in range(count)]
        self.position = 0

    def __iter__(self):
        return self

    def __next__(self):
        item = self.items[self.position]
        self.position += 1
        return item


def extra_model_runs(dev, with_dwain8b: bool, world: int = 1) -> dict:
    """metric (i): decompose_in_place wall time per model through the public API. With world > 1
    EVERY rank runs this (process_group="world": falor layers and dwain's calibration steps /
    rank trials are sharded); times are the max over ranks."""
    import torch

    import ptdeco_b200.dwain as dwain
    import ptdeco_b200.falor as falor
    from ptdeco_b200 import parallel
    from synth import cases, models, streams
    out = {}
    pg = "world" if world > 1 else None
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for name in ("deit_tiny", "convnext_tiny"):
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", f"falor_{name}.json")))
        model, stream, kw = cases.falor_case(name)
        dstream = _DeviceStream(stream.make, gold["stream_position"], dev)
        model.to(dev)
        trace = []
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cfg = falor.decompose_in_place(module=model, device=dev, data_iterator=dstream, trace=trace,
                                       process_group=pg, **kw)
        torch.cuda.synchronize()
        wall = parallel.max_over_ranks(time.perf_counter() - t0, dev)
        ranks = {n: c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels")) for n, c in cfg.items()}
        granks = {n: c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels"))
                  for n, c in gold["decompose_config"].items()}
        out[f"falor_{name}"] = {"wall_s": wall, "n_gpus": world, "targets_decomposed": len(cfg),
                                "rank_trials": len(trace), "batches": dstream.position,
                                "ranks_equal_reference_golden": ranks == granks}
        del model, dstream
        torch.cuda.empty_cache()
    if with_dwain8b:
        with torch.device(dev):
            model = models.LlamaLikeDecoder(init=False).to(torch.bfloat16)
        models.fast_init_(model, 271828)
        model.eval()
        data = streams.IndexedStream(lambda i: streams.token_batch(2, i, 1, SEQ, 128256))
        metric = streams.IndexedStream(lambda i: streams.token_batch(3, i, 1, SEQ, 128256))
        trace = []
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cfg = dwain.decompose_in_place(
            module=model, device=dev, data_iterator=data, metric_iterator=metric,
            loss_fn=models.llama_ce_loss, finetune_fn=lambda m, d, n: m, num_data_steps=8,
            num_metric_steps=1, blacklisted_module_names=["lm_head"], nsr_final_threshold=0.05,
            min_rank=32, decompose_in_float64=True, precomputing_covariance_num_splits=1, trace=trace,
            process_group=pg)
        torch.cuda.synchronize()
        wall = parallel.max_over_ranks(time.perf_counter() - t0, dev)
        hist: dict = {}
        for c in cfg.values():
            r = c["modules"]["0"]["out_features"]
            hist[r] = hist.get(r, 0) + 1
        out["dwain_llama3_8b_shape"] = {
            "wall_s": wall, "n_gpus": world, "targets": 224, "targets_decomposed": len(cfg),
            "rank_trials": len(trace), "num_data_steps": 8, "num_metric_steps": 1,
            "rank_histogram": hist, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        del model
        torch.cuda.empty_cache()
    return out


def strong_scaling_leg(dev, rank: int, world: int, n_layers: int, sequences: int) -> dict:
    """The design's real multi-GPU path, wall-timed: a FIXED number of 2048-token sequences is
    calibrated through ptdeco_b200.dwain's precompute (sharded forwards of a Llama-3-8B-shape
    replica per rank, SYRK per accumulator), the partial covariances are reduced (lower triangles,
    NCCL, asynchronously), the eigensolves run round-robin and the top-k blocks are broadcast."""
    import torch

    import ptdeco_b200.dwain.decomposition as D
    from ptdeco_b200 import parallel
    from synth import models, streams

    group = parallel.default_group()
    with torch.device(dev):
        model = models.LlamaLikeDecoder(layers=n_layers, init=False).to(torch.bfloat16)
    models.fast_init_(model, 271828)
    model.eval()
    names = D._get_decomposeable_submodule_names(model, ["lm_head"])
    gen = torch.Generator().manual_seed(1314159)
    ids = torch.randint(0, 128256, (sequences, 1, SEQ), generator=gen).to(dev)
    data = iter({"input_ids": ids[i]} for i in range(sequences))
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    u = D._precompute_covariance_matrix_decompositions(
        module=model, submodule_names=names, num_data_steps=sequences, data_iterator=data, device=dev,
        decompose_in_float64=True, reduction_factor=0.5, group=group)
    check = float(sum(float(v[0, 0]) for v in list(u.values())[:4]))  # D2H: forces completion
    torch.cuda.synchronize()
    wall = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    del model, u
    torch.cuda.empty_cache()
    return {"tokens": sequences * SEQ, "sequences": sequences, "wall_s": wall,
            "tokens_per_s": sequences * SEQ / wall, "targets": len(names), "n_gpus": world,
            "scaling": "strong", "check": check,
            "path": "dwain._precompute_covariance_matrix_decompositions: sharded calibration forwards -> "
                    "lower-triangle NCCL reduce to round-robin owners -> eigensolves -> broadcast of U[:, -k:]"}


# ------------------------------------------------------------------------------------ GPU arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ptdeco_b200")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--layers", type=int, default=N_LAYERS, help="debug: fewer decoder layers")
    ap.add_argument("--no-model-runs", action="store_true", help="skip the whole-model wall times in extra")
    ap.add_argument("--no-dwain8b", action="store_true", help="skip the 8B-shape dwain wall time in extra")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling end-to-end leg")
    ap.add_argument("--strong-sequences", type=int, default=512)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a rank that fails inside a sharded extra must not leave the others waiting for ever
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=420))
    nat.lib()
    n_layers = args.layers
    peaks, peak_src = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: SYRK over resident layer outputs -----------------------------------
    g = torch.Generator(device=dev).manual_seed(1314159 + rank)
    acts = {name: torch.randn(SEQ, d, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
            for name, d in LAYER_DIMS}
    # staging depth: the library default (up to 16384 tokens = 8 steps per launch), shortened to a
    # divisor of the step count when there is one >= 4, so that the timed region ends on a full
    # launch instead of a short remainder (a short launch pays the same accumulator traffic)
    per = max(c for c in range(1, 9) if args.steps % c == 0)
    if per < 4:
        per = 8
    accs = [[linalg.CovarianceAccumulator(d, dev, defer_rows=min(linalg.default_defer_rows(d, 2), per * SEQ))
             for _, d in LAYER_DIMS] for _ in range(n_layers)]

    def syrk_step():
        for layer in accs:
            for (name, _), acc in zip(LAYER_DIMS, layer):
                acc.update(acts[name])

    def flush_all():
        for layer in accs:
            for acc in layer:
                acc.flush()

    for _ in range(args.warmup):
        syrk_step()
    flush_all()
    launches0 = sum(acc.launches for layer in accs for acc in layer)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            syrk_step()
        flush_all()  # pending staged rows are folded in inside the timed region
        e1.record()
        barrier()
    gpu_launches = sum(acc.launches for layer in accs for acc in layer) - launches0
    ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    tokens_per_step = SEQ * world
    value = tokens_per_step / (ms * 1e-3)
    flops_step = (n_layers / N_LAYERS) * alg_flops_per_token() * SEQ  # per GPU
    achieved = flops_step / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))

    # exchange: one reduction of every covariance to its owner (once per calibration, not per step)
    exchange_ms = None
    if world > 1:
        for attempt in range(2):  # the first pass warms NCCL channels and the staging allocations
            barrier()
            e0.record()
            for li, layer in enumerate(accs):
                for ti, acc in enumerate(layer):
                    parallel.reduce_accumulator_to(acc, parallel.owner_of(li * len(LAYER_DIMS) + ti, world),
                                                   parallel.default_group(), total_steps=acc.steps)
            e1.record()
            barrier()
            exchange_ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev)
    del accs
    torch.cuda.empty_cache()

    # ---------------- extra: comparators and the other BASELINE metrics --------------------------
    extra = {}
    if not args.no_extra and rank == 0:
        for key, fn in (("eigh", lambda: extra_eigh(dev, g, world)),
                        ("reference_on_cuda_syrk", lambda: extra_reference_gpu_syrk(dev, acts)),
                        ("lowrank_forward", lambda: extra_lowrank(dev, g, peaks))):
            try:  # the headline number must survive an extra failing
                extra[key] = fn()
            except Exception as exc:
                extra[key + "_error"] = repr(exc)[:300]
            torch.cuda.empty_cache()
    torch.cuda.empty_cache()

    # ---------------- e2e: public API path with host token ids ----------------------------------
    e2e = None
    forward_only_ms = None
    if not args.no_e2e:
        import ptdeco_b200.dwain.decomposition as D
        from synth import models

        with torch.device(dev):
            model = models.LlamaLikeDecoder(layers=n_layers, init=False).to(torch.bfloat16)
        models.fast_init_(model, 271828)
        model.eval()
        names = D._get_decomposeable_submodule_names(model, ["lm_head"])
        gen = torch.Generator().manual_seed(1314159 + rank)
        host_tokens = [torch.randint(0, 128256, (1, SEQ), generator=gen).pin_memory()
                       for _ in range(args.warmup + args.steps)]
        check = torch.zeros(1, dtype=torch.float32).pin_memory()

        def fwd(ids_host):
            ids = ids_host.to(dev, non_blocking=True)
            with torch.no_grad():
                out = model({"input_ids": ids})
            return out

        for i in range(args.warmup):  # plain forward (the user's model alone), for reference
            fwd(host_tokens[i])
        barrier()
        e0.record()
        for i in range(args.steps):
            fwd(host_tokens[args.warmup + i])
        e1.record()
        barrier()
        forward_only_ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps

        originals = D._install_covariance_modules(model, names, True, reduction_factor=0.5)
        last = model.get_submodule(names[-1])
        fwd(host_tokens[0])  # probe forward: decides which targets share an accumulator
        last.units.finish_probe()

        def e2e_step(ids_host):
            fwd(ids_host)
            check.copy_(last.acc.C[0, :1], non_blocking=True)  # D2H read of the step's result
            torch.cuda.current_stream().synchronize()

        cov_units = last.units.units if last.units is not None else []

        def flush_units():
            for u_ in cov_units:
                u_.acc.flush()

        for i in range(args.warmup):
            e2e_step(host_tokens[i])
        flush_units()
        syrk0 = sum(u_.acc.launches for u_ in cov_units)
        barrier()
        e0.record()
        for i in range(args.steps):
            e2e_step(host_tokens[args.warmup + i])
        flush_units()  # staged (deferred) rows are folded in inside the timed region
        e1.record()
        barrier()
        e2e_syrk_launches = sum(u_.acc.launches for u_ in cov_units) - syrk0
        e2e_ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        e2e = {"value": tokens_per_step / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": SEQ * 8, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
               "model_forward_only_ms": forward_only_ms,
               "covariance_accumulators": len(cov_units), "syrk_launches": e2e_syrk_launches,
               "path": "ptdeco_b200.dwain covariance-computing modules in a Llama-3-8B-shape model "
                       "(q/k/v and gate/up share one input-side accumulator each, o / down output-side)"}
        D._restore_modules(model, originals)
        del model, originals
        torch.cuda.empty_cache()

    # ---------------- whole-model wall times (metric i) and the strong-scaling leg ----------------
    if not args.no_extra and not args.no_model_runs:  # every rank: the runs are sharded over the group
        try:
            extra["decompose_wall_time"] = extra_model_runs(dev, with_dwain8b=not args.no_dwain8b, world=world)
        except Exception as exc:
            extra["decompose_wall_time_error"] = repr(exc)[:300]
        torch.cuda.empty_cache()
    strong = None
    if not args.no_strong:
        try:
            strong = strong_scaling_leg(dev, rank, world, n_layers, args.strong_sequences)
        except Exception as exc:
            strong = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

    # ---------------- CPU baseline (rank 0, N = 1) ----------------------------------------------
    cpu = None
    if world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_sample(cores)
        v, sample = cpu_sample(cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE d=14336, N=16384 launch
                         # (the dominant shape: 64 of 224 launches, 88 % of the step), from the
                         # `ncu --set full` capture summarised in
                         # profiles/r01_syrk_pair_ncu_full_summary.json (gate_proj row: 5.70 GB read +
                         # 0.41 GB written). Algorithmic bytes of that launch: 0.470 GB of tokens +
                         # 0.822 GB accumulator RMW; the excess is operand re-reads between waves (the
                         # ~145 MB working set of one wave of 74 tile pairs exceeds the L2). The kernel
                         # is tensor / power bound: this traffic is ~2.5 TB/s, 38 % of the HBM peak.
                         "traffic": 6.11e9, "traffic_unit": "B/launch (d=14336, N=16384)",
                         "traffic_source": "profiles/r01_syrk_pair_ncu_full_summary.json",
                         "algorithmic_bytes_per_launch": 1.292e9,
                         "kernel": "gemm_tc2_kernel<MN,MN> (SYRK, lower triangle, 256x256 tiles on CTA pairs, cta_group::2)",
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src})",
                         "algorithmic_flop_per_token": alg_flops_per_token()},
            "gpu_launches": gpu_launches,
            "e2e": e2e, "cpu_baseline": cpu, "extra": extra,
        }
        if exchange_ms is not None:
            line["exchange_ms_all_covariances"] = exchange_ms
        if strong is not None:
            line["strong_scaling_e2e"] = strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
