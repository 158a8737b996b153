"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(TCLResearchEurope/ptdeco, imported from /root/reference/src) on CPU in the build container.

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py prim cov   # a subset

The reference cannot travel to the GPU box, so its outputs are committed here as small fixtures:
  prim_{falor,dwain}_{linear,conv}.npz  the reference's own primitive tests
                                         (tests/test_deco_primitives_{falor,dwain}.py) with the
                                         covariance handed to torch.linalg.eigh captured
  metrics.npz                            NSR / KL values on seeded logits (U/l:10-63)
  cov_eig_d192.npz                       _update_Eyyt_in_place + _get_eigenvectors on step-spectrum
                                         activations (D:147-163)
  eig_ref_d768.npz, eig_ref_d2048.npz    the same at d = 768 / 2048: eigenvalues + top-k eigenvector
                                         blocks (fp16) of the reference's torch.linalg.eigh
  falor_*.json / dwain_*.json            decompose_config + per-trial trace of decompose_in_place
Everything is seeded; re-running reproduces the files bit for bit on the same torch build
(torch 2.11.0+cu128, CPU/MKL).
"""
from __future__ import annotations

import json
import logging
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import ptdeco  # noqa: E402  (the reference)
import ptdeco.dwain.decomposition as RD  # noqa: E402
import ptdeco.falor  # noqa: E402
import ptdeco.falor.decomposition as RF  # noqa: E402

from synth import cases  # noqa: E402

torch.set_float32_matmul_precision("highest")
torch.set_num_threads(max(1, os.cpu_count() or 1))


class EighTap:
    """Records the matrix handed to torch.linalg.eigh (the covariance is not otherwise observable)."""

    def __init__(self):
        self.inputs = []
        self._orig = torch.linalg.eigh

    def __enter__(self):
        def tapped(a, *args, **kw):
            self.inputs.append(a.detach().clone())
            return self._orig(a, *args, **kw)

        torch.linalg.eigh = tapped
        return self

    def __exit__(self, *exc):
        torch.linalg.eigh = self._orig


def make_prim() -> None:
    for method in ("falor", "dwain"):
        for kind in ("linear", "conv"):
            net, stream = cases.primitive_case(kind, dict_input=(method == "dwain"))
            x = next(stream)
            with torch.no_grad():
                y0 = net(x)
            mod = RF if method == "falor" else RD
            mod._wrap_in_place(net, "mod")
            w = net.mod.get_weight_copy()
            with EighTap() as tap, torch.no_grad():
                if method == "falor":
                    u = RF._compute_decompositon_of_covariance_matrix(
                        root_module=net, decomposed_submodule_name="mod", data_iterator=stream,
                        weight=w, num_data_steps=8, device=torch.device("cpu"), use_float64=True,
                        use_mean=False, use_damping=True)
                else:
                    u = RD._compute_covariance_matrix_decomposition(
                        root_module=net, decomposed_submodule_name="mod", data_iterator=stream,
                        weight=w, num_data_steps=8, device=torch.device("cpu"),
                        decompose_in_float64=True)
                uk = u[:, u.shape[1] - 32:].to(torch.float)
                U, V = w.T @ uk, uk.T
                new = net.get_submodule("mod").get_decomposed_module(u=U.T, v=V.T)
            mod._unwrap_in_place(net, "mod")
            bias = net.mod.bias.detach().clone()
            ptdeco.utils.replace_submodule_in_place(net, "mod", new)
            with torch.no_grad():
                y1 = net(x)
            np.savez_compressed(
                os.path.join(HERE, f"prim_{method}_{kind}.npz"), weight=w.numpy(), bias=bias.numpy(),
                cov=tap.inputs[0].numpy(), u=u.numpy(), y0_head=y0[:2].numpy(),
                max_abs_diff=np.float64((y0 - y1).abs().max().item()))
            print(f"prim {method} {kind}: max|y0-y1| = {(y0 - y1).abs().max().item():.3e}")


def make_metrics() -> None:
    g = torch.Generator().manual_seed(cases.DATA_SEED)
    out = {}
    for name, shape, dims in (("falor_logits", (16, 10), (0,)), ("dwain_logits", (2, 24, 96), (0, 1))):
        y = torch.randn(*shape, generator=g) * 2.0
        x = y + 0.1 * torch.randn(*shape, generator=g)
        out[name + "_x"] = x.numpy()
        out[name + "_y"] = y.numpy()
        out[name + "_nsr"] = np.float64(ptdeco.utils.calc_per_channel_noise_to_signal_ratio(
            x=x, y=y, non_channel_dim=dims).item())
    x, y = torch.from_numpy(out["falor_logits_x"]), torch.from_numpy(out["falor_logits_y"])
    out["falor_logits_kl"] = np.float64(ptdeco.utils.calc_kl_loss(x, y).item())
    out["falor_logits_kl_rows"] = ptdeco.utils.calc_kl_divergence(x, y).numpy()
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("metrics:", {k: float(v) for k, v in out.items() if np.ndim(v) == 0})


def make_cov() -> None:
    d, n, steps = 192, 1024, 4
    Eyyt = torch.zeros((d, d), dtype=torch.float32)
    for i in range(steps):
        RD._update_Eyyt_in_place(Eyyt, cases.step_spectrum_batch(n, d, i))
    cov = Eyyt / steps
    with EighTap() as tap:
        u = RD._get_eigenvectors(cov.clone())
    evals = torch.linalg.eigvalsh(tap.inputs[0])
    np.savez_compressed(os.path.join(HERE, "cov_eig_d192.npz"), cov=cov.numpy(),
                        damped=tap.inputs[0].numpy(), u=u.numpy(), evals=evals.numpy())
    print("cov_eig_d192: lambda max/min", evals.max().item(), evals.min().item())


def make_eig_ref() -> None:
    """Reference-held K3 fixtures at the sizes of VERDICT r1 weak #3: covariance through the
    reference's _update_Eyyt_in_place, eigenvectors through its _get_eigenvectors (D:147-163).
    Only the eigenvalues and the top-k blocks at the designed spectral steps are kept (fp16)."""
    for d, n, steps, kmax in ((768, 1024, 4, 384), (2048, 2048, 4, 512)):
        Eyyt = torch.zeros((d, d), dtype=torch.float32)
        for i in range(steps):
            RD._update_Eyyt_in_place(Eyyt, cases.step_spectrum_batch(n, d, i))
        cov = Eyyt / steps
        with EighTap() as tap:
            u = RD._get_eigenvectors(cov.clone())
        evals = torch.linalg.eigvalsh(tap.inputs[0].double())
        ks = [k for k in (d // 8, d // 4, d // 2) if k <= kmax]
        np.savez_compressed(os.path.join(HERE, f"eig_ref_d{d}.npz"), evals=evals.numpy(),
                            u_top=u[:, d - kmax:].numpy().astype(np.float16), ks=np.array(ks),
                            n=np.int64(n), steps=np.int64(steps))
        print(f"eig_ref_d{d}: lambda max/min {evals.max().item():.4g} {evals.min().item():.4g}, ks {ks}")


class _Trace(logging.Handler):
    """Collects the per-trial log lines of the reference (F:371-373, D:470-486)."""

    def __init__(self):
        super().__init__(level=logging.INFO)
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def _run_with_trace(method: str, fn):
    samples = []
    mod = RF if method == "falor" else RD
    orig = mod._compute_metrics

    def tapped(**kw):
        r = orig(**kw)
        samples.append([float(t) for t in r])
        return r

    handler = _Trace()
    lg = logging.getLogger(mod.__name__)
    lg.addHandler(handler)
    lg.setLevel(logging.INFO)
    mod._compute_metrics = tapped
    try:
        cfg = fn()
    finally:
        mod._compute_metrics = orig
        lg.removeHandler(handler)
    return cfg, samples, handler.lines


def _jsonable(o):
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (tuple, list)):
        return [_jsonable(v) for v in o]
    return o


def make_falor(which=None) -> None:
    for name in cases.FALOR_CASES:
        if which and name not in which:
            continue
        model, stream, kw = cases.falor_case(name)
        steps = kw["num_metric_steps"]
        cfg, samples, lines = _run_with_trace("falor", lambda: ptdeco.falor.decompose_in_place(
            module=model, device=torch.device("cpu"), data_iterator=stream, **kw))
        trials = []
        cur = None
        pat = re.compile(r"Processing (\S+): i=(\d+) rank_width=(\d+) rank_new=(\d+)")
        si = 0
        for ln in lines:
            m = pat.search(ln)
            if m:
                cur = m.group(1)
                grp = samples[si:si + steps]
                si += steps
                trials.append({"name": cur, "rank": int(m.group(4)),
                               "nsr": sum(s[0] for s in grp) / steps,
                               "kl": sum(s[1] for s in grp) / steps})
        assert si == len(samples)
        with open(os.path.join(HERE, f"falor_{name}.json"), "w") as f:
            json.dump({"kwargs": kw, "decompose_config": _jsonable(cfg), "trace": trials,
                       "stream_position": stream.position}, f, indent=1)
        print(f"falor {name}: {len(cfg)} decomposed, {len(trials)} trials, stream at {stream.position}")
        for n_, c in cfg.items():
            print("   ", n_, c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels")),
                  c["__meta__"])


def make_dwain(which=None) -> None:
    for name in cases.DWAIN_CASES:
        if which and name not in which:
            continue
        model, stream, mstream, kw = cases.dwain_case(name)
        steps = kw["num_metric_steps"]
        cfg, samples, lines = _run_with_trace("dwain", lambda: ptdeco.dwain.decompose_in_place(
            module=model, device=torch.device("cpu"), data_iterator=stream, metric_iterator=mstream,
            loss_fn=cases.dwain_loss_fn(name), finetune_fn=lambda m, dev, names: m, **kw))
        trials = []
        pat = re.compile(r"i=(\d+) rank_new=(\d+)/(\d+) nsr_new")
        cur_name = None
        si = 0
        for ln in lines:
            m0 = re.match(r"PROCESSING (\S+) MODULE", ln)
            if m0:
                cur_name = m0.group(1)
            m = pat.search(ln)
            if m:
                grp = samples[si:si + steps]
                si += steps
                trials.append({"name": cur_name, "rank": int(m.group(2)),
                               "nsr": sum(s[0] for s in grp) / steps,
                               "ppl_diff": sum((s[1] - s[2]) / s[2] for s in grp) / steps,
                               "ppl_deco": sum(s[1] for s in grp) / steps})
        assert si == len(samples), (si, len(samples))
        with open(os.path.join(HERE, f"dwain_{name}.json"), "w") as f:
            json.dump({"kwargs": kw, "decompose_config": _jsonable(cfg), "trace": trials,
                       "stream_position": stream.position,
                       "metric_stream_position": mstream.position}, f, indent=1)
        print(f"dwain {name}: {len(cfg)} decomposed, {len(trials)} trials")
        for n_, c in cfg.items():
            print("   ", n_, c["modules"]["0"].get("out_features"), c["__meta__"])


if __name__ == "__main__":
    what = sys.argv[1:] or ["prim", "metrics", "cov", "eig", "falor", "dwain"]
    sel = [w for w in what if w not in ("prim", "metrics", "cov", "eig", "falor", "dwain")]
    if "prim" in what:
        make_prim()
    if "metrics" in what:
        make_metrics()
    if "cov" in what:
        make_cov()
    if "eig" in what:
        make_eig_ref()
    if "falor" in what:
        make_falor(sel or None)
    if "dwain" in what:
        make_dwain(sel or None)
