"""ORACLE (test infrastructure, NOT the product): CPU restatement of the falor / dwain drivers.

Restates the control flow of the reference's `decompose_in_place` (F:424-511, D:677-800) and
`_process_module` (F:284-399, D:333-537) on top of oracle.primitives (numpy) with the user model
run by torch on CPU. It reproduces the observable behaviour the drop-in must match: iterator
consumption order, bisection / geometric rank search, the falor "last tried rank" quirk
(F:346-348 vs F:379-386), reversed layer order and smallest-accepted-wins in dwain, and the
decompose_config layout (U/m:21-61). Pinned against the real reference by
tests/golden/make_golden.py + tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Any, Callable, Iterator, Optional

import numpy as np
import torch

from . import primitives as P


def _is_target(m: torch.nn.Module) -> bool:
    """F:402-408 / D:540-546."""
    return isinstance(m, torch.nn.Linear) or (
        isinstance(m, torch.nn.Conv2d) and tuple(m.kernel_size) == (1, 1) and m.groups == 1)


def _parent_and_key(root: torch.nn.Module, name: str):
    head, _, key = name.rpartition(".")
    return root.get_submodule(head), key


class _Tap(torch.nn.Module):
    """Stands in for the target layer and remembers its last input (F:51-74, F:98-126)."""

    def __init__(self, inner: torch.nn.Module):
        super().__init__()
        self.inner = inner
        self.last = None

    def forward(self, x):
        self.last = x
        return self.inner(x)

    def rows(self) -> np.ndarray:
        x = self.last
        if isinstance(self.inner, torch.nn.Conv2d):
            x = x.permute(0, 2, 3, 1)
        return x.reshape(-1, self.weight2d().shape[1]).detach().float().numpy()  # bf16 -> exact fp32

    def weight2d(self) -> torch.Tensor:
        w = self.inner.weight.detach()
        return w[..., 0, 0] if w.dim() == 4 else w

    def put_weight(self, w2d: torch.Tensor) -> None:
        with torch.no_grad():
            self.inner.weight.copy_(w2d[:, :, None, None] if self.inner.weight.dim() == 4 else w2d)


def build_two_factor(orig: torch.nn.Module, w1: torch.Tensor, w2: torch.Tensor) -> torch.nn.Sequential:
    """F:76-95, F:128-153: Sequential(first(in->k, no bias), second(k->out, bias of the original))."""
    k = w1.shape[0]
    has_bias = orig.bias is not None
    if isinstance(orig, torch.nn.Conv2d):
        a = torch.nn.Conv2d(orig.in_channels, k, kernel_size=1, bias=False)
        b = torch.nn.Conv2d(k, orig.out_channels, kernel_size=1, bias=has_bias)
        with torch.no_grad():
            a.weight.copy_(w1[:, :, None, None])
            b.weight.copy_(w2[:, :, None, None])
    else:
        a = torch.nn.Linear(orig.in_features, k, bias=False)
        b = torch.nn.Linear(k, orig.out_features, bias=has_bias)
        a.weight.data = w1
        b.weight.data = w2
    if has_bias:
        with torch.no_grad():
            b.bias.copy_(orig.bias)
    return torch.nn.Sequential(a, b)


def module_config(m: torch.nn.Module) -> dict[str, Any]:
    """U/m:21-61."""
    if isinstance(m, torch.nn.Sequential):
        return {"type": "Sequential", "modules": {k: module_config(v) for k, v in m.named_children()}}
    if isinstance(m, torch.nn.Conv2d):
        return {"type": "Conv2d", "in_channels": m.in_channels, "out_channels": m.out_channels,
                "kernel_size": m.kernel_size, "bias": m.bias is not None, "groups": m.groups,
                "padding": m.padding, "padding_mode": m.padding_mode, "stride": m.stride,
                "dilation": m.dilation}
    if isinstance(m, torch.nn.Linear):
        return {"type": "Linear", "in_features": m.in_features, "out_features": m.out_features,
                "bias": m.bias is not None}
    raise ValueError(f"no config for {type(m)}")


# ------------------------------------------------------------------------------------ falor
def falor_eigenvectors(root, name: str, it: Iterator[torch.Tensor], weight: np.ndarray,
                       num_data_steps: int, use_float64: bool, use_mean: bool,
                       use_damping: bool, return_cov: bool = False):
    """F:165-208 with the target already tapped."""
    root.eval()
    tap = root.get_submodule(name)
    acc = np.float64 if use_float64 else np.float32
    n_out = weight.shape[0]
    Ey = np.zeros(n_out, dtype=acc)
    Eyyt = np.zeros((n_out, n_out), dtype=acc)
    for _ in range(num_data_steps):
        root(next(it))
        P.accumulate_Ey_and_Eyyt(Ey, Eyyt, weight, tap.rows())
    cov = P.falor_covariance(Ey, Eyyt, num_data_steps, use_mean, use_damping)
    u = P.eigenvectors_ascending(cov)
    return (u, cov.copy()) if return_cov else u


def falor_process(root, name, it, nsr_thr, kl_thr, num_data_steps, num_metric_steps, use_float64,
                  use_mean, use_damping, trace: Optional[list] = None) -> dict[str, Any]:
    """F:284-399."""
    parent, key = _parent_and_key(root, name)
    orig = getattr(parent, key)
    tap = _Tap(orig)
    setattr(parent, key, tap)
    w_t = tap.weight2d().clone()
    w = w_t.numpy()
    d_out, d_in = w.shape
    full_rank = min(d_in, d_out)
    if full_rank == 1:
        setattr(parent, key, orig)
        return {"proportion": 1.0, "nsr_final": 0.0, "kl_final": 0.0, "decomposed_module": None}
    u = falor_eigenvectors(root, name, it, w, num_data_steps, use_float64, use_mean, use_damping)
    rank_best, width = full_rank, full_rank // 2
    nsr_new = kl_new = 0.0
    U = V = None
    while width > 0:
        rank_new = rank_best - width
        uk = P.top_k(u, rank_new).astype(np.float32)
        U, V, deco = P.factors(w, uk)
        nsr_new = kl_new = 0.0
        for _ in range(num_metric_steps):
            x = next(it)
            tap.put_weight(torch.from_numpy(np.ascontiguousarray(deco)))
            y_deco = root(x)
            tap.put_weight(w_t)
            y_orig = root(x)
            # .mean() of a scalar is the scalar (F:228-231); metrics evaluated in the model dtype
            nsr_new += float(_torch_nsr(y_deco, y_orig, (0,)))
            kl_new += float(_torch_kl_loss(y_deco, y_orig))
        nsr_new /= num_metric_steps
        kl_new /= num_metric_steps
        accepted = nsr_new < nsr_thr and kl_new < kl_thr
        if trace is not None:
            trace.append({"name": name, "rank": rank_new, "nsr": nsr_new, "kl": kl_new,
                          "accepted": accepted})
        if accepted:
            rank_best = rank_new
        width //= 2
    tap.put_weight(w_t)
    proportion = rank_best / full_rank
    new = None
    if rank_best != full_rank and P.is_num_params_reduced(proportion, d_in, d_out):
        # built from the LAST tried rank's factors (quirk)
        new = build_two_factor(orig, torch.from_numpy(np.ascontiguousarray(U)).T,
                               torch.from_numpy(np.ascontiguousarray(V)).T)
    setattr(parent, key, orig)
    return {"proportion": proportion, "nsr_final": nsr_new, "kl_final": kl_new,
            "decomposed_module": new}


def falor_decompose_in_place(*, module, data_iterator, blacklisted_module_names=None,
                             proportion_threshold, nsr_final_threshold, kl_final_threshold,
                             num_data_steps, num_metric_steps, use_float64, use_mean, use_damping,
                             trace: Optional[list] = None) -> dict[str, Any]:
    """F:424-511 (CPU)."""
    black = blacklisted_module_names or []
    names = [n for n, m in module.named_modules() if _is_target(m)]
    results = {}
    with torch.no_grad():
        for n in names:
            if n in black:
                continue
            results[n] = falor_process(module, n, data_iterator, nsr_final_threshold,
                                       kl_final_threshold, num_data_steps, num_metric_steps,
                                       use_float64, use_mean, use_damping, trace)
    cfg = {}
    for n in names:
        if n in black:
            continue
        r = results[n]
        new = r["decomposed_module"]
        if new is None or not (r["proportion"] < proportion_threshold):
            continue
        parent, key = _parent_and_key(module, n)
        setattr(parent, key, new)
        c = module_config(new)
        c["__meta__"] = {k: v for k, v in r.items() if k != "decomposed_module"}
        cfg[n] = c
    return cfg


# ------------------------------------------------------------------------------------ dwain
def dwain_eigenvectors(root, name: str, it, weight: np.ndarray, num_data_steps: int,
                       decompose_in_float64: bool, return_cov: bool = False, bf16: bool = False):
    """D:211-244 with the target already tapped. bf16: the model's dtype is bfloat16, so
    y = x W^T is a bf16 GEMM and every per-step product is rounded to bf16 (D:152, D:239-240)."""
    root.eval()
    tap = root.get_submodule(name)
    acc = np.float64 if decompose_in_float64 else np.float32
    Eyyt = np.zeros((weight.shape[0], weight.shape[0]), dtype=acc)
    for _ in range(num_data_steps):
        root(next(it))
        if bf16:
            P.update_Eyyt_in_place_bf16(Eyyt, P.bf16_round(tap.rows() @ weight.T))
            continue
        P.update_Eyyt_in_place(Eyyt, tap.rows() @ weight.T)
    cov = Eyyt / num_data_steps
    u = P.dwain_get_eigenvectors(cov)
    return (u, cov.copy()) if return_cov else u


class _CovLinear(torch.nn.Module):
    """D:166-208: layer forward that also folds its (bias-free) output into Eyyt."""

    def __init__(self, lin: torch.nn.Linear, f64: bool):
        super().__init__()
        self.lin = lin
        self.Eyyt = np.zeros((lin.out_features, lin.out_features), np.float64 if f64 else np.float32)
        self.steps = 0

    def forward(self, x):
        y = x @ self.lin.weight.T
        if y.dtype == torch.bfloat16:
            P.update_Eyyt_in_place_bf16(self.Eyyt, y.reshape(-1, self.lin.out_features).detach().float().numpy())
        else:
            P.update_Eyyt_in_place(self.Eyyt, y.reshape(-1, self.lin.out_features).detach().numpy())
        if self.lin.bias is not None:
            y = y + self.lin.bias
        self.steps += 1
        return y


def dwain_precompute(module, names, num_splits, num_data_steps, it, f64) -> dict[str, np.ndarray]:
    """D:580-674: chunked one-pass covariance for every target of the chunk."""
    out: dict[str, np.ndarray] = {}
    chunk = len(names) // num_splits
    if chunk == 0:
        chunk, num_splits = 1, len(names)
    parts = num_splits if len(names) % num_splits == 0 else num_splits + 1
    for p in range(parts):
        sub = names[p * chunk:(p + 1) * chunk]
        saved = {}
        for n in sub:
            parent, key = _parent_and_key(module, n)
            saved[n] = getattr(parent, key)
            setattr(parent, key, _CovLinear(saved[n], f64))
        module.eval()
        with torch.no_grad():
            for _ in range(num_data_steps):
                module(next(it))
        for n in sub:
            c = module.get_submodule(n)
            u = P.dwain_get_eigenvectors(c.Eyyt / c.steps)
            # D:208: cast to the weight dtype (bf16 values kept in a float32 array)
            out[n] = (P.bf16_round(u) if saved[n].weight.dtype == torch.bfloat16
                      else u.astype(_np_dtype(saved[n].weight.dtype)))
        for n in sub:
            parent, key = _parent_and_key(module, n)
            setattr(parent, key, saved[n])
    return out


def _np_dtype(t: torch.dtype):
    return {torch.float32: np.float32, torch.float64: np.float64}[t]


def dwain_process(root, name, it, loss_fn, nsr_thr, num_data_steps, num_metric_steps, metric_it,
                  num_params, min_rank, trade_off, reduction, max_ppl_diff, f64, u_matrix,
                  trace: Optional[list] = None) -> dict[str, Any]:
    """D:333-537."""
    parent, key = _parent_and_key(root, name)
    orig = getattr(parent, key)
    tap = _Tap(orig)
    setattr(parent, key, tap)
    w_t = tap.weight2d().clone()
    bf16 = w_t.dtype == torch.bfloat16
    w = w_t.float().numpy()
    d_out, d_in = w.shape
    full_rank = min(d_in, d_out)
    if full_rank == 1:
        setattr(parent, key, orig)
        return {"proportion": 1.0, "nsr_final": 0.0, "ppl_final": 0.0, "decomposed_module": None}
    if u_matrix is None:
        u_matrix = dwain_eigenvectors(root, name, it, w, num_data_steps, f64, bf16=bf16)

    def factors(rank):
        if bf16:
            return P.factors_bf16(w, P.top_k(u_matrix, rank))
        return P.factors(w, P.top_k(u_matrix, rank).astype(w.dtype))

    def as_weight(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(w_t.dtype)
    rank_best = rank_new = full_rank
    nsr_best = ppl_best = 0.0
    tried = False
    while rank_new > min_rank:
        rank_new = int(rank_new * reduction)
        drop = P.get_params_for_proportion(1.0, d_in, d_out) - P.get_params_for_proportion(
            rank_new / full_rank, d_in, d_out)
        ppl_thr = drop / num_params * trade_off
        if drop == 0:
            continue
        _, _, deco = factors(rank_new)
        tried = True
        nsr_new = ppl_new = diff_new = 0.0
        for _ in range(num_metric_steps):
            d = next(metric_it)
            tap.put_weight(as_weight(deco))
            y_deco = root(d)
            tap.put_weight(w_t)
            y_orig = root(d)
            ppl_d = torch.exp(loss_fn(d, y_deco)).mean()
            ppl_o = torch.exp(loss_fn(d, y_orig)).mean()
            diff_new += float((ppl_d - ppl_o) / ppl_o)
            nsr_new += float(_torch_nsr(y_deco, y_orig, (0, 1)))
            ppl_new += float(ppl_d)
        nsr_new /= num_metric_steps
        ppl_new /= num_metric_steps
        diff_new /= num_metric_steps
        accepted = not (diff_new >= ppl_thr or diff_new >= max_ppl_diff or nsr_new >= nsr_thr)
        if trace is not None:
            trace.append({"name": name, "rank": rank_new, "nsr": nsr_new, "ppl_diff": diff_new,
                          "ppl_thr": ppl_thr, "accepted": accepted})
        if accepted:
            rank_best, nsr_best, ppl_best = rank_new, nsr_new, ppl_new
    proportion = rank_best / full_rank
    if tried and rank_best != full_rank and P.is_num_params_reduced(proportion, d_in, d_out):
        U, V, _ = factors(rank_best)
        new = build_two_factor(orig, as_weight(U).T, as_weight(V).T)
        drop = P.get_params_for_proportion(1.0, d_in, d_out) - P.get_params_for_proportion(
            proportion, d_in, d_out)
        # on success the tap is overwritten by the caller's swap (D:779)
        return {"proportion": proportion, "nsr_final": nsr_best, "ppl_final": ppl_best,
                "drop_in_params": drop, "decomposed_module": new}
    setattr(parent, key, orig)
    return {"proportion": 1.0, "nsr_final": 0.0, "ppl_final": 0.0, "drop_in_params": 0,
            "decomposed_module": None}


def dwain_decompose_in_place(*, module, data_iterator, loss_fn, num_data_steps, metric_iterator,
                             num_metric_steps, blacklisted_module_names=None, nsr_final_threshold,
                             finetune_fn: Callable, min_rank=32, trade_off_factor=0.5,
                             reduction_factor=0.5, max_accepted_ppl_diff=0.1,
                             decompose_in_float64=True, precomputing_covariance_num_splits=None,
                             trace: Optional[list] = None) -> dict[str, Any]:
    """D:677-800 (CPU)."""
    params = {p.data_ptr(): p for p in module.parameters()}
    num_params = sum(p.numel() for p in params.values())
    black = blacklisted_module_names or []
    names = [n for n, m in module.named_modules() if _is_target(m) and n not in black]
    u_dict = {}
    if precomputing_covariance_num_splits is not None and precomputing_covariance_num_splits > 0:
        u_dict = dwain_precompute(module, names, precomputing_covariance_num_splits, num_data_steps,
                                  data_iterator, decompose_in_float64)
    cfg = {}
    done = []
    for n in reversed(names):
        with torch.no_grad():
            r = dwain_process(module, n, data_iterator, loss_fn, nsr_final_threshold,
                              num_data_steps, num_metric_steps, metric_iterator, num_params,
                              min_rank, trade_off_factor, reduction_factor, max_accepted_ppl_diff,
                              decompose_in_float64, u_dict.pop(n) if len(u_dict) > 0 else None,
                              trace)
        new = r["decomposed_module"]
        if new is not None:
            done.append(n)
            parent, key = _parent_and_key(module, n)
            setattr(parent, key, new)
            module = finetune_fn(module, torch.device("cpu"), done)
            c = module_config(new)
            c["__meta__"] = {k: v for k, v in r.items() if k != "decomposed_module"}
            cfg[n] = c
    return cfg


# ------------------------------------------------------------------ metrics on model outputs
def _torch_nsr(x: torch.Tensor, y: torch.Tensor, dims) -> torch.Tensor:
    """U/l:10-22, evaluated in the model dtype like the reference."""
    var = torch.square(torch.std(y, dim=dims))
    msd = torch.square(x - y).mean(dim=dims)
    return torch.divide(msd, var + 1e-3).mean()


def _torch_kl(q_logits, p_logits):
    q = torch.softmax(q_logits, dim=-1)
    p = torch.softmax(p_logits, dim=-1)
    return (p * torch.log(p / q)).sum(dim=1)


def _torch_kl_loss(s, t):
    """U/l:57-63."""
    return torch.maximum(_torch_kl(s, t), _torch_kl(t, s)).mean()
