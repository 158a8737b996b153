"""Pins the oracle (oracle/, numpy + torch-CPU restatement) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py). CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import drivers as OD
from oracle import primitives as P
from synth import cases

torch.set_float32_matmul_precision("highest")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("method", ["falor", "dwain"])
@pytest.mark.parametrize("kind", ["linear", "conv"])
def test_primitive_covariance_and_reconstruction(golden_dir, method, kind):
    """The reference's own primitive tests (tests/test_deco_primitives_*.py): covariance handed
    to eigh, and full-rank reconstruction <= 1e-6."""
    g = _load(golden_dir, f"prim_{method}_{kind}.npz")
    net, stream = cases.primitive_case(kind, dict_input=(method == "dwain"))
    x = next(stream)
    with torch.no_grad():
        y0 = net(x)
    np.testing.assert_allclose(y0[:2].numpy(), g["y0_head"], rtol=0, atol=1e-6)
    w = (net.mod.weight.detach()[..., 0, 0] if kind == "conv" else net.mod.weight.detach()).numpy()
    np.testing.assert_array_equal(w, g["weight"])
    tap = OD._Tap(net.mod)
    net.mod = tap
    it = stream
    with torch.no_grad():
        if method == "falor":
            u, cov = OD.falor_eigenvectors(net, "mod", it, w, 8, True, False, True, return_cov=True)
        else:
            u, cov = OD.dwain_eigenvectors(net, "mod", it, w, 8, True, return_cov=True)
    rel = np.linalg.norm(cov - g["cov"]) / np.linalg.norm(g["cov"])
    assert rel < 1e-6, rel
    # eigenvector parity up to sign: |u^T u_ref| = I where gaps exist; use projector of the top 8
    cosine = P.min_principal_cosine(P.top_k(u, 8), P.top_k(g["u"].astype(np.float64), 8))
    assert cosine > 0.999999, cosine
    uk = P.top_k(u, 32).astype(np.float32)
    U, V, _ = P.factors(w, uk)
    new = OD.build_two_factor(tap.inner, torch.from_numpy(U.T.copy()), torch.from_numpy(V.T.copy()))
    net.mod = new
    with torch.no_grad():
        y1 = net(x)
    assert (y0 - y1).abs().max().item() < 1e-6


def test_metrics(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    assert abs(P.nsr(g["falor_logits_x"], g["falor_logits_y"], (0,)) - g["falor_logits_nsr"]) < 1e-6 * g["falor_logits_nsr"] + 1e-9
    assert abs(P.nsr(g["dwain_logits_x"], g["dwain_logits_y"], (0, 1)) - g["dwain_logits_nsr"]) < 1e-5 * g["dwain_logits_nsr"]
    assert abs(P.kl_loss(g["falor_logits_x"], g["falor_logits_y"]) - g["falor_logits_kl"]) < 1e-5 * g["falor_logits_kl"]
    np.testing.assert_allclose(P.kl_divergence(g["falor_logits_x"], g["falor_logits_y"]),
                               g["falor_logits_kl_rows"], rtol=2e-4, atol=1e-7)


def test_covariance_and_eigen_d192(golden_dir):
    g = _load(golden_dir, "cov_eig_d192.npz")
    d, n, steps = 192, 1024, 4
    Eyyt = np.zeros((d, d), np.float32)
    for i in range(steps):
        P.update_Eyyt_in_place(Eyyt, cases.step_spectrum_batch(n, d, i).numpy())
    cov = Eyyt / steps
    assert np.linalg.norm(cov - g["cov"]) / np.linalg.norm(g["cov"]) < 1e-6
    u = P.dwain_get_eigenvectors(cov)  # damps in place
    assert np.linalg.norm(cov - g["damped"]) / np.linalg.norm(g["damped"]) < 1e-6
    ev = np.linalg.eigvalsh(cov.astype(np.float64))
    assert np.abs(ev - g["evals"]).max() / g["evals"].max() < 1e-5
    for k in (d // 8, d // 4, d // 2):
        assert P.min_principal_cosine(P.top_k(u, k), P.top_k(g["u"], k)) > 0.9999


def _same_structure(cfg, gold, rel=2e-3):
    assert list(cfg.keys()) == list(gold.keys())
    for name in cfg:
        a, b = json.loads(json.dumps(cfg[name])), gold[name]
        meta_a, meta_b = a.pop("__meta__"), b.pop("__meta__") if "__meta__" in b else {}
        assert a == b, name
        assert meta_a.keys() == meta_b.keys()
        for k in meta_a:
            exact = k in ("proportion", "drop_in_params")
            assert meta_a[k] == pytest.approx(meta_b[k], rel=0 if exact else rel, abs=0 if exact else 1e-7), (name, k)


@pytest.mark.parametrize("name", ["mlp", "convmlp", "deit_small"])
def test_falor_driver_matches_reference(golden_dir, name):
    gold = json.load(open(os.path.join(golden_dir, f"falor_{name}.json")))
    model, stream, kw = cases.falor_case(name)
    trace = []
    cfg = OD.falor_decompose_in_place(module=model, data_iterator=stream, trace=trace, **kw)
    assert stream.position == gold["stream_position"]
    assert [(t["name"], t["rank"]) for t in trace] == [(t["name"], t["rank"]) for t in gold["trace"]]
    for t, g in zip(trace, gold["trace"]):
        assert t["nsr"] == pytest.approx(g["nsr"], rel=2e-3, abs=1e-7)
        assert t["kl"] == pytest.approx(g["kl"], rel=5e-3, abs=1e-7)
    _same_structure(cfg, gold["decompose_config"])


@pytest.mark.parametrize("name", list(cases.DWAIN_CASES))
def test_dwain_driver_matches_reference(golden_dir, name):
    gold = json.load(open(os.path.join(golden_dir, f"dwain_{name}.json")))
    model, stream, mstream, kw = cases.dwain_case(name)
    trace = []
    cfg = OD.dwain_decompose_in_place(
        module=model, data_iterator=stream, metric_iterator=mstream,
        loss_fn=cases.dwain_loss_fn(name), finetune_fn=lambda m, dev, names: m, trace=trace, **kw)
    assert stream.position == gold["stream_position"]
    assert mstream.position == gold["metric_stream_position"]
    assert [(t["name"], t["rank"]) for t in trace] == [(t["name"], t["rank"]) for t in gold["trace"]]
    # bf16 model: the numpy restatement rounds to bf16 at the same places as torch (D:152, D:208,
    # D:423-429) but sums in a different order, so individual bf16 roundings flip; the reference's
    # own fp32 and bf16 runs of this model differ by up to 7 % in NSR. Decisions must be identical.
    rel = 0.15 if "bf16" in name else 5e-3
    for t, g in zip(trace, gold["trace"]):
        assert t["nsr"] == pytest.approx(g["nsr"], rel=rel, abs=1e-7)
    thr = kw["nsr_final_threshold"]
    assert [t["nsr"] < thr for t in trace] == [g["nsr"] < thr for g in gold["trace"]]
    _same_structure(cfg, gold["decompose_config"], rel=0.15 if "bf16" in name else 2e-3)
