"""GPU parity tests: the CUDA path (through the C-ABI, via ptdeco_b200.linalg / falor / dwain)
against the oracle and the golden fixtures of the unmodified reference. Tolerances are the
north_star's: covariance <= 1e-5 relative, eigenvalues <= 1e-4 (normwise), top-k principal-angle
cosines >= 0.9999, identical ranks."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import primitives as P
from synth import cases

pytestmark = pytest.mark.gpu

COV_TOL = 1e-5
EVAL_TOL = 1e-4
COS_TOL = 0.9999


@pytest.fixture(scope="module")
def dev():
    # the user models run on torch: keep their GPU forwards at fp32 like the golden CPU runs
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64))
                 / np.linalg.norm(np.asarray(b, np.float64)))


# ------------------------------------------------------------------------------------ K1 / K2
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,d", [(1, 8), (37, 10), (1576, 192), (1000, 200), (3136, 1000), (2048, 1024)])
def test_syrk_matches_oracle(dev, dtype, n, d):
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(n * 7919 + d)
    acc = linalg.CovarianceAccumulator(d, dev, with_mean=True)
    Ey = np.zeros(d, np.float64)
    Eyyt = np.zeros((d, d), np.float64)
    steps = 3
    for _ in range(steps):
        y = (torch.randn(n, d, generator=g) * torch.logspace(0, -2, d)).to(dtype)
        acc.update(y.to(dev))
        yd = y.double().numpy()
        Eyyt += yd.T @ yd / n
        Ey += yd.mean(0)
    assert acc.steps == steps
    cov = acc.finalize(use_mean=False, damp_factor=0.0).cpu().numpy()
    assert _rel(cov, Eyyt / steps) < COV_TOL
    assert np.array_equal(cov, cov.T)
    assert np.abs(acc.colsum.cpu().numpy() / steps - Ey / steps).max() < 1e-5 * (np.abs(Ey).max() / steps + 1e-3)


def test_syrk_bias_subtraction_and_strided_rows(dev):
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(5)
    n, d = 500, 96
    y = torch.randn(n, d + 8, generator=g)
    b = torch.randn(d, generator=g)
    acc = linalg.CovarianceAccumulator(d, dev)
    acc.update(y.to(dev)[:, :d], sub=b.to(dev))  # row pitch d+8
    ref = ((y[:, :d] - b).double().T @ (y[:, :d] - b).double() / n).numpy()
    assert _rel(acc.finalize(False, 0.0).cpu().numpy(), ref) < COV_TOL


@pytest.mark.parametrize("use_mean,use_damping", [(False, True), (True, True), (True, False), (False, False)])
def test_finalize_matches_falor_covariance(dev, use_mean, use_damping):
    """F:192-205 including the damping quirk (a no-op when use_mean=True)."""
    from ptdeco_b200 import linalg
    d, n, steps = 64, 300, 4
    g = torch.Generator().manual_seed(11)
    w = np.eye(d, dtype=np.float32)
    Ey, Eyyt = np.zeros(d, np.float32), np.zeros((d, d), np.float32)
    acc = linalg.CovarianceAccumulator(d, dev, with_mean=True)
    for _ in range(steps):
        y = torch.randn(n, d, generator=g) + 0.5
        acc.update(y.to(dev))
        P.accumulate_Ey_and_Eyyt(Ey, Eyyt, w, y.numpy())
    ref = P.falor_covariance(Ey, Eyyt, steps, use_mean, use_damping)
    damp = P.EIGEN_DAMPEN_FACTOR if (use_damping and not use_mean) else 0.0
    cov = acc.finalize(use_mean=use_mean, damp_factor=damp).cpu().numpy()
    assert _rel(cov, ref) < COV_TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_deferred_updates_match_immediate(dev, dtype):
    """Staging several equal-N batches and folding them in with one SYRK call gives the covariance
    of per-batch calls (sum_steps y^T y / N is a SYRK over the concatenated rows)."""
    from ptdeco_b200 import linalg
    d, n = 200, 96
    g = torch.Generator().manual_seed(9)
    ys = [torch.randn(n, d, generator=g).to(dtype).to(dev) for _ in range(7)]
    ys.append(torch.randn(n + 8, d, generator=g).to(dtype).to(dev))  # a different N forces a flush
    now = linalg.CovarianceAccumulator(d, dev, with_mean=True)
    later = linalg.CovarianceAccumulator(d, dev, with_mean=True, defer_rows=3 * n)
    for y in ys:
        now.update(y)
        later.update(y)
    assert later.launches < now.launches and later.steps == now.steps == 8
    a = now.finalize(True, 0.01).cpu().numpy()
    b = later.finalize(True, 0.01).cpu().numpy()
    assert _rel(b, a) < 2e-6
    ref = sum((y.double().T @ y.double() / y.shape[0]) for y in ys) / 8
    mean = sum(y.double().mean(0) for y in ys) / 8
    ref = (ref - torch.outer(mean, mean)).cpu().numpy()
    ref = ref + 0.01 * np.mean(np.diag(ref)) * np.eye(d)
    assert _rel(b, ref) < COV_TOL


def test_syrk_empty_and_bad_arguments(dev):
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    acc = linalg.CovarianceAccumulator(16, dev)
    with pytest.raises(ValueError):
        acc.update(torch.zeros(0, 16, device=dev))
    with pytest.raises(ValueError):
        acc.update(torch.zeros(4, 17, device=dev))
    with pytest.raises(nat.NativeError):
        acc.update(torch.zeros(4, 16))  # CPU tensor: no CPU path
    with pytest.raises(nat.NativeError):
        linalg.CovarianceAccumulator(16, torch.device("cpu"))


# ------------------------------------------------------------------------------------ K3
def _check_eigh(cov64, ev, u, k, cos_ks=()):
    d = cov64.shape[0]
    ref_ev, ref_u = np.linalg.eigh(cov64)
    assert np.abs(ev - ref_ev).max() / np.abs(ref_ev).max() < EVAL_TOL
    assert np.abs(u.T @ u - np.eye(k)).max() < 2e-5
    resid = np.abs(cov64 @ u - u * ref_ev[d - k:]).max() / np.abs(ref_ev).max()
    assert resid < 2e-5, resid
    for kk in cos_ks:
        assert P.min_principal_cosine(ref_u[:, d - kk:], u[:, k - kk:]) >= COS_TOL


def test_eigh_golden_d192(dev, golden_dir):
    from ptdeco_b200 import linalg
    g = np.load(os.path.join(golden_dir, "cov_eig_d192.npz"))
    ev, u = linalg.eigh(torch.from_numpy(g["damped"]).to(dev))
    ev, u = ev.cpu().numpy().astype(np.float64), u.cpu().numpy().astype(np.float64)
    assert np.abs(ev - g["evals"]).max() / g["evals"].max() < EVAL_TOL
    for k in (24, 48, 96):
        assert P.min_principal_cosine(P.top_k(g["u"], k), P.top_k(u, k)) >= COS_TOL
    _check_eigh(g["damped"].astype(np.float64), ev, u, 192)


@pytest.mark.parametrize("d", [768, 2048])
def test_eigh_matches_reference_fixture(dev, golden_dir, d):
    """K3 against the REFERENCE's output at d = 768 / 2048: tests/golden/make_golden.py ran
    _update_Eyyt_in_place + _get_eigenvectors (D:147-163, torch.linalg.eigh on CPU) on seeded
    step-spectrum activations and committed the eigenvalues and the top-k eigenvector blocks
    (fp16 to keep the fixture small). The covariance is regenerated here from the same seeds."""
    from ptdeco_b200 import linalg
    g = np.load(os.path.join(golden_dir, f"eig_ref_d{d}.npz"))
    steps, n = int(g["steps"]), int(g["n"])
    acc = linalg.CovarianceAccumulator(d, dev)
    for i in range(steps):
        acc.update(cases.step_spectrum_batch(n, d, i).to(dev))
    cov = acc.finalize(use_mean=False, damp_factor=P.EIGEN_DAMPEN_FACTOR)
    kmax = int(g["u_top"].shape[1])
    ev, u = linalg.eigh(cov, k=kmax)
    ev = ev.double().cpu().numpy()
    assert np.abs(ev - g["evals"]).max() / g["evals"].max() < EVAL_TOL
    u = u.double().cpu().numpy()
    u_ref = g["u_top"].astype(np.float64)
    for k in [int(x) for x in g["ks"]]:
        # the fp16 block is re-orthonormalised: its rounding (5e-4 per entry) tilts the SPAN by
        # ~1e-7 only, but would enter the cosines at first order through the lost orthonormality
        q, _ = np.linalg.qr(u_ref[:, kmax - k:])
        assert P.min_principal_cosine(q, u[:, kmax - k:]) >= COS_TOL, k


@pytest.mark.parametrize("d,k", [(1, 1), (2, 2), (3, 2), (10, 10), (32, 32), (33, 7), (96, 96), (97, 97),
                                 (128, 128), (130, 65), (200, 50), (576, 288), (768, 384), (1000, 125)])
def test_eigh_step_spectrum(dev, d, k):
    from ptdeco_b200 import linalg
    y = cases.step_spectrum_batch(4 * d + 8, d, 3).double()
    cov = (y.T @ y / y.shape[0])
    cov = cov + 0.01 * cov.diagonal().mean() * torch.eye(d, dtype=torch.float64)
    c32 = cov.float()
    ev, u = linalg.eigh(c32.to(dev), k=k)
    assert tuple(u.shape) == (d, k) and tuple(ev.shape) == (d,)
    cos_ks = [kk for kk in (d // 8, d // 4, d // 2) if 1 <= kk <= k and d >= 32]
    _check_eigh(c32.double().numpy(), ev.cpu().numpy().astype(np.float64),
                u.cpu().numpy().astype(np.float64), k, cos_ks)


@pytest.mark.parametrize("d,k", [(1000, 250), (2050, 1025)])
def test_eigh_lower_triangle_symv_path(dev, d, k):
    """The panel kernel's lower-triangle symv (default for trailing sizes >= 5120, where the
    one-stage reduction is HBM-bound) forced on at test sizes: same eigenvalue / subspace /
    residual bars as the full-row path, and agreement of the two paths' eigenvalues."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    y = cases.step_spectrum_batch(4 * d + 8, d, 5).double()
    cov = (y.T @ y / y.shape[0])
    cov = cov + 0.01 * cov.diagonal().mean() * torch.eye(d, dtype=torch.float64)
    c32 = cov.float()
    L = nat.lib()
    try:
        L.ptdeco_debug_set(102, 0)  # blocked panel kernel all the way (these sizes are otherwise resident)
        L.ptdeco_debug_set(101, 128)
        ev, u = linalg.eigh(c32.to(dev), k=k)
        L.ptdeco_debug_set(101, 0)
        ev_full, _ = linalg.eigh(c32.to(dev), k=k)
    finally:
        L.ptdeco_debug_set(101, 5120)
        L.ptdeco_debug_set(102, 1)
    cos_ks = [kk for kk in (d // 8, d // 4) if kk <= k]
    _check_eigh(c32.double().numpy(), ev.cpu().numpy().astype(np.float64),
                u.cpu().numpy().astype(np.float64), k, cos_ks)
    assert float((ev - ev_full).abs().max() / ev_full.abs().max()) < 2e-6


@pytest.mark.parametrize("d,k,rows", [(64, 64, 4), (130, 130, 2), (500, 125, 4), (1000, 1000, 8),
                                      (2050, 300, 4), (2700, 340, 4)])
def test_eigh_resident_matches_blocked(dev, d, k, rows):
    """The shared-memory-resident tridiagonalisation (one packet exchange per Householder column;
    whole matrix for d <= ~2560, the tail of the reduction beyond) against the blocked panel
    kernel: same bars, eigenvalues of the two paths equal to fp32 rounding. d = 2700 runs blocked
    panels first and hands the trailing block over at a panel boundary."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    y = cases.step_spectrum_batch(4 * d + 8, d, 7).double()
    cov = (y.T @ y / y.shape[0])
    cov = cov + 0.01 * cov.diagonal().mean() * torch.eye(d, dtype=torch.float64)
    c32 = cov.float()
    L = nat.lib()
    try:
        L.ptdeco_debug_set(104, 0)     # no Jacobi: the general path also at d <= 96
        L.ptdeco_debug_set(103, rows)  # resident on, rows-per-CTA target
        ev, u = linalg.eigh(c32.to(dev), k=k)
        L.ptdeco_debug_set(102, 0)
        ev_blocked, _ = linalg.eigh(c32.to(dev), k=k)
    finally:
        L.ptdeco_debug_set(103, 4)
        L.ptdeco_debug_set(104, 32)
    cos_ks = [kk for kk in (d // 8, d // 4) if 1 <= kk <= k]
    _check_eigh(c32.double().numpy(), ev.cpu().numpy().astype(np.float64),
                u.cpu().numpy().astype(np.float64), k, cos_ks)
    assert float((ev - ev_blocked).abs().max() / ev_blocked.abs().max()) < 2e-6


def test_eigh_rank_deficient_with_damping(dev):
    """Covariance of rank d/4 plus damping: a (d - d/4)-fold near-degenerate cluster."""
    from ptdeco_b200 import linalg
    d, r = 320, 80
    g = torch.Generator().manual_seed(2)
    y = (torch.randn(4 * d, r, generator=g) @ torch.randn(r, d, generator=g)).float()
    acc = linalg.CovarianceAccumulator(d, dev)
    acc.update(y.to(dev))
    cov = acc.finalize(False, 0.01).clone()
    ev, u = linalg.eigh(cov)
    c64 = cov.double().cpu().numpy()
    _check_eigh(c64, ev.cpu().numpy().astype(np.float64), u.cpu().numpy().astype(np.float64), d, (r // 2, r))
    assert torch.equal(cov, cov.T)


@pytest.mark.parametrize("d", [40, 160])
def test_eigh_degenerate_inputs(dev, d):
    """Zero, scaled-identity and diagonal matrices (all-1x1 blocks after splitting), k = 1."""
    from ptdeco_b200 import linalg
    z = torch.zeros(d, d, device=dev)
    ev, u = linalg.eigh(z)
    assert torch.all(ev == 0) and not torch.isnan(u).any()
    assert (u.T @ u - torch.eye(d, device=dev)).abs().max().item() < 1e-6
    ev, u = linalg.eigh(3.5 * torch.eye(d, device=dev), k=5)
    assert torch.allclose(ev, torch.full_like(ev, 3.5)) and tuple(u.shape) == (d, 5)
    assert (u.T @ u - torch.eye(5, device=dev)).abs().max().item() < 1e-6
    diag = torch.arange(1, d + 1, dtype=torch.float32)
    perm = torch.randperm(d, generator=torch.Generator().manual_seed(0))
    ev, u = linalg.eigh(torch.diag(diag[perm]).to(dev), k=1)
    assert torch.allclose(ev.cpu(), diag) and tuple(u.shape) == (d, 1)
    assert abs(abs(u[:, 0].cpu()[perm.tolist().index(d - 1)].item()) - 1.0) < 1e-6  # e_j of the largest entry
    g = torch.Generator().manual_seed(1)
    a = torch.randn(d, d, generator=g)
    sym = (a @ a.T / d).to(dev)
    ev, u = linalg.eigh(sym, k=1)
    ref = torch.linalg.eigh(sym.double())
    assert abs(u[:, 0].double() @ ref.eigenvectors[:, -1]).item() > 0.99999


def test_eigh_input_not_modified_and_uplo(dev):
    from ptdeco_b200 import linalg
    d = 150
    g = torch.Generator().manual_seed(4)
    a = torch.randn(d, d, generator=g)
    sym = (a + a.T) / 2
    garbage_upper = torch.tril(sym) + torch.triu(torch.randn(d, d, generator=g), 1)  # only lower counts
    x = garbage_upper.to(dev)
    keep = x.clone()
    ev, _ = linalg.eigh(x)
    assert torch.equal(x, keep)
    ref = np.linalg.eigvalsh(sym.double().numpy())
    assert np.abs(ev.cpu().numpy() - ref).max() / np.abs(ref).max() < EVAL_TOL


@pytest.mark.parametrize("in_f,out_f,k", [(48, 96, 24), (64, 176, 32), (192, 768, 96), (256, 1024, 128)])
def test_input_side_eigenvectors_match_output_side(dev, in_f, out_f, k):
    """C = W S W^T: the in x in route (linalg.eigvecs_from_input_covariance) spans the same top-k
    subspace as the eigensolve of the out x out covariance the reference accumulates."""
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(in_f + out_f)
    w = torch.randn(out_f, in_f, generator=g) / in_f ** 0.5
    x = torch.randn(6 * in_f, in_f, generator=g) * torch.logspace(0, -1.5, in_f)
    y = x @ w.T
    acc_s = linalg.CovarianceAccumulator(in_f, dev)
    acc_s.update(x.to(dev))
    u_in = linalg.eigvecs_from_input_covariance(acc_s.finalize(False, 0.0), w.to(dev), k)
    c = (y.double().T @ y.double() / y.shape[0]).numpy()
    c += 0.01 * np.mean(np.diag(c)) * np.eye(out_f)
    u_ref = P.eigenvectors_ascending(c)
    u_in = u_in.double().cpu().numpy()
    assert u_in.shape == (out_f, k)
    assert np.abs(u_in.T @ u_in - np.eye(k)).max() < 2e-5
    for kk in sorted({k, max(1, k // 2), max(1, k // 4)}):
        assert P.min_principal_cosine(P.top_k(u_ref, kk), u_in[:, k - kk:]) >= COS_TOL


# ------------------------------------------------------------------------------------ K4 / K5 / K7
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("out_f,in_f,k", [(32, 64, 32), (96, 48, 13), (192, 768, 96), (576, 192, 100)])
def test_factors_match_oracle(dev, dtype, out_f, in_f, k):
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(out_f + in_f + k)
    w = (torch.randn(out_f, in_f, generator=g) / in_f ** 0.5).to(dtype)
    q, _ = torch.linalg.qr(torch.randn(out_f, out_f, generator=g))
    uk = q[:, out_f - k:].contiguous().to(dtype)
    w1 = linalg.factor_w1(w.to(dev), uk.to(dev))
    deco = linalg.deco_weight(uk.to(dev), w1)
    U, V, deco_ref = P.factors(w.double().numpy(), uk.double().numpy())
    tol = 3e-6 if dtype == torch.float32 else 1e-2  # bf16x3 products + chunked TMEM accumulation
    assert np.abs(w1.double().cpu().numpy() - U.T).max() <= tol * np.abs(U).max()
    # K5 consumes the (possibly bf16-rounded) W1 the kernel produced
    deco_ref2 = uk.double().numpy() @ w1.double().cpu().numpy()
    assert np.abs(deco.double().cpu().numpy() - deco_ref2).max() <= tol * np.abs(deco_ref).max()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("n,in_f,k,out_f", [(1, 64, 8, 32), (16, 192, 48, 160), (300, 200, 50, 168),
                                            (1024, 768, 96, 3072), (2048, 1024, 256, 1024),
                                            # >= 148 token tiles: the persistent phase-overlapped kernel
                                            (19000, 320, 64, 520), (37999, 256, 128, 768), (18944, 128, 100, 256)])
def test_lowrank_forward(dev, dtype, tol, n, in_f, k, out_f):
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(n, in_f, generator=g).to(dtype)
    w1 = (torch.randn(k, in_f, generator=g) / in_f ** 0.5).to(dtype)
    w2 = (torch.randn(out_f, k, generator=g) / k ** 0.5).to(dtype)
    b = torch.randn(out_f, generator=g)
    y = linalg.lowrank_forward(x.to(dev), w1.to(dev), w2.to(dev), b.to(dev))
    ref = (x.double() @ w1.double().T) @ w2.double().T + b.double()
    assert y.dtype == dtype and tuple(y.shape) == (n, out_f)
    assert (y.double().cpu() - ref).abs().max() <= tol * ref.abs().max()


@pytest.mark.parametrize("tile_m", [8, 40, 56, 104, 128, 0])
@pytest.mark.parametrize("n,in_f,k,out_f", [(1000, 256, 64, 520), (5000, 192, 128, 256), (19000, 128, 96, 264),
                                            (150 * 56 + 3, 128, 64, 128), (300, 320, 200, 1000)])
def test_lowrank_forward_balanced_token_tiles(dev, tile_m, n, in_f, k, out_f):
    """The fused / persistent kernels with token tiles of `tile_m` rows (the host normally sizes
    them so the tiles fill whole waves of SMs; 0 = that cost model): X boxes and Y stores are
    clipped to the tile, rows past it are never written, ragged last tiles included."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(n + k + tile_m)
    x = torch.randn(n, in_f, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(k, in_f, generator=g) / in_f ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(out_f, k, generator=g) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(out_f, generator=g)
    L = nat.lib()
    try:
        L.ptdeco_debug_set(207, tile_m)
        xd, w1d, w2d, bd = x.to(dev), w1.to(dev), w2.to(dev), b.to(dev)
        y = linalg.lowrank_forward(xd, w1d, w2d, bd)
        # a sentinel-filled output must come back fully overwritten and nothing else touched
        big = torch.full((n + 16, out_f), 7.0, dtype=torch.bfloat16, device=dev)
        view = big[8:8 + n]
        nat.check(L.ptdeco_lowrank_forward(
            xd.data_ptr(), in_f, w1d.data_ptr(), in_f, w2d.data_ptr(), k, bd.data_ptr(),
            view.data_ptr(), out_f, nat.BF16, n, in_f, k, out_f, None, 0, nat.stream_ptr(dev)),
            "ptdeco_lowrank_forward")
        torch.cuda.synchronize()
    finally:
        L.ptdeco_debug_set(207, 0)
    h = (x.double() @ w1.double().T).to(torch.bfloat16).double()
    ref = h @ w2.double().T + b.double()
    assert (y.double().cpu() - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert torch.equal(view, y) and bool((big[:8] == 7.0).all()) and bool((big[8 + n:] == 7.0).all())


@pytest.mark.parametrize("n,in_f,k,out_f", [(8192, 4096, 128, 4096), (2048, 4096, 32, 4096), (1000, 768, 96, 3072),
                                            (4100, 200, 64, 1024), (300, 4096, 128, 512)])
def test_lowrank_forward_k_split_clusters(dev, n, in_f, k, out_f):
    """K7 with pairs of out-groups sharing GEMM 1 over a 2-CTA cluster (each accumulates H over
    half of `in`, the fp32 partials cross through distributed shared memory): forced wherever the
    shape allows it (knob 208 = 2) and switched off (1), both against the fp64 product of the bf16
    operands; both CTAs of a pair add the halves in the same order, so reruns are bit-identical;
    ragged token counts and an `in` that is not a multiple of the k-block included."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(n + k + out_f)
    x = torch.randn(n, in_f, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(k, in_f, generator=g) / in_f ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(out_f, k, generator=g) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(out_f, generator=g)
    xd, w1d, w2d, bd = x.to(dev), w1.to(dev), w2.to(dev), b.to(dev)
    L = nat.lib()
    outs = {}
    try:
        for mode in (2, 2, 1):
            L.ptdeco_debug_set(208, mode)
            outs.setdefault(mode, []).append(linalg.lowrank_forward(xd, w1d, w2d, bd))
        torch.cuda.synchronize()
    finally:
        L.ptdeco_debug_set(208, 0)
    h = (x.double() @ w1.double().T).to(torch.bfloat16).double()
    ref = h @ w2.double().T + b.double()
    for mode, ys in outs.items():
        for y in ys:
            assert (y.double().cpu() - ref).abs().max() <= 2e-2 * ref.abs().max(), mode
    assert torch.equal(outs[2][0], outs[2][1])
    # H differs from the unsplit sum only by fp32 rounding before its bf16 rounding
    assert (outs[2][0].float() - outs[1][0].float()).abs().max() <= 2e-2 * ref.abs().max()


@pytest.mark.parametrize("n,in_f,k,out_f,bias", [
    (1, 256, 32, 256, False), (5, 320, 96, 1000, True), (16, 4096, 512, 4096, True),
    (33, 768, 200, 3072, True), (128, 2048, 1024, 2048, False), (100, 1024, 1000, 520, True),
    (7, 4096, 40, 14336, True)])
def test_lowrank_forward_decode_kernel(dev, n, in_f, k, out_f, bias):
    """N <= 128: the single-launch weight-streaming kernel (swap-AB, split-`in` phase 1 with fp32
    red.add + ticketed bf16 rounding, grid barrier, phase 2). Forced on for every shape here (small
    factors are normally routed to the fused kernel), compared with the two-launch path and with an
    fp64 product that rounds the rank-k intermediate to bf16 like the module pair does."""
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(1000 * n + k)
    x = torch.randn(n, in_f, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(k, in_f, generator=g) / in_f ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(out_f, k, generator=g) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(out_f, generator=g) if bias else None
    args = (x.to(dev), w1.to(dev), w2.to(dev), None if b is None else b.to(dev))
    from ptdeco_b200 import _native as nat
    L = nat.lib()
    try:
        L.ptdeco_debug_set(200, 1)  # force the decode kernel
        y = linalg.lowrank_forward(*args)
        y_again = linalg.lowrank_forward(*args)  # the workspace is reused: tickets / H must be reset
        L.ptdeco_debug_set(200, 0)
        L.ptdeco_debug_set(201, 1)  # no decode kernel
        y_two_launch = linalg.lowrank_forward(*args)
    finally:
        L.ptdeco_debug_set(200, 0)
        L.ptdeco_debug_set(201, 0)
    h = (x.double() @ w1.double().T).to(torch.bfloat16).double()
    ref = h @ w2.double().T + (0.0 if b is None else b.double())
    scale = ref.abs().max()
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (n, out_f)
    assert (y.double().cpu() - ref).abs().max() <= 1e-2 * scale
    assert (y_again.double().cpu() - ref).abs().max() <= 1e-2 * scale
    assert (y.double() - y_two_launch.double()).abs().max().cpu() <= 2e-2 * scale


@pytest.mark.parametrize("m,n,k", [(8192, 4096, 1024), (1000, 520, 328), (300, 264, 72)])
def test_gemm_cta_pair_matches_single_cta(dev, m, n, k):
    """The 256 x 256 CTA-pair (cta_group::2) instance of the engine against the single-CTA one and
    an fp64 product: ragged M / N / K go through TMA zero fill and the row / column guards."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g).to(torch.bfloat16).to(dev)
    w = torch.randn(n, k, generator=g).to(torch.bfloat16).to(dev)
    L = nat.lib()
    try:
        L.ptdeco_debug_set(7, 0)
        y_pair = linalg.linear_nt(x, w, out_dtype=torch.float32)
        assert L.ptdeco_debug_get(5) == 1
        L.ptdeco_debug_set(7, 1)
        y_single = linalg.linear_nt(x, w, out_dtype=torch.float32)
        assert L.ptdeco_debug_get(5) == 0
    finally:
        L.ptdeco_debug_set(7, 0)
    ref = x.double() @ w.double().T
    assert (y_pair.double() - ref).abs().max() <= 2e-6 * ref.abs().max() * (k ** 0.5)
    assert (y_pair - y_single).abs().max() <= 1e-5 * ref.abs().max()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_lowrank_sequential_matches_torch_sequential(dev, dtype, tol):
    from ptdeco_b200 import modules
    g = torch.Generator().manual_seed(21)
    lin = torch.nn.Sequential(torch.nn.Linear(96, 24, bias=False), torch.nn.Linear(24, 80)).to(dev).to(dtype)
    conv = torch.nn.Sequential(torch.nn.Conv2d(32, 8, 1, bias=False), torch.nn.Conv2d(8, 40, 1)).to(dev).to(dtype)
    holder = torch.nn.ModuleDict({"lin": lin, "conv": conv})
    xl = torch.randn(3, 17, 96, generator=g).to(dtype).to(dev)
    xc = torch.randn(2, 32, 9, 7, generator=g).to(dtype).to(dev)
    with torch.no_grad():
        ref_l, ref_c = lin(xl).float(), conv(xc).float()
        assert modules.fuse_decomposed_modules_in_place(holder) == 2
        out_l, out_c = holder["lin"](xl), holder["conv"](xc)
    assert out_l.shape == ref_l.shape and out_c.shape == ref_c.shape and out_l.dtype == dtype
    assert (out_l.float() - ref_l).abs().max() <= tol * ref_l.abs().max()
    assert (out_c.float() - ref_c).abs().max() <= tol * ref_c.abs().max()


def test_paired_trials_verify_every_layer(dev):
    """_wrap.PairState: one forward of the doubled batch replaces a trial's two forwards only after
    the first batch OF THAT LAYER was evaluated both ways and agreed; a model whose forward mixes
    batch elements, or one that folds the batch into another dimension before the layer (the
    wrapper then sees a leading dim that is not 2b), keeps the reference's two-forward path. The
    trial runs the two-factor op: the layer's weight is never written."""
    import ptdeco_b200.falor.decomposition as F
    from ptdeco_b200 import _wrap, linalg

    class Net(torch.nn.Module):
        def __init__(self, mode):
            super().__init__()
            self.fc1, self.fc2, self.mode = torch.nn.Linear(24, 40), torch.nn.Linear(40, 7), mode

        def forward(self, x):
            if self.mode == "fold":  # [B, 24] -> [B/2, 48] -> fc1 on [.., 24] rows in another order
                h = torch.relu(self.fc1(x.reshape(-1, 2, 24).transpose(0, 1))).transpose(0, 1).reshape(x.shape[0], 40)
            else:
                h = torch.relu(self.fc1(x))
            if self.mode == "mix":
                h = h - h.mean(0, keepdim=True)
            return self.fc2(h)

    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 24, generator=g).to(dev)
    for mode, want in (("plain", "on"), ("mix", "off"), ("fold", "off")):
        torch.manual_seed(3)
        net = Net(mode).to(dev).eval()
        F._wrap_in_place(net, "fc1")
        wrapper = net.get_submodule("fc1")
        w = wrapper.get_weight_copy()
        q, _ = torch.linalg.qr(torch.randn(40, 40, generator=g))
        uk = q[:, :10].contiguous().to(dev)
        w1 = linalg.factor_w1(w, uk)
        deco = uk @ w1
        st = _wrap.PairState(net)
        st.begin_layer()
        with torch.no_grad():
            y_deco, y_orig = st.forward_pair(net, wrapper, x, (w1, uk))
            assert st.mode == want, (mode, st.mode, st.probe)
            ref_orig = net(x)
            wrapper.set_weight(deco)
            ref_deco = net(x)
            wrapper.set_weight(w)
            assert torch.allclose(y_deco, ref_deco, atol=2e-5) and torch.allclose(y_orig, ref_orig, atol=1e-6)
            y_deco2, y_orig2 = st.forward_pair(net, wrapper, x, (w1, uk))  # second batch: paired iff verified
            assert (st.paired_forwards == 1) == (want == "on")
            assert torch.allclose(y_deco2, ref_deco, atol=2e-5) and torch.allclose(y_orig2, ref_orig, atol=1e-5)
            assert torch.equal(wrapper.get_weight_copy(), w) and wrapper.trial_factors is None
            st.begin_layer()
            assert st.mode == "unverified"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_trial_factors_match_materialised_weight(dev, dtype):
    """f2: a rank trial through the wrapper's two-factor op (ptdeco_lowrank_forward) gives the
    layer output of the reference's route, deco_weight = uk uk^T W copied into the layer
    (F:347-348,222 / D:427-429,262), for Linear and 1x1 conv targets."""
    import ptdeco_b200.falor.decomposition as F
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(17)
    lin = torch.nn.Sequential(torch.nn.Linear(96, 160)).to(dev).to(dtype).eval()
    conv = torch.nn.Sequential(torch.nn.Conv2d(32, 48, 1)).to(dev).to(dtype).eval()
    for net, x in ((lin, torch.randn(3, 7, 96, generator=g)), (conv, torch.randn(2, 32, 5, 6, generator=g))):
        x = x.to(dev).to(dtype)
        F._wrap_in_place(net, "0")
        wr = net.get_submodule("0")
        w = wr.get_weight_copy()
        q, _ = torch.linalg.qr(torch.randn(w.shape[0], w.shape[0], generator=g))
        uk = q[:, :24].contiguous().to(dev)
        w1 = linalg.factor_w1(w.float(), uk)
        with torch.no_grad():
            wr.set_trial(w1, uk)
            y = net(x).float()
            wr.clear_trial()
            wr.set_weight((uk @ w1).to(dtype))
            ref = net(x).float()
            wr.set_weight(w)
        tol = 1e-5 if dtype == torch.float32 else 3e-2
        assert y.shape == ref.shape and (y - ref).abs().max() <= tol * ref.abs().max()


def test_strided_conv_covariance_matches_reference_formulation(dev):
    """Strided / padded 1x1 conv targets: the covariance is the reference's, formed from ALL input
    positions (F:125-126,159), not from the (subsampled / bordered) layer output."""
    import ptdeco_b200.falor.decomposition as F
    from oracle import drivers as OD
    for kw in (dict(stride=2), dict(padding=1)):
        torch.manual_seed(7)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.GELU(),
                                  torch.nn.Conv2d(16, 24, 1, **kw), torch.nn.Flatten(), torch.nn.LazyLinear(5)).eval()
        stream_a = cases.IndexedStream(lambda i: cases.streams.lowrank_image_batch(9, i, 6, 3, 8, 12))
        stream_b = cases.IndexedStream(lambda i: cases.streams.lowrank_image_batch(9, i, 6, 3, 8, 12))
        with torch.no_grad():
            net(next(cases.IndexedStream(lambda i: cases.streams.lowrank_image_batch(9, i, 6, 3, 8, 12))))  # lazy init
        w = net[2].weight.detach()[..., 0, 0].clone()
        tap = OD._Tap(net[2])
        net[2] = tap
        with torch.no_grad():
            u_ref = OD.falor_eigenvectors(net, "2", stream_a, w.numpy(), 3, True, False, True)
        net[2] = tap.inner
        net.to(dev)
        F._wrap_in_place(net, "2")
        with torch.no_grad():
            u = F._compute_decompositon_of_covariance_matrix(
                root_module=net, decomposed_submodule_name="2", data_iterator=stream_b, weight=w.to(dev),
                num_data_steps=3, device=dev, use_float64=True, use_mean=False, use_damping=True)
        assert stream_a.position == stream_b.position == 3
        for k in (2, 4, 8):
            assert P.min_principal_cosine(P.top_k(u_ref, k), P.top_k(u.double().cpu().numpy(), k)) >= COS_TOL


@pytest.mark.parametrize("input_side", [False, True])
def test_bf16_covariance_module_matches_fp32_einsum_of_same_activations(dev, input_side):
    """dwain's CovarianceComputingLinearModule on a bf16 layer against einsum(y.float(), y.float())
    of the SAME bf16 activations (the reference's own bf16 einsum rounds every
    per-step product to bf16 and is ~1e-3 noisy, so it is not the yardstick for a 1e-5 bar)."""
    import ptdeco_b200.dwain.decomposition as D
    g = torch.Generator().manual_seed(23)
    in_f, out_f = (64, 192) if input_side else (192, 96)
    lin = torch.nn.Linear(in_f, out_f, bias=False).to(dev).to(torch.bfloat16)
    mod = D.CovarianceComputingLinearModule(lin.weight, None, True, 16 if input_side else None)
    assert mod.input_side == input_side
    ref = torch.zeros(mod.acc.d, mod.acc.d, dtype=torch.float64, device=dev)
    steps = 3
    with torch.no_grad():
        for _ in range(steps):
            x = (torch.randn(2, 50, in_f, generator=g) * torch.logspace(0, -1, in_f)).to(torch.bfloat16).to(dev)
            y = mod(x)  # no bias: the module's output IS the bf16 y = x W^T it folded in
            assert y.dtype == torch.bfloat16 and tuple(y.shape) == (2, 50, out_f)
            rows = (x if input_side else y).reshape(-1, mod.acc.d).double()
            ref += rows.T @ rows / rows.shape[0]
            y_ref = torch.nn.functional.linear(x.float(), lin.weight.float())
            assert (y.float() - y_ref).abs().max() <= 1e-2 * y_ref.abs().max()
    assert mod.num_data_steps == steps
    cov = mod.acc.finalize(False, 0.0)
    assert _rel(cov.double().cpu().numpy(), (ref / steps).cpu().numpy()) < COV_TOL


def test_deterministic_flag_is_per_call_and_bit_reproducible(dev):
    """PTDECO_FLAG_DETERMINISTIC (the `flags` of the *_ex entry points, set by
    nat.set_deterministic): no split-K, so repeated runs agree bit for bit -- the SYRK of a short,
    wide batch (where the default splits the token dimension over CTAs), a GEMM, and the whole
    eigensolver; and the flag is not process state of the library (two interleaved calls with
    different flags do not disturb each other)."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(77)
    y = torch.randn(8192, 256, generator=g).to(torch.bfloat16).to(dev)
    w = torch.randn(300, 256, generator=g).to(dev)

    def run():
        acc = linalg.CovarianceAccumulator(256, dev)
        acc.update(y)
        cov = acc.finalize(False, 0.01).clone()
        ev, u = linalg.eigh(cov, k=64)
        f = linalg.linear_nt(w, cov)
        return cov, ev, u, f

    try:
        nat.set_deterministic(True)
        a = run()
        nat.set_deterministic(False)
        other = run()  # default mode in between
        nat.set_deterministic(True)
        b = run()
    finally:
        nat.set_deterministic(False)
    for x1, x2 in zip(a, b):
        assert torch.equal(x1, x2)
    assert _rel(other[0].cpu().numpy(), a[0].cpu().numpy()) < 1e-6  # same numbers up to summation order


def test_canonical_shards_make_the_covariance_independent_of_the_gpu_count(dev):
    """in deterministic mode calibration step i goes to canonical shard i mod 8
    whatever GPU ran it and the shards are added in index order, so a 1-, 2-, 4- or 8-way split of
    the steps yields the same bits. The split is emulated on one GPU with the accumulators of every
    "rank" (tools/dist_check.py repeats it across real GPUs with NCCL point-to-point gathers); a
    plain sharded sum of the same steps differs in the low bits."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    d, n, steps = 320, 1024, 11
    g = torch.Generator().manual_seed(5)
    ys = [(torch.randn(n, d, generator=g) * 10.0 ** float(torch.randint(-2, 3, (1,), generator=g)))
          .to(torch.bfloat16).to(dev) for _ in range(steps)]
    try:
        nat.set_deterministic(True)
        V = linalg.canonical_shards(1)
        assert V == linalg.CANONICAL_SHARDS == 8
        with pytest.raises(ValueError):
            linalg.canonical_shards(3)
        covs = {}
        for world in (1, 2, 4, 8):
            accs = [linalg.CovarianceAccumulator(d, dev, shards=V, rank=r, world=world) for r in range(world)]
            for i, y in enumerate(ys):
                accs[i % world].update(y)            # default step numbering: rank + j * world
            owner = accs[0]
            if world > 1:                            # what parallel.gather_shards_to does over NCCL
                owner.C.zero_()
                for v in range(min(V, steps)):
                    owner.C += accs[v % world].shard_C[v]
                owner._collapsed, owner.steps = True, steps
            covs[world] = owner.finalize(False, 0.01).clone()
        # explicit step numbers (the dwain precompute passes them) give the same shards
        acc = linalg.CovarianceAccumulator(d, dev, shards=V, rank=1, world=2)
        for i in range(1, steps, 2):
            acc.update(ys[i], step=i)
        ref = linalg.CovarianceAccumulator(d, dev, shards=V, rank=1, world=2)
        for i in range(1, steps, 2):
            ref.update(ys[i])
        assert all((a is None) == (b is None) and (a is None or torch.equal(a, b))
                   for a, b in zip(acc.shard_C, ref.shard_C))
        plain = linalg.CovarianceAccumulator(d, dev)
        for y in ys:
            plain.update(y)
        cov_plain = plain.finalize(False, 0.01).clone()
    finally:
        nat.set_deterministic(False)
    assert linalg.canonical_shards(2) == 1  # default mode: plain partial sums
    for world in (2, 4, 8):
        assert torch.equal(covs[world], covs[1]), world
    assert _rel(cov_plain.cpu().numpy(), covs[1].cpu().numpy()) < 1e-6
    cov_ref = sum((y.double().T @ y.double()) / n for y in ys) / steps
    cov_ref += 0.01 * cov_ref.diagonal().mean() * torch.eye(d, device=dev, dtype=torch.float64)
    assert _rel(covs[1].double().cpu().numpy(), cov_ref.cpu().numpy()) < COV_TOL


def test_staged_updates_survive_in_place_reuse_of_the_batch(dev):
    """The staging copies and group SYRKs of CovarianceAccumulator run on side streams; the batch
    handed to update() may still be overwritten in place right afterwards (an inplace activation
    behind a hooked layer output), reused for the next step, or freed. Covariance must be that of
    the values at call time, for several interleaved accumulators, with a flush in the middle."""
    from ptdeco_b200 import linalg
    d, n, steps = 1024, 2048, 21
    g = torch.Generator(device=dev).manual_seed(11)
    accs = [linalg.CovarianceAccumulator(d, dev, defer_rows=8 * n) for _ in range(3)]
    refs = [torch.zeros(d, d, dtype=torch.float64, device=dev) for _ in accs]
    buf = torch.empty(n, d, dtype=torch.bfloat16, device=dev)   # one buffer reused for every batch
    for i in range(steps):
        for a, (acc, ref) in enumerate(zip(accs, refs)):
            buf.copy_(torch.randn(n, d, generator=g, device=dev) * (1.0 + a))
            ref += buf.double().T @ buf.double() / n
            acc.update(buf)
            buf.relu_()                                         # in place, right after the call
            tmp = torch.randn(n, d, generator=g, device=dev).to(torch.bfloat16)
            ref += tmp.double().T @ tmp.double() / n
            acc.update(tmp)
            del tmp                                             # freed while its copy may be pending
        if i == 9:
            accs[1].flush()
    for acc, ref in zip(accs, refs):
        assert acc.steps == 2 * steps
        cov = acc.finalize(False, 0.0)
        want = ref / (2 * steps)
        assert _rel(torch.tril(cov).double().cpu().numpy(), torch.tril(want).cpu().numpy()) < COV_TOL


def test_workspaces_are_per_stream(dev):
    """the scratch buffer (the decode kernel keeps its grid-barrier / ticket words and
    the rank-k intermediate in it) is keyed by (device, stream), so forwards issued on two
    streams concurrently neither share control words nor free each other's buffer."""
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg
    g = torch.Generator().manual_seed(3)
    shapes = [(16, 2048, 512, 2048), (48, 1024, 1024, 4096)]
    data = []
    for n, fin, k, fout in shapes:
        x = torch.randn(n, fin, generator=g).to(torch.bfloat16).to(dev)
        w1 = (torch.randn(k, fin, generator=g) / fin ** 0.5).to(torch.bfloat16).to(dev)
        w2 = (torch.randn(fout, k, generator=g) / k ** 0.5).to(torch.bfloat16).to(dev)
        h = (x.double() @ w1.double().T).to(torch.bfloat16).double()
        data.append((x, w1, w2, h @ w2.double().T))
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    outs = [[], []]
    L = nat.lib()
    try:
        L.ptdeco_debug_set(200, 1)  # decode kernel for both
        torch.cuda.synchronize()
        for _ in range(25):
            for si, st in enumerate(streams):
                with torch.cuda.stream(st):
                    x, w1, w2, _ = data[si]
                    outs[si].append(linalg.lowrank_forward(x, w1, w2, None))
        torch.cuda.synchronize()
    finally:
        L.ptdeco_debug_set(200, 0)
    keys = {k for k in nat.WORKSPACE._buf if k[1] in (streams[0].cuda_stream, streams[1].cuda_stream)}
    assert len(keys) == 2
    for si in range(2):
        ref = data[si][3]
        for y in outs[si]:
            assert (y.double() - ref).abs().max() <= 1e-2 * ref.abs().max()


def test_kernels_run_on_a_non_current_device():
    """kernel attributes / SM counts are cached per device and the Python layer makes
    the tensor's device current for the call. Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from ptdeco_b200 import linalg
    d1 = torch.device("cuda:1")
    assert torch.cuda.current_device() == 0
    g = torch.Generator().manual_seed(9)
    y = torch.randn(600, 320, generator=g)
    acc = linalg.CovarianceAccumulator(320, d1)
    acc.update(y.to(d1))
    cov = acc.finalize(False, 0.01)
    ref = (y.double().T @ y.double() / 600).numpy()
    ref = ref + 0.01 * np.mean(np.diag(ref)) * np.eye(320)
    assert _rel(cov.cpu().numpy(), ref) < COV_TOL
    ev, u = linalg.eigh(cov, k=40)
    assert u.device == d1 and np.abs(ev.cpu().numpy() - np.linalg.eigvalsh(ref)).max() / ref.max() < EVAL_TOL
    x = torch.randn(300, 320, generator=g).to(torch.bfloat16).to(d1)
    w1 = torch.randn(64, 320, generator=g).to(torch.bfloat16).to(d1)
    w2 = torch.randn(512, 64, generator=g).to(torch.bfloat16).to(d1)
    yl = linalg.lowrank_forward(x, w1, w2, None)
    refl = (x.double() @ w1.double().T).to(torch.bfloat16).double() @ w2.double().T
    assert (yl.double() - refl).abs().max() <= 2e-2 * refl.abs().max()
    assert torch.cuda.current_device() == 0


# ------------------------------------------------------------------------------------ K6
def test_metrics_match_reference(dev, golden_dir):
    from ptdeco_b200 import utils
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    fx, fy = torch.from_numpy(g["falor_logits_x"]).to(dev), torch.from_numpy(g["falor_logits_y"]).to(dev)
    dx, dy = torch.from_numpy(g["dwain_logits_x"]).to(dev), torch.from_numpy(g["dwain_logits_y"]).to(dev)
    assert utils.calc_per_channel_noise_to_signal_ratio(x=fx, y=fy, non_channel_dim=(0,)).item() == \
        pytest.approx(float(g["falor_logits_nsr"]), rel=1e-5)
    assert utils.calc_per_channel_noise_to_signal_ratio(x=dx, y=dy, non_channel_dim=(0, 1)).item() == \
        pytest.approx(float(g["dwain_logits_nsr"]), rel=1e-5)
    assert utils.calc_kl_loss(fx, fy).item() == pytest.approx(float(g["falor_logits_kl"]), rel=1e-5)
    np.testing.assert_allclose(utils.calc_kl_divergence(fx, fy).cpu().numpy(), g["falor_logits_kl_rows"],
                               rtol=2e-4, atol=1e-6)


# ------------------------------------------------------------------------------------ primitives
@pytest.mark.parametrize("method", ["falor", "dwain"])
@pytest.mark.parametrize("kind,bound", [("linear", 1e-6), ("conv", 9e-4)])
def test_reference_primitive_tests(dev, golden_dir, method, kind, bound):
    """Mirror of tests/test_deco_primitives_{falor,dwain}.py (full-rank reconstruction), with the
    reference's own GPU bounds, plus covariance / eigenvector parity against the captured golden."""
    import ptdeco_b200.dwain.decomposition as D
    import ptdeco_b200.falor.decomposition as F
    from ptdeco_b200 import linalg, utils
    gold = np.load(os.path.join(golden_dir, f"prim_{method}_{kind}.npz"))
    net, stream = cases.primitive_case(kind, dict_input=(method == "dwain"))
    net.to(dev)
    mod = F if method == "falor" else D
    x = utils.to_device(next(stream), dev)
    with torch.no_grad():
        y0 = net(x)
    mod._wrap_in_place(net, "mod")
    w = net.mod.get_weight_copy()
    with torch.no_grad():
        if method == "falor":
            u = F._compute_decompositon_of_covariance_matrix(
                root_module=net, decomposed_submodule_name="mod", data_iterator=stream, weight=w,
                num_data_steps=8, device=dev, use_float64=True, use_mean=False, use_damping=True)
        else:
            u = D._compute_covariance_matrix_decomposition(
                root_module=net, decomposed_submodule_name="mod", data_iterator=stream, weight=w,
                num_data_steps=8, device=dev, decompose_in_float64=True)
        assert tuple(u.shape) == (32, 32)
        for k in (4, 8, 16):
            assert P.min_principal_cosine(P.top_k(gold["u"].astype(np.float64), k),
                                          P.top_k(u.double().cpu().numpy(), k)) >= COS_TOL
        uk = u[:, u.shape[1] - 32:].to(torch.float)
        w1 = linalg.factor_w1(w, uk)
        new = net.get_submodule("mod").get_decomposed_module(u=w1, v=uk)
        new.to(dev)
    mod._unwrap_in_place(net, "mod")
    utils.replace_submodule_in_place(net, "mod", new)
    with torch.no_grad():
        y1 = net(x)
    assert (y0 - y1).abs().max().item() < bound


# ------------------------------------------------------------------------------------ drivers
def _ranks(cfg):
    return {n: c["modules"]["0"].get("out_features", c["modules"]["0"].get("out_channels"))
            for n, c in cfg.items()}


@pytest.mark.parametrize("name", ["mlp", "convmlp", "deit_small"])
def test_falor_decompose_in_place_matches_reference(dev, golden_dir, name):
    import ptdeco_b200.falor as falor
    from ptdeco_b200 import utils
    gold = json.load(open(os.path.join(golden_dir, f"falor_{name}.json")))
    model, stream, kw = cases.falor_case(name)
    model.to(dev)
    trace = []
    cfg = falor.decompose_in_place(module=model, device=dev, data_iterator=stream, trace=trace, **kw)
    assert stream.position == gold["stream_position"]  # iterator consumption order (fact 10)
    assert [(t["name"], t["rank"]) for t in trace] == [(t["name"], t["rank"]) for t in gold["trace"]]
    for t, g in zip(trace, gold["trace"]):
        assert t["nsr"] == pytest.approx(g["nsr"], rel=5e-3, abs=1e-6)
        assert t["kl"] == pytest.approx(g["kl"], rel=1e-2, abs=1e-6)
    assert list(cfg.keys()) == list(gold["decompose_config"].keys())
    assert _ranks(cfg) == _ranks(gold["decompose_config"])
    for n in cfg:
        a = json.loads(json.dumps(cfg[n]))
        b = gold["decompose_config"][n]
        ma, mb = a.pop("__meta__"), b.pop("__meta__")
        assert a == b
        assert ma["proportion"] == mb["proportion"]
        assert ma["nsr_final"] == pytest.approx(mb["nsr_final"], rel=5e-3, abs=1e-6)
    # the config round-trips into a fresh model and the state dict loads strictly (U/m:114-130)
    fresh, _, _ = cases.falor_case(name)
    utils.apply_decompose_config_in_place(fresh, json.loads(json.dumps(cfg)))
    fresh.load_state_dict(model.state_dict(), strict=True)
    fresh.to(dev).eval()
    xb = next(stream).to(dev)
    with torch.no_grad():
        torch.testing.assert_close(fresh(xb), model(xb))


def test_falor_on_a_bf16_model_keeps_the_model_dtype(dev):
    """falor's replacement modules follow the layer's dtype (the reference leaves fp32
    factors in place, F:387, which breaks the next forward of a bf16 model); the decomposed bf16
    model runs, its config round-trips, and its logits stay close to the original's."""
    import ptdeco_b200.falor as falor
    from ptdeco_b200 import utils
    model, stream, kw = cases.falor_case("mlp")
    model = model.to(torch.bfloat16).to(dev)
    ref_model, _, _ = cases.falor_case("mlp")
    ref_model = ref_model.to(torch.bfloat16).to(dev)

    class Cast:
        def __init__(self, it):
            self.it = it

        def __iter__(self):
            return self

        def __next__(self):
            return next(self.it).to(torch.bfloat16)

    kw = dict(kw, nsr_final_threshold=0.08, kl_final_threshold=0.05)
    cfg = falor.decompose_in_place(module=model, device=dev, data_iterator=Cast(stream), **kw)
    assert len(cfg) >= 1
    for name in cfg:
        sub = model.get_submodule(name)
        assert all(p.dtype == torch.bfloat16 for p in sub.parameters())
    x = next(Cast(stream)).to(dev)
    with torch.no_grad():
        y, y_ref = model(x).float(), ref_model(x).float()
    assert y.dtype == torch.float32 and torch.isfinite(y).all()
    assert float((y - y_ref).norm() / y_ref.norm()) < 0.5
    fresh, _, _ = cases.falor_case("mlp")
    utils.apply_decompose_config_in_place(fresh, json.loads(json.dumps(cfg)))
    fresh.to(torch.bfloat16).load_state_dict(model.state_dict(), strict=True)


def test_falor_edge_cases_match_oracle(dev):
    """Rank-1 target, grouped / 3x3 convs, blacklist (incl. a name that does not exist), a layer
    that cannot reduce parameters, the proportion_threshold gate: product (GPU) vs oracle (CPU)."""
    import ptdeco_b200.falor as falor
    from oracle import drivers as OD
    m_cpu, s_cpu, kw = cases.falor_case("edge")
    t_cpu = []
    cfg_cpu = OD.falor_decompose_in_place(module=m_cpu, data_iterator=s_cpu, trace=t_cpu, **kw)
    m_gpu, s_gpu, kw = cases.falor_case("edge")
    m_gpu.to(dev)
    t_gpu = []
    cfg_gpu = falor.decompose_in_place(module=m_gpu, device=dev, data_iterator=s_gpu, trace=t_gpu, **kw)
    assert s_gpu.position == s_cpu.position
    assert [(t["name"], t["rank"], t["accepted"]) for t in t_gpu] == \
        [(t["name"], t["rank"], t["accepted"]) for t in t_cpu]
    assert {t["name"] for t in t_gpu} == {"pw", "wide", "tiny", "head"}  # gate: rank 1; keep: blacklisted
    assert list(cfg_gpu.keys()) == list(cfg_cpu.keys()) and _ranks(cfg_gpu) == _ranks(cfg_cpu)
    assert json.loads(json.dumps(cfg_gpu))["pw"]["modules"]["0"]["type"] == "Conv2d"
    assert isinstance(m_gpu.gate, torch.nn.Linear) and isinstance(m_gpu.keep, torch.nn.Linear)
    for t, o in zip(t_gpu, t_cpu):
        # `tiny` / `head` see an effectively rank-1 input: their second eigenvector comes out of the
        # degenerate damping-level cluster (eigenvalue spread 1e-9), so only decisions are compared
        if t["name"] in ("pw", "wide"):
            assert t["nsr"] == pytest.approx(o["nsr"], rel=2e-2, abs=2e-6)


def test_dwain_edge_cases_match_oracle(dev):
    """min_rank above some layers' full rank (no trial runs), a candidate rank that does not reduce
    parameters (skipped without consuming metric batches), float32 decomposition flag."""
    import ptdeco_b200.dwain as dwain
    from oracle import drivers as OD
    outs = []
    for side in ("cpu", "gpu"):
        model, stream, mstream, kw = cases.dwain_case("llama_tiny")
        kw.update(min_rank=20, decompose_in_float64=False, nsr_final_threshold=0.05,
                  blacklisted_module_names=["lm_head", "model.layers.0.mlp.up_proj"])
        trace = []
        if side == "cpu":
            cfg = OD.dwain_decompose_in_place(
                module=model, data_iterator=stream, metric_iterator=mstream,
                loss_fn=cases.dwain_loss_fn("llama_tiny"), finetune_fn=lambda m, d, n: m, trace=trace, **kw)
        else:
            model.to(dev)
            cfg = dwain.decompose_in_place(
                module=model, device=dev, data_iterator=stream, metric_iterator=mstream,
                loss_fn=cases.dwain_loss_fn("llama_tiny"), finetune_fn=lambda m, d, n: m, trace=trace, **kw)
        outs.append((cfg, trace, stream.position, mstream.position))
    (c0, t0, p0, q0), (c1, t1, p1, q1) = outs
    assert (p0, q0) == (p1, q1)
    assert [(t["name"], t["rank"], t["accepted"]) for t in t0] == [(t["name"], t["rank"], t["accepted"]) for t in t1]
    assert list(c0.keys()) == list(c1.keys()) and _ranks(c0) == _ranks(c1)
    assert "model.layers.0.mlp.up_proj" not in c1
    for n in c1:
        assert c1[n]["__meta__"]["drop_in_params"] == c0[n]["__meta__"]["drop_in_params"]


# Trials of the golden runs whose accept / reject decision sits within 0.5 % of a threshold (a
# bisection ends next to the threshold by construction): (layer, tested rank) -> relative margin
# |metric - threshold| / threshold of the reference's own run. Only after one of THESE trials may a
# layer's trial sequence leave the golden one; everything else must be identical. The list is
# re-derived from the committed golden traces below, so it cannot drift.
FRAGILE_TRIALS = {
    "deit_tiny": {("blocks.0.mlp.fc1", 114): 0.0007, ("blocks.6.mlp.fc1", 24): 0.0028},
    "convnext_tiny": {("features.5.5.block.5", 17): 0.0004, ("features.5.6.block.3", 96): 0.0016,
                      ("features.7.1.block.5", 216): 0.0019, ("features.7.1.block.3", 417): 0.0021,
                      ("features.7.2.block.5", 183): 0.0028, ("features.5.4.block.3", 81): 0.0029,
                      ("features.5.3.block.5", 30): 0.0034, ("features.7.1.block.3", 414): 0.0034,
                      ("features.7.0.block.5", 240): 0.0038, ("features.5.1.block.5", 44): 0.0049},
}
FRAGILE_MARGIN = 0.005


def _golden_margin(t, nsr_thr, kl_thr):
    """Relative distance of a golden trial from flipping its decision (accept iff both below)."""
    mn, mk = abs(t["nsr"] - nsr_thr) / nsr_thr, abs(t["kl"] - kl_thr) / kl_thr
    if t["nsr"] < nsr_thr and t["kl"] < kl_thr:
        return min(mn, mk)
    return max(m for m, bad in ((mn, t["nsr"] >= nsr_thr), (mk, t["kl"] >= kl_thr)) if bad)


@pytest.mark.parametrize("name", ["deit_tiny", "convnext_tiny"])
def test_falor_baseline_configs_match_reference(dev, golden_dir, name):
    """BASELINE.json configs[0] (DeiT-tiny layout, (5,3,224,224) inputs) and configs[1]
    (torchvision convnext_tiny): the whole trial sequence (339 / 285 rank trials) and the chosen
    ranks must be the unmodified reference's. No blanket tolerance on ranks: a layer may differ
    only downstream of a trial listed by name in FRAGILE_TRIALS (golden margin < 0.5 %)."""
    import ptdeco_b200.falor as falor
    gold = json.load(open(os.path.join(golden_dir, f"falor_{name}.json")))
    model, stream, kw = cases.falor_case(name)
    nsr_thr, kl_thr = kw["nsr_final_threshold"], kw["kl_final_threshold"]
    derived = {(g["name"], g["rank"]) for g in gold["trace"] if _golden_margin(g, nsr_thr, kl_thr) < FRAGILE_MARGIN}
    assert derived == set(FRAGILE_TRIALS[name]), sorted(derived ^ set(FRAGILE_TRIALS[name]))
    model.to(dev)
    trace = []
    cfg = falor.decompose_in_place(module=model, device=dev, data_iterator=stream, trace=trace, **kw)
    assert stream.position == gold["stream_position"]  # trial COUNTS are data independent
    by_layer_gold: dict = {}
    for g in gold["trace"]:
        by_layer_gold.setdefault(g["name"], []).append(g)
    by_layer_mine: dict = {}
    for t in trace:
        by_layer_mine.setdefault(t["name"], []).append(t)
    assert list(by_layer_mine) == list(by_layer_gold)
    diverged = []
    for name_, gl in by_layer_gold.items():
        ml = by_layer_mine[name_]
        assert len(ml) == len(gl)
        if [t["rank"] for t in ml] == [g["rank"] for g in gl]:
            continue
        first = next(i for i, (t, g) in enumerate(zip(ml, gl)) if t["rank"] != g["rank"])
        assert first > 0, name_  # the first tested rank is data independent
        flipped = gl[first - 1]  # the decision that went the other way
        assert (name_, flipped["rank"]) in FRAGILE_TRIALS[name], (name_, flipped, ml[first - 1])
        diverged.append(name_)
    ranks, granks = _ranks(cfg), _ranks(gold["decompose_config"])
    assert all(ranks.get(n) == granks.get(n) for n in set(ranks) | set(granks) if n not in diverged)
    # metrics of the trials both runs evaluated at the same rank. Decision-relevant ones (golden NSR
    # within 20 % of the threshold) must agree tightly; the rest of the distribution is bounded by
    # its 90th percentile (ranks cutting through a flat, ill-conditioned part of a random-init
    # spectrum have subspace-dependent NSR: the reference's own fp32 / fp64 runs differ there too).
    pairs = [(t, g) for t, g in zip(trace, gold["trace"]) if (t["name"], t["rank"]) == (g["name"], g["rank"])]
    assert len(pairs) >= 0.95 * len(gold["trace"])
    rel = sorted(abs(t["nsr"] - g["nsr"]) / max(g["nsr"], 1e-9) for t, g in pairs)
    near = [abs(t["nsr"] - g["nsr"]) / g["nsr"] for t, g in pairs if abs(g["nsr"] - nsr_thr) < 0.2 * nsr_thr]
    assert rel[len(rel) // 2] < 1e-3, rel[len(rel) // 2]
    assert rel[int(0.9 * len(rel))] < 5e-2, rel[int(0.9 * len(rel))]
    assert near and max(near) < 5e-2, max(near)
    print(f"falor {name}: {len(diverged)} diverged layers {diverged}; rel NSR diff median {rel[len(rel) // 2]:.2e} "
          f"p90 {rel[int(0.9 * len(rel))]:.2e} max {rel[-1]:.2e}; near-threshold max {max(near):.2e} over {len(near)}")


@pytest.mark.parametrize("name", list(cases.DWAIN_CASES))
def test_dwain_decompose_in_place_matches_reference(dev, golden_dir, name):
    import ptdeco_b200.dwain as dwain
    gold = json.load(open(os.path.join(golden_dir, f"dwain_{name}.json")))
    model, stream, mstream, kw = cases.dwain_case(name)
    model.to(dev)
    trace = []
    cfg = dwain.decompose_in_place(module=model, device=dev, data_iterator=stream,
                                   metric_iterator=mstream, loss_fn=cases.dwain_loss_fn(name),
                                   finetune_fn=lambda m, d, names: m, trace=trace, **kw)
    assert stream.position == gold["stream_position"]
    assert mstream.position == gold["metric_stream_position"]
    assert [(t["name"], t["rank"]) for t in trace] == [(t["name"], t["rank"]) for t in gold["trace"]]
    # bf16 cases: the golden run is the reference's bf16 arithmetic on CPU (every per-step y^T y
    # rounded to bf16, D:152; bf16 factor GEMMs, D:423-429); the product forms exact bf16 products
    # with fp32 accumulation and the user model runs on cuBLAS instead of oneDNN. The reference's
    # own fp32 and bf16 runs of this model differ by up to 7 % in NSR; decisions have >= 1.9x margin.
    rel = 0.15 if "bf16" in name else 1e-2
    for t, g in zip(trace, gold["trace"]):
        assert t["nsr"] == pytest.approx(g["nsr"], rel=rel, abs=1e-6)
    assert list(cfg.keys()) == list(gold["decompose_config"].keys())  # reversed module order
    assert _ranks(cfg) == _ranks(gold["decompose_config"])
    for n in cfg:
        assert cfg[n]["__meta__"]["drop_in_params"] == gold["decompose_config"][n]["__meta__"]["drop_in_params"]


# ------------------------------------------------------------------------------------ full size
def test_full_size_properties_d4096(dev):
    """BASELINE-size checks through size-independent properties: trace(C) = mean ||y||^2, symmetry,
    linearity in the number of steps; eigh: orthonormal top-k and small residual."""
    from ptdeco_b200 import linalg
    d, n, k = 4096, 8192, 512
    y = cases.step_spectrum_batch(n, d, 0).to(torch.bfloat16).to(dev)
    acc = linalg.CovarianceAccumulator(d, dev)
    acc.update(y)
    acc.update(y)
    cov = acc.finalize(False, 0.0)
    tr = cov.diagonal().double().sum().item()
    expect = (y.double() ** 2).sum().item() / n
    assert abs(tr - expect) / expect < 1e-5
    assert torch.equal(cov, cov.T)
    ev, u = linalg.eigh(cov, k=k)
    ud = u.double()
    assert (ud.T @ ud - torch.eye(k, dtype=torch.float64, device=dev)).abs().max().item() < 5e-5
    assert abs(ev.double().sum().item() - tr) / tr < 1e-4
    resid = (cov.double() @ ud - ud * ev[d - k:].double()).abs().max().item() / ev.max().item()
    assert resid < 5e-5


def test_full_size_syrk_d14336_properties(dev):
    """The bench's dominant launch (d = 14336, 16384 staged bf16 tokens, CTA-pair kernel) checked
    through size-independent properties against fp64 on the device: trace(C) = mean ||y||^2,
    C v = Y^T (Y v) / N for random v, symmetry after finalize, and linearity in the batches
    (accumulating two batches = one launch over their concatenation)."""
    from ptdeco_b200 import linalg
    d, n = 14336, 8192
    g = torch.Generator(device=dev).manual_seed(99)
    scale = torch.logspace(0, -1.5, d, device=dev)
    ya = (torch.randn(n, d, generator=g, device=dev) * scale).to(torch.bfloat16)
    yb = (torch.randn(n, d, generator=g, device=dev) * scale).to(torch.bfloat16)
    acc = linalg.CovarianceAccumulator(d, dev, defer_rows=2 * n)  # one 16384-token launch
    acc.update(ya)
    acc.update(yb)
    two = linalg.CovarianceAccumulator(d, dev)  # two 8192-token launches
    two.update(ya)
    two.update(yb)
    assert acc.launches == 1 and two.launches == 2
    c1 = acc.finalize(False, 0.0)
    c2 = two.finalize(False, 0.0)
    assert torch.equal(c1, c1.T)
    tr_ref = ((ya.double() ** 2).sum() + (yb.double() ** 2).sum()) / n / 2
    assert abs(float(torch.trace(c1.double()) - tr_ref)) <= COV_TOL * float(tr_ref)
    v = torch.randn(d, 4, generator=g, device=dev, dtype=torch.float64)
    ref = (ya.double().T @ (ya.double() @ v) + yb.double().T @ (yb.double() @ v)) / n / 2
    got = c1.double() @ v
    assert float((got - ref).norm() / ref.norm()) <= COV_TOL
    assert float((c1 - c2).norm() / c2.norm()) <= 2e-6


def test_full_size_decode_forward_70b_shapes(dev):
    """Decode-size forward at a Llama-3-70B MLP shape (in 8192, out 28672, k 3584, N = 16 and 128):
    linearity in x and agreement with an fp32 torch product that rounds the intermediate to bf16."""
    from ptdeco_b200 import linalg
    in_f, out_f, k = 8192, 28672, 3584
    g = torch.Generator(device=dev).manual_seed(123)
    w1 = (torch.randn(k, in_f, generator=g, device=dev) / in_f ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(out_f, k, generator=g, device=dev) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(out_f, generator=g, device=dev)
    for n in (16, 128):
        x = torch.randn(n, in_f, generator=g, device=dev).to(torch.bfloat16)
        y = linalg.lowrank_forward(x, w1, w2, b).float()
        h = (x.float() @ w1.float().T).to(torch.bfloat16).float()
        ref = h @ w2.float().T + b
        assert float((y - ref).abs().max() / ref.abs().max()) <= 1e-2
        y2 = linalg.lowrank_forward((2 * x), w1, w2, None).float()  # exact in bf16: doubling x
        y1 = linalg.lowrank_forward(x, w1, w2, None).float()
        assert float((y2 - 2 * y1).abs().max() / y2.abs().max()) <= 1e-2


def test_config_round_trip_on_a_torchvision_model(dev):
    """The reference's tests/test_config_torchvision_timm.py::check_config with falor producing the
    config (lockd, which that test uses, is out of scope): decompose model 1, rebuild the architecture
    of an untouched model 2 from the returned decompose_config, load model 1's state dict strictly,
    and both must give the same output."""
    import torchvision

    import ptdeco_b200.falor as falor
    from ptdeco_b200 import modules, utils
    from synth import streams
    torch.manual_seed(271828)
    model1 = torchvision.models.get_model("convnext_tiny", weights=None, num_classes=10).to(dev).eval()
    model2 = torchvision.models.get_model("convnext_tiny", weights=None, num_classes=10).to(dev).eval()
    data = streams.IndexedStream(lambda i: torch.rand(4, 3, 64, 64, generator=torch.Generator().manual_seed(1314159 + i)))
    dc = falor.decompose_in_place(
        module=model1, device=dev, data_iterator=data, proportion_threshold=10.0, nsr_final_threshold=0.05,
        kl_final_threshold=0.05, num_data_steps=2, num_metric_steps=1, use_float64=False, use_mean=False,
        use_damping=True)
    assert len(dc) >= 5
    json.loads(json.dumps(dc))  # JSON-serialisable like the reference's
    sd = model1.state_dict()
    x = torch.rand(5, 3, 64, 64, generator=torch.Generator().manual_seed(5)).to(dev)
    with torch.no_grad():
        y1 = model1(x)
    utils.apply_decompose_config_in_place(model2, dc)
    model2.load_state_dict(sd, strict=True)
    model2.to(dev).eval()
    with torch.no_grad():
        y2 = model2(x)
    # model 1 holds LowRankSequential modules (fused kernel, fp32 operands through the bf16x3 split),
    # model 2 plain nn.Sequential pairs on cuBLAS: same weights, fp32 rounding apart
    torch.testing.assert_close(y1, y2, rtol=1e-4, atol=1e-4)
    assert modules.fuse_decomposed_modules_in_place(model2) == len(dc)
    with torch.no_grad():
        torch.testing.assert_close(model2(x), y1, rtol=1e-4, atol=1e-4)
