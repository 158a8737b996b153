"""CPU tests of the host logic: config schema, wrappers, rank arithmetic, the C-ABI surface, the
no-CPU-fallback rule, and the world_size-2 gloo path of ptdeco_b200.parallel."""
import collections
import ctypes
import json
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol():
    from ptdeco_b200 import _native as nat
    header = open(os.path.join(ROOT, "include", "ptdeco_b200.h")).read()
    declared = set(re.findall(r"\b(ptdeco_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    L = nat.lib()
    assert L.ptdeco_version() >= 100
    assert L.ptdeco_strerror(-22).decode().startswith("invalid argument")
    # pure size queries need no GPU
    assert L.ptdeco_syrk_workspace_bytes(nat.BF16, 2048, 4096) == 2 * 2048 * 4096
    assert L.ptdeco_syrk_workspace_bytes(nat.F32, 100, 10) >= 3 * 2 * 100 * 16
    assert L.ptdeco_eigh_workspace_bytes(64, 64) > 0
    assert L.ptdeco_eigh_workspace_bytes(4096, 2048) > 4096 * 4096 * 4
    assert L.ptdeco_eigh_workspace_bytes(10, 11) == 0


def test_no_cpu_fallback():
    from ptdeco_b200 import _native as nat
    from ptdeco_b200 import linalg, utils
    import ptdeco_b200.falor as falor
    with pytest.raises(nat.NativeError):
        linalg.CovarianceAccumulator(8, torch.device("cpu"))
    with pytest.raises(nat.NativeError):
        linalg.eigh(torch.eye(4))
    with pytest.raises(nat.NativeError):
        utils.calc_kl_loss(torch.zeros(2, 3), torch.zeros(2, 3))
    with pytest.raises(nat.NativeError):
        falor.decompose_in_place(module=torch.nn.Linear(4, 4), device=torch.device("cpu"),
                                 data_iterator=iter([]), proportion_threshold=0.9,
                                 nsr_final_threshold=0.1, kl_final_threshold=0.1, num_data_steps=1,
                                 num_metric_steps=1, use_float64=False, use_mean=False, use_damping=True)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ptdeco_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert not re.search(r"^\s*(from|import)\s+synth\b", src, re.M), f


def test_public_api_names():
    import ptdeco_b200
    import ptdeco_b200.falor
    assert hasattr(ptdeco_b200, "dwain") and hasattr(ptdeco_b200, "utils")
    for name in ("decompose_in_place", "is_decomposeable_module"):
        assert hasattr(ptdeco_b200.falor, name) and hasattr(ptdeco_b200.dwain, name)
    import ptdeco_b200.dwain.decomposition as D
    import ptdeco_b200.falor.decomposition as F
    for name in ("_wrap_in_place", "_unwrap_in_place", "_compute_decompositon_of_covariance_matrix",
                 "WrappedFALORLinear", "WrappedFALORConv2d1x1", "_process_module", "_compute_metrics"):
        assert hasattr(F, name), name
    for name in ("_wrap_in_place", "_unwrap_in_place", "_compute_covariance_matrix_decomposition",
                 "CovarianceComputingLinearModule", "_get_params_for_proportion", "_is_num_params_reduced",
                 "_precompute_covariance_matrix_decompositions_in_splits", "WrappedDWAINLinear"):
        assert hasattr(D, name), name
    for name in ("apply_decompose_config_in_place", "get_module_config", "build_module_from_config",
                 "MODCONFIG_META_KEY", "replace_submodule_in_place", "calc_per_channel_noise_to_signal_ratio",
                 "calc_kl_divergence", "calc_kl_loss", "to_device", "get_num_params", "free_gpu_reserved_memory"):
        assert hasattr(ptdeco_b200.utils, name), name
    import inspect
    sig = inspect.signature(ptdeco_b200.dwain.decompose_in_place)
    assert sig.parameters["min_rank"].default == 32 and sig.parameters["reduction_factor"].default == 0.5
    assert sig.parameters["decompose_in_float64"].default is True
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in sig.parameters.values())


def test_is_decomposeable_and_wrapping():
    import ptdeco_b200.falor.decomposition as F
    assert F.is_decomposeable_module(torch.nn.Linear(3, 4))
    assert F.is_decomposeable_module(torch.nn.Conv2d(3, 4, 1))
    assert F.is_decomposeable_module(torch.nn.Conv2d(3, 4, 1, stride=2))  # accepted, like the reference
    assert not F.is_decomposeable_module(torch.nn.Conv2d(4, 4, 1, groups=2))
    assert not F.is_decomposeable_module(torch.nn.Conv2d(3, 4, 3))
    assert F.is_decomposeable_module(torch.nn.modules.linear.NonDynamicallyQuantizableLinear(3, 3))
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 3))
    F._wrap_in_place(net, "0")
    assert isinstance(net[0], F.WrappedFALORConv2d1x1)
    x = torch.randn(2, 3, 5, 5)
    y = net[0](x)
    assert tuple(net[0].get_last_input().shape) == (50, 3)
    assert torch.equal(net[0].get_last_input(), x.permute(0, 2, 3, 1).reshape(-1, 3))
    assert tuple(net[0].get_weight_copy().shape) == (8, 3)
    net[0].capture_output = True
    y = net[0](x)
    assert torch.equal(net[0].get_last_output_rows(), y.permute(0, 2, 3, 1).reshape(-1, 8))
    with torch.no_grad():
        two = net[0].get_decomposed_module(u=torch.randn(2, 3), v=torch.randn(8, 2))
    assert [tuple(p.shape) for p in two.parameters()] == [(2, 3, 1, 1), (8, 2, 1, 1), (8,)]
    assert torch.equal(two[1].bias, net[0].get_orig_module().bias) and two[0].bias is None
    F._unwrap_in_place(net, "0")
    assert isinstance(net[0], torch.nn.Conv2d)
    with pytest.raises(ValueError):
        F._wrap_in_place(net, "2")
    with pytest.raises(AttributeError):
        F._wrap_in_place(net, "nope")


def test_rank_arithmetic_matches_reference_formulas():
    import ptdeco_b200.dwain.decomposition as D
    assert D._get_params_for_proportion(1.0, 4096, 14336) == 4096 * 14336
    assert D._get_params_for_proportion(0.5, 4096, 4096) == int((8192) * 0.5 * 4096)
    assert D._get_params_for_proportion(100 / 192, 192, 576) == int((192 + 576) * (100 / 192) * 192)
    assert D._is_num_params_reduced(0.25, 64, 64) and not D._is_num_params_reduced(0.5, 64, 64)
    assert D._max_rank_consumed(4096, 14336, 0.5) == 2048
    assert D._max_rank_consumed(3, 3, 0.5) == 1


def test_modconfig_roundtrip_and_schema():
    from ptdeco_b200 import utils
    seq = torch.nn.Sequential(torch.nn.Linear(6, 3, bias=False), torch.nn.Linear(3, 5))
    cfg = utils.get_module_config(seq)
    assert cfg == {"type": "Sequential", "modules": {
        "0": {"type": "Linear", "in_features": 6, "out_features": 3, "bias": False},
        "1": {"type": "Linear", "in_features": 3, "out_features": 5, "bias": True}}}
    conv = torch.nn.Sequential(torch.nn.Conv2d(4, 2, 1, bias=False), torch.nn.Conv2d(2, 7, 1))
    ccfg = json.loads(json.dumps(utils.get_module_config(conv)))  # tuples -> lists, like the artifact
    assert list(ccfg["modules"]["0"].keys()) == ["type", "in_channels", "out_channels", "kernel_size", "bias",
                                                 "groups", "padding", "padding_mode", "stride", "dilation"]
    ccfg["__meta__"] = {"proportion": 0.5}
    rebuilt = utils.build_module_from_config(ccfg)
    assert isinstance(rebuilt[1], torch.nn.Conv2d) and rebuilt[1].out_channels == 7
    model = torch.nn.ModuleDict({"a": torch.nn.Linear(6, 5), "b": torch.nn.Conv2d(4, 7, 1)})
    utils.apply_decompose_config_in_place(model, {"a": cfg, "b": ccfg})
    assert isinstance(model["a"], torch.nn.Sequential) and isinstance(model["b"], torch.nn.Sequential)
    model["a"].load_state_dict(seq.state_dict(), strict=True)
    named = torch.nn.Sequential()
    named.add_module("first", torch.nn.Linear(2, 2))
    assert list(utils.build_module_from_config(utils.get_module_config(named))._modules) == ["first"]
    with pytest.raises(ValueError):
        utils.build_module_from_config({"type": "GRU"})
    with pytest.raises(ValueError):
        utils.get_module_config(torch.nn.ReLU())
    with pytest.raises(ValueError):
        utils.to_device([1, 2], torch.device("cpu"))
    tied = torch.nn.Linear(4, 4)
    holder = torch.nn.ModuleList([tied, tied])
    assert utils.get_num_params(holder) == 20


def test_shard_plan():
    from ptdeco_b200 import parallel
    # periodic job lists: round-robin would give one rank every heavy job; LPT balances them
    costs = [3.0, 1.0, 9.0, 1.0] * 8
    owners = parallel.balanced_owners(costs, 2)
    loads = [sum(c for c, o in zip(costs, owners) if o == r) for r in range(2)]
    assert max(loads) - min(loads) <= 1.0 and sorted(set(owners)) == [0, 1]
    assert parallel.balanced_owners(costs, 2) == owners  # deterministic
    assert parallel.steps_of_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((parallel.steps_of_rank(7, r, 3) for r in range(3)), [])) == list(range(7))
    assert [parallel.owner_of(i, 8) for i in (0, 7, 8, 225)] == [0, 7, 0, 1]
    assert parallel.rank_and_world(None) == (0, 1)


class _FakeAcc:
    """Stands in for linalg.CovarianceAccumulator on CPU: the partial sums come from the oracle."""

    def __init__(self, d):
        self.C = torch.zeros(d, d)
        self.colsum = None
        self.steps = 0


class _FakeShardedAcc(_FakeAcc):
    """The canonical-shard side of CovarianceAccumulator (attributes parallel.gather_shards_to uses)."""

    def __init__(self, d, shards, rank, world):
        super().__init__(d)
        self.shards, self.shard_rank, self.shard_world = shards, rank, world
        self.shard_C = [None] * shards
        self._collapsed = False

    def update(self, part, step):
        v = step % self.shards
        if self.shard_C[v] is None:
            self.shard_C[v] = torch.zeros_like(self.C)
        self.shard_C[v] += part
        self.steps += 1


def _canonical_parts(d, n, steps, layer):
    from synth import cases
    g = torch.Generator().manual_seed(77 + layer)
    # badly scaled partial products: any change of summation order shows in the low bits
    return [(cases.step_spectrum_batch(n, d, 10 * layer + i).T @ cases.step_spectrum_batch(n, d, 10 * layer + i))
            * float(10.0 ** torch.randint(-3, 4, (1,), generator=g).item()) for i in range(steps)]


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import primitives as P
    from ptdeco_b200 import parallel
    from synth import cases
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    group = parallel.default_group()
    assert parallel.rank_and_world(group) == (rank, world)
    d, n, steps, k = 24, 64, 5, 6
    results = {}
    for layer in range(3):
        acc = _FakeAcc(d)
        for i in parallel.steps_of_rank(steps, rank, world):
            y = cases.step_spectrum_batch(n, d, 10 * layer + i).numpy()
            part = np.zeros((d, d), np.float32)
            P.update_Eyyt_in_place(part, y)
            acc.C += torch.from_numpy(part)
            acc.steps += 1

        def compute():
            cov = (acc.C / acc.steps).numpy()
            return torch.from_numpy(P.top_k(P.dwain_get_eigenvectors(cov), k).astype(np.float32).copy())

        u = parallel.owner_computes(layer, group, compute, acc, (d, k))
        assert acc.steps == steps
        results[layer] = u
    # the pipelined variant (all reductions enqueued up front as lower-triangle bands, owners solve
    # while later layers are still in flight, asynchronous broadcasts) must give the same blocks
    accs, jobs = [], []
    for layer in range(3):
        acc = _FakeAcc(d)
        for i in parallel.steps_of_rank(steps, rank, world):
            part = np.zeros((d, d), np.float32)
            P.update_Eyyt_in_place(part, cases.step_spectrum_batch(n, d, 10 * layer + i).numpy())
            acc.C += torch.from_numpy(part)
            acc.steps += 1
        acc.C += torch.triu(torch.full((d, d), 1e3 * (rank + 1)), 1)  # the strict upper triangle is
        accs.append(acc)                                              # unspecified: never sent

        def compute(acc=acc):
            cov = (torch.tril(acc.C) / acc.steps).numpy()
            cov = cov + np.tril(cov, -1).T
            return torch.from_numpy(P.top_k(P.dwain_get_eigenvectors(cov), k).astype(np.float32).copy())

        jobs.append((acc, compute, (d, k)))
    piped = parallel.owners_compute_pipelined(jobs, group, total_steps=steps)
    for layer in range(3):
        assert all(a.steps == steps for a in accs)
        assert torch.equal(piped[layer], results[layer]), layer
    # canonical shards (deterministic mode): the owner adds the V shard matrices in index order,
    # which is bit for bit the sum a single process holding all V shards forms
    shards, c_steps = 4, 7
    jobs = []
    for layer in range(3):
        acc = _FakeShardedAcc(d, shards, rank, world)
        parts = _canonical_parts(d, n, c_steps, layer)
        for i in parallel.steps_of_rank(c_steps, rank, world):
            acc.update(parts[i], i)
        jobs.append((acc, (lambda acc=acc: acc.C.clone()), (d, d)))
    canon = parallel.owners_compute_pipelined(jobs, group, total_steps=c_steps, costs=[1.0, 5.0, 2.0])
    assert all(a.steps == c_steps and all(c is None for c in a.shard_C[rank::world]) for a, _, _ in jobs)
    results["canonical"] = torch.stack(canon)
    parallel.check_identical_batches({"ids": torch.arange(6).reshape(2, 3)}, group)
    with pytest.raises(RuntimeError):
        parallel.check_identical_batches(torch.full((2, 2), float(rank)), group)
    assert parallel.resolve_group(None) is None and parallel.resolve_group("world") is group
    t = parallel.mean_over_ranks(torch.tensor([float(rank)], dtype=torch.float64), group)
    assert t.item() == pytest.approx((world - 1) / 2)
    assert parallel.max_over_ranks(float(rank), torch.device("cpu")) == world - 1
    torch.save(results, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_covariance(tmp_path):
    """world_size 2 on CPU: sharded steps + reduce-to-owner + owner eigensolve + broadcast gives
    every rank the same subspace as the single-process computation."""
    from oracle import primitives as P
    from synth import cases
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    d, n, steps, k = 24, 64, 5, 6
    for layer in range(3):
        assert torch.equal(r0[layer], r1[layer])
        full = np.zeros((d, d), np.float32)
        for i in range(steps):
            P.update_Eyyt_in_place(full, cases.step_spectrum_batch(n, d, 10 * layer + i).numpy())
        u = P.top_k(P.dwain_get_eigenvectors(full / steps), k)
        assert P.min_principal_cosine(u, r0[layer].numpy()) > 0.9999
    # canonical shards: both ranks hold exactly the single-process sum ((C_0 + C_1) + C_2) + C_3
    assert torch.equal(r0["canonical"], r1["canonical"])
    shards, c_steps = 4, 7
    for layer in range(3):
        parts = _canonical_parts(d, n, c_steps, layer)
        single = _FakeShardedAcc(d, shards, 0, 1)
        for i in range(c_steps):
            single.update(parts[i], i)
        total = torch.zeros(d, d)
        for v in range(shards):
            total += single.shard_C[v]
        assert torch.equal(r0["canonical"][layer], total), layer
        plain = sum(parts[1:], parts[0])
        assert not torch.equal(plain, total)  # the order does matter for these inputs


def test_lowrank_sequential_is_a_sequential():
    """The fused module keeps the artifact contract: config type, state-dict keys, CPU behaviour."""
    from ptdeco_b200 import modules, utils
    model = torch.nn.ModuleDict({
        "a": torch.nn.Sequential(torch.nn.Linear(6, 3, bias=False), torch.nn.Linear(3, 5)),
        "b": torch.nn.Sequential(torch.nn.Conv2d(4, 2, 1, bias=False), torch.nn.Conv2d(2, 7, 1)),
        "c": torch.nn.Sequential(torch.nn.Linear(6, 3), torch.nn.ReLU()),
        "d": torch.nn.Sequential(torch.nn.Linear(6, 3), torch.nn.Linear(3, 5)),  # first has a bias
    })
    before = {k: v.clone() for k, v in model.state_dict().items()}
    cfg_before = utils.get_module_config(model["a"])
    x = torch.randn(2, 6)
    y_before = model["a"](x)
    assert modules.fuse_decomposed_modules_in_place(model) == 2
    assert isinstance(model["a"], modules.LowRankSequential) and isinstance(model["b"], modules.LowRankSequential)
    assert type(model["c"]) is torch.nn.Sequential and type(model["d"]) is torch.nn.Sequential
    assert utils.get_module_config(model["a"]) == cfg_before
    assert list(model.state_dict().keys()) == list(before.keys())
    torch.testing.assert_close(model["a"](x), y_before)  # CPU input: the reference's own path


def test_capi_argument_errors_return_codes_without_touching_the_gpu():
    """Error convention of the boundary: negative errno-style codes, no exceptions, validated before
    any CUDA call (so these run on the CPU-only box)."""
    from ptdeco_b200 import _native as nat
    L = nat.lib()
    buf = ctypes.create_string_buffer(64)  # a non-null dummy pointer; never dereferenced
    p = ctypes.addressof(buf)
    EINVAL, ENOMEM = -22, -12
    assert L.ptdeco_syrk_accumulate(None, nat.F32, 8, 4, 4, None, p, 4, None, 1.0, None, 0, None) == EINVAL
    assert L.ptdeco_syrk_accumulate(p, 7, 8, 4, 4, None, p, 4, None, 1.0, None, 0, None) == EINVAL  # dtype
    assert L.ptdeco_syrk_accumulate(p, nat.F32, 8, 4, 2, None, p, 4, None, 1.0, None, 0, None) == EINVAL  # ldy < d
    assert L.ptdeco_syrk_accumulate(p, nat.F32, 0, 4, 4, None, p, 4, None, 1.0, None, 0, None) == 0  # empty batch
    assert L.ptdeco_syrk_accumulate(p, nat.F32, 8, 4, 4, None, p, 4, None, 1.0, None, 0, None) == ENOMEM  # no workspace
    assert L.ptdeco_cov_finalize(None, 4, 4, None, 1, 0, 0.0, None, None) == EINVAL
    assert L.ptdeco_cov_finalize(p, 4, 4, None, 0, 0, 0.0, None, None) == EINVAL  # n_steps
    assert L.ptdeco_eigh(p, 200, 200, 0, p, p, 8, None, 0, None) == EINVAL       # k out of range
    assert L.ptdeco_eigh(p, 200, 100, 10, p, p, 16, None, 0, None) == EINVAL     # lda < d
    assert L.ptdeco_eigh(p, 200, 200, 10, p, p, 16, None, 0, None) == ENOMEM     # no workspace
    assert L.ptdeco_gemm(None, 0, 0, 4, p, 0, 0, 4, 4, 4, 4, 1.0, None, p, 0, 4, 0, None, 0, None) == EINVAL
    assert L.ptdeco_gemm(p, 0, 0, 4, p, 0, 0, 4, 4, 4, 4, 1.0, None, p, nat.BF16, 4, 1, None, 0, None) == EINVAL
    assert L.ptdeco_lowrank_forward(None, 4, p, 4, p, 4, None, p, 4, nat.BF16, 4, 4, 2, 4, None, 0, None) == EINVAL
    assert L.ptdeco_lowrank_forward(p, 4, p, 4, p, 4, None, p, 4, 9, 4, 4, 2, 4, None, 0, None) == EINVAL
    assert L.ptdeco_nsr_metric(p, p, 0, 4, 4, 1e-3, None, 0, p, None) == ENOMEM
    assert L.ptdeco_kl_metric(None, p, 0, 4, 4, p, None) == EINVAL
    for code in (EINVAL, ENOMEM, -5, -34, -38, -1003):
        assert len(L.ptdeco_strerror(code)) > 3


def test_lowrank_workspace_covers_the_decode_kernel():
    """Size query only (no GPU): for bf16 batches of up to 128 tokens the workspace must hold the
    decode kernel's control words, the fp32 partial sums and the bf16 intermediate
    (4096 + 6 * round_up(n, 16) * round_up(k, 128) bytes), for larger batches the two-GEMM path's
    intermediate."""
    from ptdeco_b200 import _native as nat
    L = nat.lib()
    for n, in_f, k, out_f in ((1, 4096, 512, 4096), (16, 8192, 3584, 28672), (128, 4096, 1000, 4096)):
        npad, ldh = -(-n // 16) * 16, -(-k // 128) * 128
        assert L.ptdeco_lowrank_workspace_bytes(nat.BF16, n, in_f, k, out_f) >= 4096 + 6 * npad * ldh
    assert L.ptdeco_lowrank_workspace_bytes(nat.BF16, 8192, 4096, 512, 4096) >= 2 * 8192 * 512
    assert L.ptdeco_lowrank_workspace_bytes(nat.F32, 64, 256, 32, 256) >= 3 * 2 * 64 * 32


def test_pair_state_host_logic():
    """_wrap.PairState on the CPU: batch doubling of tensors and dict batches, refusal of batches it
    cannot double, and the reference's two-forward path (weights restored) when pairing is off."""
    from ptdeco_b200 import _wrap
    import ptdeco_b200.falor.decomposition as F
    x = torch.arange(12.0).reshape(3, 4)
    doubled, b = _wrap.PairState._double(x)
    assert b == 3 and tuple(doubled.shape) == (6, 4) and torch.equal(doubled[:3], doubled[3:])
    batch = {"input_ids": torch.ones(2, 5, dtype=torch.long), "mask": torch.ones(2, 5), "tag": "keep"}
    doubled, b = _wrap.PairState._double(batch)
    assert b == 2 and tuple(doubled["input_ids"].shape) == (4, 5) and doubled["tag"] == "keep"
    assert _wrap.PairState._double({"a": torch.ones(2, 3), "b": torch.ones(3, 3)}) == (None, 0)
    assert _wrap.PairState._double({"a": torch.ones(2, 3), "s": torch.tensor(1.0)}) == (None, 0)
    assert _wrap.PairState._double([x]) == (None, 0)

    # policy: a fixed rule, never a timing measurement: env forces, "auto" goes by size
    net = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.ReLU(), torch.nn.Linear(6, 2)).eval()
    assert _wrap.PairState(net).enabled and _wrap.PairState(net).mode == "unverified"
    assert not _wrap.PairState(None).enabled
    big = _wrap.PairState.AUTO_MAX_PARAMS
    try:
        _wrap.PairState.AUTO_MAX_PARAMS = 10
        assert not _wrap.PairState(net).enabled and _wrap.PairState(net).mode == "off"
    finally:
        _wrap.PairState.AUTO_MAX_PARAMS = big

    # the wrapper's paired forward on CPU with the two-factor op stubbed by torch (the kernel is
    # CUDA only): first half through the factors, second half through the layer, and a leading
    # dimension that is not the doubled batch is refused (seq-first layouts)
    F._wrap_in_place(net, "0")
    wrapper = net.get_submodule("0")
    w = wrapper.get_weight_copy()
    w1, w2 = 0.5 * torch.eye(4), w  # W2 W1 = 0.5 W
    wrapper._factored_forward = lambda t: torch.nn.functional.linear(
        torch.nn.functional.linear(t, wrapper.trial_factors[0]), wrapper.trial_factors[1], wrapper.get_bias())
    with torch.no_grad():
        y_orig = net(x)
        wrapper.set_trial(w1, w2)
        y_deco = net(x)
        wrapper.clear_trial()
        assert not torch.allclose(y_deco, y_orig)
        wrapper.set_trial(w1, w2, pair_batch=3)
        yy = net(torch.cat([x, x], 0))
        assert torch.allclose(yy[:3], y_deco, atol=1e-6) and torch.allclose(yy[3:], y_orig, atol=1e-6)
        with pytest.raises(_wrap.PairLayoutError):
            net(torch.cat([x, x, x], 0))
        wrapper.clear_trial()
        assert torch.equal(wrapper.get_weight_copy(), w) and wrapper.trial_factors is None
        st = _wrap.PairState(net)
        st.begin_layer()
        a, b_ = st.forward_pair(net, wrapper, x, (w1, w2))  # first batch of the layer: verified both ways
        assert st.mode == "on" and st.probe["rel_err"] <= 1e-4
        assert torch.allclose(a, y_deco, atol=1e-6) and torch.allclose(b_, y_orig, atol=1e-6)
        a, b_ = st.forward_pair(net, wrapper, x, (w1, w2))
        assert st.paired_forwards == 1 and torch.allclose(a, y_deco, atol=1e-6)
        st.begin_layer()
        assert st.mode == "unverified"  # every layer re-verifies


def test_calibration_forward_stops_after_the_target():
    """_wrap.calibration_forward: the reference discards the model output of calibration forwards
    (F:189, D:237), so the forward is cut right after the wrapped layer; the layers behind it do
    not run, the captured input / output are those of a full forward."""
    from ptdeco_b200 import _wrap
    import ptdeco_b200.dwain.decomposition as D
    calls = []

    class Tail(torch.nn.Module):
        def forward(self, t):
            calls.append(1)
            return t * 2

    net = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.ReLU(), torch.nn.Linear(6, 5), Tail()).eval()
    x = torch.randn(3, 4)
    D._wrap_in_place(net, "2")
    wrapper = net.get_submodule("2")
    wrapper.capture_output = True
    with torch.no_grad():
        full = net(x)
        assert len(calls) == 1
        want_in, want_out = wrapper.get_last_input().clone(), wrapper.get_last_output_rows().clone()
        _wrap.calibration_forward(net, x + 1, wrapper)  # first one: in full, counts the wrapper's calls
        assert len(calls) == 2 and wrapper._calls_per_forward == 1
        _wrap.calibration_forward(net, x + 1, wrapper)
        _wrap.calibration_forward(net, x, wrapper)
    assert len(calls) == 2 and not wrapper.capture_only  # the tail never ran again
    assert torch.equal(wrapper.get_last_input(), want_in) and torch.equal(wrapper.get_last_output_rows(), want_out)
    with torch.no_grad():
        assert torch.equal(net(x), full)  # and the model is untouched

    class Twice(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 4)

        def forward(self, t):
            return self.lin(self.lin(t) + 1)

    tw = Twice().eval()
    D._wrap_in_place(tw, "lin")
    wr = tw.get_submodule("lin")
    with torch.no_grad():
        _wrap.calibration_forward(tw, x, wr)
        _wrap.calibration_forward(tw, x, wr)
        last_in = wr.get_last_input().clone()
    assert wr._calls_per_forward == 2  # called twice per forward: no early exit,
    with torch.no_grad():
        assert torch.equal(last_in, tw.lin.get_orig_module()(x) + 1)  # the LAST call's input like F:63-65


def test_strided_conv_wrapper_uses_input_positions():
    """Strided / padded 1x1 convs: the reference's covariance runs over ALL input
    positions (F:125-126,159) and the rebuilt module drops stride / padding (F:136-147). The wrapper
    says so (the driver then forms y = x W^T from the input instead of hooking the layer output)."""
    from ptdeco_b200 import _wrap
    plain = _wrap.WrappedConv2d1x1(torch.nn.Conv2d(4, 8, 1), "a")
    strided = _wrap.WrappedConv2d1x1(torch.nn.Conv2d(4, 8, 1, stride=2), "b")
    padded = _wrap.WrappedConv2d1x1(torch.nn.Conv2d(4, 8, 1, padding=1), "c")
    assert plain.output_covers_input_positions()
    assert not strided.output_covers_input_positions() and not padded.output_covers_input_positions()
    x = torch.randn(2, 4, 6, 6)
    strided(x)
    assert tuple(strided.get_last_input().shape) == (2 * 6 * 6, 4)  # every input position
    new = strided.get_decomposed_module(torch.randn(3, 4), torch.randn(8, 3))
    assert new[0].stride == (1, 1) and tuple(new(x).shape) == (2, 8, 6, 6)  # the reference's shape change


def test_memory_pressure_gate_without_cuda_is_a_no_op():
    from ptdeco_b200 import utils
    if not torch.cuda.is_available():
        assert utils.relieve_gpu_memory_pressure() is False
    assert "relieve_gpu_memory_pressure" in utils.__all__ and "free_gpu_reserved_memory" in utils.__all__


def test_artifact_round_trip_like_the_reference_readme(tmp_path):
    """The reference's save / load recipe (README.md:56-105): decompose_config.json +
    decompose_state_dict.pt written from a decomposed model, then a FRESH original model is
    rebuilt, the config applied, the state dict loaded strictly -- and the fused fast-path modules
    can be swapped in on either side without changing config or state-dict keys."""
    import ptdeco_b200 as ptdeco
    from ptdeco_b200 import modules

    def build():
        torch.manual_seed(11)
        return torch.nn.Sequential(collections.OrderedDict(
            stem=torch.nn.Conv2d(3, 8, 1), act=torch.nn.ReLU(), pool=torch.nn.AdaptiveAvgPool2d(1),
            flat=torch.nn.Flatten(), fc1=torch.nn.Linear(8, 12), act2=torch.nn.ReLU(),
            head=torch.nn.Linear(12, 5)))

    # a decomposed model: two targets replaced by the two-factor modules the wrappers build
    import ptdeco_b200.falor.decomposition as F
    model = build().eval()
    decompose_config = {}
    for name, k in (("stem", 2), ("fc1", 3)):
        F._wrap_in_place(model, name)
        wrapper = model.get_submodule(name)
        w = wrapper.get_weight_copy()
        out_f, in_f = w.shape
        q = torch.linalg.qr(torch.randn(out_f, out_f, generator=torch.Generator().manual_seed(k)))[0]
        uk = q[:, :k].contiguous()
        new = wrapper.get_decomposed_module(u=uk.T @ w, v=uk)
        F._unwrap_in_place(model, name)
        ptdeco.utils.replace_submodule_in_place(model, name, new)
        cfg = ptdeco.utils.get_module_config(new)
        cfg[ptdeco.utils.MODCONFIG_META_KEY] = {"proportion": k / min(in_f, out_f), "nsr_final": 0.0, "kl_final": 0.0}
        decompose_config[name] = cfg
    x = torch.randn(4, 3, 6, 6, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        y = model(x)
    with open(tmp_path / "decompose_config.json", "wt") as f:
        json.dump(decompose_config, f)
    torch.save(model.state_dict(), tmp_path / "decompose_state_dict.pt")

    # loading side, exactly the README's steps
    fresh = build().eval()
    with open(tmp_path / "decompose_config.json", "rt") as f:
        loaded_cfg = json.load(f)
    ptdeco.utils.apply_decompose_config_in_place(fresh, loaded_cfg)
    fresh.load_state_dict(torch.load(tmp_path / "decompose_state_dict.pt"), strict=True)
    assert modules.fuse_decomposed_modules_in_place(fresh) == 2
    assert list(fresh.state_dict().keys()) == list(model.state_dict().keys())
    assert ptdeco.utils.get_module_config(fresh.get_submodule("fc1")) == {
        k: v for k, v in loaded_cfg["fc1"].items() if k != ptdeco.utils.MODCONFIG_META_KEY}
    with torch.no_grad():
        torch.testing.assert_close(fresh(x), y)
    # and a checkpoint written from the fused model loads into plain reference-style modules
    torch.save(fresh.state_dict(), tmp_path / "fused_state_dict.pt")
    plain = build().eval()
    ptdeco.utils.apply_decompose_config_in_place(plain, loaded_cfg)
    plain.load_state_dict(torch.load(tmp_path / "fused_state_dict.pt"), strict=True)
    with torch.no_grad():
        torch.testing.assert_close(plain(x), y)


def test_covariance_units_share_input_accumulators(monkeypatch):
    """dwain precompute: targets that read the SAME tensor (q/k/v, gate/up) share one input-side
    accumulator, decided on the first forward and checked on every later one; a module called
    twice per forward switches the split back to private accumulators. Host logic only: the
    accumulator and the layer GEMM are stubbed with torch-CPU doubles."""
    import ptdeco_b200.dwain.decomposition as D
    from ptdeco_b200 import linalg

    class FakeAcc:
        def __init__(self, d, device, with_mean=False, defer_rows=0, shards=1, rank=0, world=1):
            self.d, self.C, self.steps = d, torch.zeros(d, d, dtype=torch.float64), 0

        def update(self, y, sub=None, step=None):
            y = y.double()
            self.C += y.T @ y / y.shape[0]
            self.steps += 1

    monkeypatch.setattr(linalg, "CovarianceAccumulator", FakeAcc)
    monkeypatch.setattr(linalg, "linear_nt", lambda rows, w: rows @ w.T)

    class Block(torch.nn.Module):
        def __init__(self, twice=False):
            super().__init__()
            self.q, self.k, self.v = torch.nn.Linear(16, 16), torch.nn.Linear(16, 4), torch.nn.Linear(16, 4)
            self.gate, self.up, self.down = torch.nn.Linear(16, 48), torch.nn.Linear(16, 48), torch.nn.Linear(48, 16)
            self.twice = twice

        def forward(self, x):
            a = self.q(x) + torch.cat([self.k(x), self.v(x)] * 2, -1)
            h = a * 0.5  # a NEW tensor: gate / up share it, q / k / v do not
            out = self.down(torch.relu(self.gate(h)) * self.up(h))
            return self.down(torch.relu(out).repeat(1, 3)) if self.twice else out

    torch.manual_seed(0)
    names = ["q", "k", "v", "gate", "up", "down"]
    xs = [torch.randn(5, 16) for _ in range(3)]
    net = Block().eval()
    ref = [net(x) for x in xs]
    originals = D._install_covariance_modules(net, names, True, reduction_factor=0.5)
    units = net.q.units
    with torch.no_grad():
        for x, r in zip(xs, ref):
            assert torch.allclose(net(x), r, atol=1e-6)  # the wrapped forward is still the layer forward
    units.finish_probe()
    kinds = [(u.kind, [n for n in names if getattr(net, n) in u.members]) for u in units.units]
    assert kinds == [("input", ["q", "k", "v"]), ("input", ["gate", "up"]), ("output", ["down"])]
    assert all(u.acc.steps == 3 for u in units.units)
    s_ref = sum(x.double().T @ x.double() / 5 for x in xs)
    assert torch.allclose(units.units[0].acc.C, s_ref)
    assert net.k.acc is net.q.acc and net.up.acc is net.gate.acc and net.down.acc.d == 16
    D._restore_modules(net, originals)
    assert isinstance(net.q, torch.nn.Linear)

    # weight sharing inside the model: `down` runs twice per forward -> no input sharing at all,
    # and like the reference (D:204) every call counts as a step
    net = Block(twice=True).eval()
    D._install_covariance_modules(net, names, True, reduction_factor=0.5)
    with torch.no_grad():
        for x in xs:
            net(x)
    net.q.units.finish_probe()
    assert all(len(u.members) == 1 for u in net.q.units.units) and len(net.q.units.units) == 6
    assert net.down.acc.steps == 6 and net.q.acc.steps == 3

    # PTDECO_B200_SHARE_INPUTS=0: the reference's one-accumulator-per-target layout
    monkeypatch.setenv("PTDECO_B200_SHARE_INPUTS", "0")
    net = Block().eval()
    D._install_covariance_modules(net, names, True, reduction_factor=0.5)
    with torch.no_grad():
        for x in xs:
            net(x)
    net.q.units.finish_probe()
    assert len(net.q.units.units) == 6 and net.k.acc is not net.q.acc
