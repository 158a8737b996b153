"""Named, seeded parity cases shared by tests/golden/make_golden.py (reference run), the oracle
tests (CPU) and the GPU parity tests (product run): same models, same streams, same keyword
arguments, so that the three can be compared trial by trial.
"""
from __future__ import annotations

from typing import Any, Iterator

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import models, streams
from .streams import DATA_SEED, MODEL_SEED, IndexedStream  # noqa: F401


# ------------------------------------------------------------------ the reference's primitive tests
def primitive_case(kind: str, dict_input: bool = False):
    """tests/test_deco_primitives_{falor,dwain}.py: fin 64, fout 32, h = w = 16, bs 8, weight seed
    271828, ONE sequential data generator seeded 1314159 (the first batch is the probe input, the
    next eight feed the covariance). Returns (net, iterator)."""
    fin, fout, h, w, bs = 64, 32, 16, 16, 8
    gen = torch.Generator()
    gen.manual_seed(MODEL_SEED)
    torch.manual_seed(0)  # the reference draws the Linear bias from the global RNG
    if kind == "linear":
        net = models.PrimitiveLinearNet(fin, fout, gen, dict_input=dict_input)
        shape = (bs, h, w, fin)
    else:
        net = models.PrimitiveConv1x1Net(fin, fout, gen, dict_input=dict_input)
        shape = (bs, fin, h, w)

    def it() -> Iterator:
        g = torch.Generator()
        g.manual_seed(DATA_SEED)
        while True:
            x = torch.rand(*shape, generator=g)
            yield {"inp": x} if dict_input else x

    return net, it()


def step_spectrum_batch(n: int, d: int, index: int) -> torch.Tensor:
    return streams.step_spectrum_activations(n, d, seed=index)


# ------------------------------------------------------------------ falor cases
class TinyMLP(nn.Module):
    def __init__(self, fin=48, h1=96, h2=64, classes=10, seed=MODEL_SEED):
        super().__init__()
        self.fc1 = nn.Linear(fin, h1)
        self.fc2 = nn.Linear(h1, h2)
        self.head = nn.Linear(h2, classes)
        models._seeded_init(self, seed)

    def forward(self, x):
        return self.head(F.gelu(self.fc2(F.gelu(self.fc1(x)))))


def _lowrank_vectors(index: int, bs: int, dim: int, latent: int, noise: float) -> torch.Tensor:
    gb = torch.Generator()
    gb.manual_seed(DATA_SEED - 11)
    basis = torch.randn(latent, dim, generator=gb) / latent ** 0.5
    g = torch.Generator()
    g.manual_seed(DATA_SEED + index)
    z = torch.randn(bs, latent, generator=g) * torch.logspace(0, -1.5, latent)
    return 3.0 * z @ basis + noise * torch.randn(bs, dim, generator=g)


class EdgeNet(nn.Module):
    """Edge cases of target selection in one small net: a rank-1 Linear (skipped, F:309-317), a
    1x1 conv, a grouped 1x1 conv and a 3x3 conv (not targets), a blacklisted Linear, a layer whose
    decomposition cannot reduce parameters, and a head."""

    def __init__(self, seed=MODEL_SEED):
        super().__init__()
        self.stem = nn.Conv2d(3, 12, kernel_size=3, padding=1)
        self.pw = nn.Conv2d(12, 20, kernel_size=1)
        self.grouped = nn.Conv2d(20, 20, kernel_size=1, groups=4)
        self.gate = nn.Linear(20, 1)
        self.wide = nn.Linear(20, 36)
        self.keep = nn.Linear(36, 36)
        self.tiny = nn.Linear(36, 3)
        self.head = nn.Linear(3, 10)
        models._seeded_init(self, seed)

    def forward(self, x):
        h = F.gelu(self.grouped(F.gelu(self.pw(F.gelu(self.stem(x))))))
        h = h.mean(dim=(2, 3))
        h = h * torch.sigmoid(self.gate(h))
        h = F.gelu(self.keep(F.gelu(self.wide(h))))
        return 30.0 * self.head(self.tiny(h))  # random-init logits are tiny next to NSR epsilon 1e-3


FALOR_CASES = ("mlp", "convmlp", "deit_small", "deit_tiny", "convnext_tiny")


def falor_case(name: str):
    """Returns (model, stream, kwargs for falor.decompose_in_place except module/device/iterator)."""
    kw: dict[str, Any] = dict(blacklisted_module_names=None, proportion_threshold=0.9,
                              use_float64=True, use_mean=False, use_damping=True)
    if name == "mlp":
        model = TinyMLP()
        stream = IndexedStream(lambda i: _lowrank_vectors(i, 64, 48, 12, 0.02))
        kw.update(nsr_final_threshold=0.055, kl_final_threshold=0.02, num_data_steps=4,
                  num_metric_steps=2)
    elif name == "convmlp":
        model = models.ConvMLPNet(dims=(16, 32), depths=(1, 1), expand=4, num_classes=10)
        stream = IndexedStream(lambda i: streams.lowrank_image_batch(0, i, 8, 3, 32, 24))
        kw.update(nsr_final_threshold=0.055, kl_final_threshold=0.02, num_data_steps=4,
                  num_metric_steps=2, use_mean=True)
    elif name == "deit_small":
        model = models.DeiTLike(img=32, patch=8, dim=48, depth=2, heads=3, num_classes=10)
        stream = IndexedStream(lambda i: streams.lowrank_image_batch(1, i, 8, 3, 32, 24))
        kw.update(nsr_final_threshold=0.055, kl_final_threshold=0.02, num_data_steps=4,
                  num_metric_steps=2, blacklisted_module_names=["head"])
    elif name == "deit_tiny":
        # BASELINE.json configs[0]: deit_tiny_patch16_224 layout, random init, (5,3,224,224) inputs,
        # 10 classes (the convention of the reference's tests/test_decompose_torchvision_timm.py:28-34)
        model = models.DeiTLike(num_classes=10)
        stream = IndexedStream(lambda i: streams.image_batch(0, i, 5))
        # use_float64=True (the reference examples' setting): with fp32 LAPACK the reference's own
        # eigenvectors are noise-determined in the flat tail of these random-init spectra (its fp32
        # and fp64 runs disagree by 20 % in NSR on blocks.0.attn.proj), which makes parity ill-posed.
        kw.update(nsr_final_threshold=0.05, kl_final_threshold=0.02, num_data_steps=8,
                  num_metric_steps=2, use_float64=True)
    elif name == "edge":  # not a golden case: compared against the oracle directly
        model = EdgeNet()
        stream = IndexedStream(lambda i: streams.lowrank_image_batch(5, i, 16, 3, 8, 40, noise=0.3))
        kw.update(nsr_final_threshold=0.0005, kl_final_threshold=0.05, num_data_steps=3,
                  num_metric_steps=2, blacklisted_module_names=["keep", "nonexistent.name"],
                  proportion_threshold=0.8)
    elif name == "convnext_tiny":
        # BASELINE.json configs[1]: torchvision convnext_tiny (random init, 37 Linear targets; it has
        # no 1x1 convs), synthetic ImageNet-shape batches; layer_scale set to 1 so that the block
        # MLPs matter at random init (with the default 1e-6 every rank collapses to 2)
        import torchvision

        torch.manual_seed(MODEL_SEED)
        model = torchvision.models.convnext_tiny(weights=None)
        with torch.no_grad():
            for n_, p_ in model.named_parameters():
                if n_.endswith("layer_scale"):
                    p_.fill_(1.0)
        # 16 calibration batches of 8 images: the last stage sees 8*49*16 = 6272 rows >= d = 3072, so
        # no tested rank falls inside the rank-deficient (arbitrary-basis) part of a covariance
        stream = IndexedStream(lambda i: streams.image_batch(4, i, 8))
        kw.update(nsr_final_threshold=0.055, kl_final_threshold=0.02, num_data_steps=16,
                  num_metric_steps=1)
    else:
        raise KeyError(name)
    model.eval()
    return model, stream, kw


# ------------------------------------------------------------------ dwain cases
class DictInputNet(nn.Module):
    """dwain's model contract for vision (examples/trainer_vision/dwain_wrapper_module.py): the
    model takes the batch dict and returns logits; `raw_model` holds the actual network."""

    def __init__(self, raw_model: nn.Module):
        super().__init__()
        self.raw_model = raw_model

    def forward(self, d):
        return self.raw_model(d["inp"])


def vision_ce_loss(input_dict, logits):
    return F.cross_entropy(logits.float(), input_dict["labels"])


DWAIN_CASES = ("llama_tiny", "llama_tiny_splits", "convmlp", "llama_tiny_bf16", "llama_tiny_bf16_splits")


def dwain_case(name: str):
    """Returns (model, data stream, metric stream, kwargs). The two streams are distinct objects
    over distinct index ranges so that consumption order of each is observable."""
    if name not in DWAIN_CASES:
        raise KeyError(name)
    if name == "convmlp":  # 1x1-conv targets through dwain's per-layer path (no precompute)
        model = DictInputNet(models.ConvMLPNet(dims=(16, 32), depths=(1, 1), expand=4, num_classes=10))
        model.eval()

        def batch(stream_id):
            def make(i):
                g = torch.Generator()
                g.manual_seed(DATA_SEED + 977 * stream_id + i)
                return {"inp": streams.lowrank_image_batch(stream_id, i, 8, 3, 32, 24),
                        "labels": torch.randint(0, 10, (8,), generator=g)}
            return make

        kw = dict(num_data_steps=4, num_metric_steps=2, blacklisted_module_names=["raw_model.head"],
                  nsr_final_threshold=3.15e-5, min_rank=4, trade_off_factor=0.5, reduction_factor=0.5,
                  max_accepted_ppl_diff=0.1, decompose_in_float64=True,
                  precomputing_covariance_num_splits=None)
        return model, IndexedStream(batch(6)), IndexedStream(batch(7)), kw
    model = models.LlamaLikeDecoder(vocab=256, hidden=64, inter=176, layers=2, heads=4, kv_heads=2,
                                    theta=10000.0)
    if "bf16" in name:
        # the headline dtype (BASELINE configs[2] is a bf16 LLM): the same tiny decoder in bfloat16.
        # Every accepted trial sits >= 2.3x below and every rejected one >= 1.9x above the NSR
        # threshold in the reference's own run, far outside bf16 noise (the reference's fp32 and
        # bf16 runs of this model differ by <= 7 % in NSR), and the perplexity gates are loose.
        model = model.to(torch.bfloat16)
    model.eval()
    data = IndexedStream(lambda i: streams.token_batch(2, i, 2, 64, 256))
    metric = IndexedStream(lambda i: streams.token_batch(3, i, 2, 64, 256))
    kw: dict[str, Any] = dict(num_data_steps=4, num_metric_steps=2,
                              blacklisted_module_names=["lm_head"], nsr_final_threshold=0.004,
                              min_rank=8, trade_off_factor=0.5, reduction_factor=0.5,
                              max_accepted_ppl_diff=0.1, decompose_in_float64=True,
                              precomputing_covariance_num_splits=2 if name.endswith("splits") else None)
    if "bf16" in name:
        kw["max_accepted_ppl_diff"] = 0.5  # NSR is the governing gate
        kw["trade_off_factor"] = 50.0
    return model, data, metric, kw


def dwain_loss_fn(name: str):
    return vision_ce_loss if name == "convmlp" else models.llama_ce_loss
