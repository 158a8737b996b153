"""Random-init synthetic models of the shapes BASELINE.json names (no checkpoints, no network).

These are *user models* from the library's point of view: plain torch modules whose Linear / 1x1
Conv2d submodules are the decomposition targets.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _seeded_init(module: nn.Module, seed: int) -> None:
    g = torch.Generator()
    g.manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            nn.init.kaiming_uniform_(m.weight, a=5 ** 0.5, generator=g)
            if m.bias is not None:
                fan_in, _ = nn.init._calculate_fan_in_and_fan_out(m.weight)
                bound = fan_in ** -0.5 if fan_in > 0 else 0
                nn.init.uniform_(m.bias, -bound, bound, generator=g)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, std=0.5, generator=g)


# ---------------------------------------------------------------------------------------------
# The two nets of the reference's primitive tests (tests/test_deco_primitives_falor.py:34-73):
# same layer, same init order and seed usage, so outputs are reproducible against the reference.
class PrimitiveLinearNet(nn.Module):
    def __init__(self, in_features: int, out_features: int, gen: torch.Generator, dict_input=False):
        super().__init__()
        self.mod = nn.Linear(in_features, out_features)
        self.dict_input = dict_input
        nn.init.kaiming_uniform_(self.mod.weight, a=5 ** 0.5, generator=gen)
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.mod.weight)
        bound = fan_in ** -0.5 if fan_in > 0 else 0
        nn.init.uniform_(self.mod.bias, -bound, bound)  # (the reference draws this un-seeded too)

    def forward(self, x):
        if self.dict_input:
            x = x["inp"]
        return torch.flatten(self.mod(x), start_dim=1)


class PrimitiveConv1x1Net(nn.Module):
    def __init__(self, in_features: int, out_features: int, gen: torch.Generator, dict_input=False):
        super().__init__()
        self.mod = nn.Conv2d(in_features, out_features, kernel_size=(1, 1))
        self.dict_input = dict_input
        nn.init.kaiming_uniform_(self.mod.weight, a=5 ** 0.5, generator=gen)
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.mod.weight)
        if fan_in != 0:
            bound = fan_in ** -0.5
            nn.init.uniform_(self.mod.bias, -bound, bound, generator=gen)

    def forward(self, x):
        if self.dict_input:
            x = x["inp"]
        return torch.flatten(self.mod(x), start_dim=1)


# ---------------------------------------------------------------------------------------------
class _Attention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        return self.proj(out.transpose(1, 2).reshape(b, n, c))


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, in_ch: int, dim: int, patch: int):
        super().__init__()
        self.proj = nn.Conv2d(in_ch, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class DeiTLike(nn.Module):
    """timm `VisionTransformer` module layout (deit_tiny_patch16_224 = dim 192, depth 12, heads 3):
    patch_embed.proj / blocks.{i}.attn.qkv / attn.proj / mlp.fc1 / mlp.fc2 / head, cls token."""

    def __init__(self, img: int = 224, patch: int = 16, dim: int = 192, depth: int = 12,
                 heads: int = 3, mlp_ratio: float = 4.0, num_classes: int = 10, in_ch: int = 3,
                 seed: int = 271828):
        super().__init__()
        self.patch_embed = _PatchEmbed(in_ch, dim, patch)
        n = (img // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, dim))
        self.blocks = nn.Sequential(*[_Block(dim, heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, num_classes)
        _seeded_init(self, seed)
        g = torch.Generator()
        g.manual_seed(seed + 1)
        with torch.no_grad():
            self.pos_embed.normal_(std=0.02, generator=g)
            self.cls_token.normal_(std=0.02, generator=g)

    def forward(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        x = self.norm(self.blocks(x))
        return self.head(x[:, 0])


class ConvMLPNet(nn.Module):
    """ConvNeXt-flavoured stages whose MLPs are 1x1 convs (exercises the Conv2d target path that
    torchvision's convnext_tiny, Linear-only, does not): stem 4x4/4, then per stage
    [depthwise 3x3 -> 1x1 expand -> GELU -> 1x1 project] blocks, 2x2/2 downsample between stages."""

    def __init__(self, dims=(96, 192), depths=(1, 1), expand: int = 4, num_classes: int = 10,
                 in_ch: int = 3, seed: int = 271828):
        super().__init__()
        layers = [nn.Conv2d(in_ch, dims[0], kernel_size=4, stride=4)]
        for si, (d, n) in enumerate(zip(dims, depths)):
            if si > 0:
                layers.append(nn.Conv2d(dims[si - 1], d, kernel_size=2, stride=2))
            for _ in range(n):
                layers.append(_ConvBlock(d, expand))
        self.features = nn.Sequential(*layers)
        self.head = nn.Linear(dims[-1], num_classes)
        _seeded_init(self, seed)

    def forward(self, x):
        return self.head(self.features(x).mean(dim=(2, 3)))


class _ConvBlock(nn.Module):
    def __init__(self, dim: int, expand: int):
        super().__init__()
        self.dw = nn.Conv2d(dim, dim, kernel_size=3, padding=1, groups=dim)
        self.pw1 = nn.Conv2d(dim, dim * expand, kernel_size=1)
        self.pw2 = nn.Conv2d(dim * expand, dim, kernel_size=1)

    def forward(self, x):
        return x + self.pw2(F.gelu(self.pw1(self.dw(x))))


# ---------------------------------------------------------------------------------------------
class _RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return (x.float() * torch.rsqrt(v + self.eps)).to(x.dtype) * self.weight


def _rope(x: torch.Tensor, theta: float) -> torch.Tensor:
    b, h, s, d = x.shape
    pos = torch.arange(s, device=x.device, dtype=torch.float32)
    inv = 1.0 / (theta ** (torch.arange(0, d, 2, device=x.device, dtype=torch.float32) / d))
    ang = pos[:, None] * inv[None, :]
    cos, sin = ang.cos().to(x.dtype), ang.sin().to(x.dtype)
    x1, x2 = x[..., 0::2], x[..., 1::2]
    return torch.stack([x1 * cos - x2 * sin, x1 * sin + x2 * cos], dim=-1).flatten(-2)


class _LlamaAttention(nn.Module):
    def __init__(self, hidden: int, heads: int, kv_heads: int, theta: float):
        super().__init__()
        self.heads, self.kv_heads, self.hd, self.theta = heads, kv_heads, hidden // heads, theta
        self.q_proj = nn.Linear(hidden, heads * self.hd, bias=False)
        self.k_proj = nn.Linear(hidden, kv_heads * self.hd, bias=False)
        self.v_proj = nn.Linear(hidden, kv_heads * self.hd, bias=False)
        self.o_proj = nn.Linear(heads * self.hd, hidden, bias=False)

    def forward(self, x):
        b, s, _ = x.shape
        q = self.q_proj(x).view(b, s, self.heads, self.hd).transpose(1, 2)
        k = self.k_proj(x).view(b, s, self.kv_heads, self.hd).transpose(1, 2)
        v = self.v_proj(x).view(b, s, self.kv_heads, self.hd).transpose(1, 2)
        q, k = _rope(q, self.theta), _rope(k, self.theta)
        rep = self.heads // self.kv_heads
        if rep > 1:
            k = k.repeat_interleave(rep, dim=1)
            v = v.repeat_interleave(rep, dim=1)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.o_proj(o.transpose(1, 2).reshape(b, s, -1))


class _LlamaMLP(nn.Module):
    def __init__(self, hidden: int, inter: int):
        super().__init__()
        self.gate_proj = nn.Linear(hidden, inter, bias=False)
        self.up_proj = nn.Linear(hidden, inter, bias=False)
        self.down_proj = nn.Linear(inter, hidden, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class _LlamaLayer(nn.Module):
    def __init__(self, hidden, inter, heads, kv_heads, theta):
        super().__init__()
        self.input_layernorm = _RMSNorm(hidden)
        self.self_attn = _LlamaAttention(hidden, heads, kv_heads, theta)
        self.post_attention_layernorm = _RMSNorm(hidden)
        self.mlp = _LlamaMLP(hidden, inter)

    def forward(self, x):
        x = x + self.self_attn(self.input_layernorm(x))
        return x + self.mlp(self.post_attention_layernorm(x))


class _LlamaBody(nn.Module):
    def __init__(self, vocab, hidden, inter, layers, heads, kv_heads, theta):
        super().__init__()
        self.embed_tokens = nn.Embedding(vocab, hidden)
        self.layers = nn.ModuleList(
            [_LlamaLayer(hidden, inter, heads, kv_heads, theta) for _ in range(layers)])
        self.norm = _RMSNorm(hidden)


class LlamaLikeDecoder(nn.Module):
    """HF `LlamaForCausalLM` module names (model.layers.{i}.self_attn.q_proj ... , lm_head) and
    maths (RMSNorm, RoPE, GQA, SwiGLU), dict input like the reference's LLM WrapperModule
    (examples/trainer_llm/dwain_wrapper_module.py:21-30). Llama-3-8B shape = defaults."""

    def __init__(self, vocab: int = 128256, hidden: int = 4096, inter: int = 14336,
                 layers: int = 32, heads: int = 32, kv_heads: int = 8, theta: float = 500000.0,
                 seed: int = 271828, init: bool = True):
        super().__init__()
        self.model = _LlamaBody(vocab, hidden, inter, layers, heads, kv_heads, theta)
        self.lm_head = nn.Linear(hidden, vocab, bias=False)
        if init:
            _seeded_init(self, seed)

    def forward(self, d):
        ids = d["input_ids"] if isinstance(d, dict) else d
        x = self.model.embed_tokens(ids)
        for layer in self.model.layers:
            x = layer(x)
        return self.lm_head(self.model.norm(x))


def llama_ce_loss(input_dict: dict, logits: torch.Tensor) -> torch.Tensor:
    """Mean next-token cross entropy, the reference's `ce_loss` contract
    (examples/trainer_llm/dwain_wrapper_module.py:33-46)."""
    labels = input_dict["labels"][:, 1:].reshape(-1)
    lg = logits[:, :-1, :].reshape(-1, logits.shape[-1]).float()
    return F.cross_entropy(lg, labels)


def fast_init_(module: nn.Module, seed: int, std: float = 0.02) -> None:
    """Cheap in-place random init for big (GPU-resident) models: normal(0, std) weights."""
    dev = next(module.parameters()).device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, std if p.dim() == 2 and p.shape[0] < 100000 else std, generator=g)


def sqrt_fan_in_std(m: nn.Linear) -> float:
    return 1.0 / math.sqrt(m.in_features)
