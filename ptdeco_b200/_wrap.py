"""Capture wrappers shared by falor and dwain (reference: F:29-153, D:19-144).

A wrapper stands in for the target layer while it is analysed: it remembers the last input (the
reference's contract, `get_last_input`), optionally the last output (what the covariance kernel
consumes directly, so the projection y = x W^T is not recomputed outside the model), exposes the
2-D weight, and builds the two-factor replacement module.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Optional

import torch


class WrappedModule(torch.nn.Module):
    """Method table of the reference's WrappedFALORModule / WrappedDWAINModule (F:29-48, D:19-38)."""

    def __init__(self) -> None:
        super().__init__()
        self.input = torch.zeros(size=(0,))
        self.output: Optional[torch.Tensor] = None
        self.capture_output = False
        # paired rank trial (see PairState): when set, the first half of the batch goes through
        # this [out, in] weight and the second half through the layer's own
        self.pair_weight: Optional[torch.Tensor] = None

    def get_weight_copy(self) -> torch.Tensor:
        raise NotImplementedError()

    def set_weight(self, weights: torch.Tensor) -> None:
        raise NotImplementedError()

    def get_last_input(self) -> torch.Tensor:
        raise NotImplementedError()

    def get_orig_module(self) -> torch.nn.Module:
        raise NotImplementedError()

    def get_decomposed_module(self, u: torch.Tensor, v: torch.Tensor) -> torch.nn.Module:
        raise NotImplementedError()

    # --- additions over the reference's table -----------------------------------------------
    def get_bias(self) -> Optional[torch.Tensor]:
        return self.get_orig_module().bias

    def get_last_output_rows(self) -> torch.Tensor:
        """Last layer output as [N, out] rows (bias still included)."""
        raise NotImplementedError()


class WrappedLinear(WrappedModule):
    def __init__(self, lin_orig: torch.nn.Module, name: Optional[str] = None):
        super().__init__()
        assert isinstance(lin_orig, torch.nn.Linear)
        self.lin_orig = lin_orig
        self.name = name

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.input = x
        if self.pair_weight is not None:
            b = x.shape[0] // 2
            bias = self.lin_orig.bias
            return torch.cat([torch.nn.functional.linear(x[:b], self.pair_weight, bias),
                              torch.nn.functional.linear(x[b:], self.lin_orig.weight, bias)], 0)
        y = self.lin_orig(x)
        if self.capture_output:
            self.output = y
        return y

    def get_weight_copy(self) -> torch.Tensor:
        return self.lin_orig.weight.detach().clone()

    def set_weight(self, weights: torch.Tensor) -> None:
        self.lin_orig.weight.copy_(weights)

    def get_last_input(self) -> torch.Tensor:
        return self.input.reshape(-1, self.lin_orig.in_features)

    def get_last_output_rows(self) -> torch.Tensor:
        assert self.output is not None, "no forward pass captured"
        return self.output.reshape(-1, self.lin_orig.out_features)

    def get_orig_module(self) -> torch.nn.Module:
        return self.lin_orig

    def get_decomposed_module(self, u: torch.Tensor, v: torch.Tensor) -> torch.nn.Module:
        """u = W1 [k, in], v = W2 [out, k] -> Sequential(Linear(in->k, no bias), Linear(k->out, bias))
        (F:76-95, D:66-85); the second layer carries the original bias."""
        use_bias = self.lin_orig.bias is not None
        k = u.shape[0]
        lin_1 = torch.nn.Linear(self.lin_orig.in_features, k, bias=False)
        lin_2 = torch.nn.Linear(k, self.lin_orig.out_features, bias=use_bias)
        lin_1.weight.data = u.detach().contiguous()
        lin_2.weight.data = v.detach().contiguous()
        if use_bias:
            lin_2.bias.data = self.lin_orig.bias.detach().clone()
        return torch.nn.Sequential(lin_1, lin_2)


class WrappedConv2d1x1(WrappedModule):
    def __init__(self, conv_orig: torch.nn.Module, name: Optional[str] = None):
        super().__init__()
        assert (isinstance(conv_orig, torch.nn.Conv2d) and conv_orig.kernel_size[0] == 1
                and conv_orig.kernel_size[1] == 1 and conv_orig.groups == 1)
        self.conv_orig = conv_orig
        self.name = name

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.input = x
        if self.pair_weight is not None:
            b = x.shape[0] // 2
            conv = self.conv_orig
            return torch.cat([conv._conv_forward(x[:b], self.pair_weight[:, :, None, None], conv.bias),
                              conv._conv_forward(x[b:], conv.weight, conv.bias)], 0)
        y = self.conv_orig(x)
        if self.capture_output:
            self.output = y
        return y

    def get_weight_copy(self) -> torch.Tensor:
        return self.conv_orig.weight.detach().data[..., 0, 0].clone()

    def set_weight(self, weights: torch.Tensor) -> None:
        self.conv_orig.weight.copy_(weights[:, :, None, None])

    def get_last_input(self) -> torch.Tensor:
        return self.input.permute(0, 2, 3, 1).reshape(-1, self.conv_orig.in_channels)

    def get_last_output_rows(self) -> torch.Tensor:
        assert self.output is not None, "no forward pass captured"
        return self.output.permute(0, 2, 3, 1).reshape(-1, self.conv_orig.out_channels)

    def get_orig_module(self) -> torch.nn.Module:
        return self.conv_orig

    def get_decomposed_module(self, u: torch.Tensor, v: torch.Tensor) -> torch.nn.Module:
        """Two 1x1 convs built with default stride/padding/dilation, exactly like the reference
        (F:128-153, D:118-144) -- a strided 1x1 target therefore changes shape, as it does there."""
        use_bias = self.conv_orig.bias is not None
        k = u.shape[0]
        conv_1 = torch.nn.Conv2d(self.conv_orig.in_channels, k, kernel_size=1, bias=False)
        conv_2 = torch.nn.Conv2d(k, self.conv_orig.out_channels, kernel_size=1, bias=use_bias)
        conv_1.weight.data = u.detach()[:, :, None, None].contiguous()
        conv_2.weight.data = v.detach()[:, :, None, None].contiguous()
        if use_bias:
            conv_2.bias.data = self.conv_orig.bias.detach().clone()
        return torch.nn.Sequential(conv_1, conv_2)


def is_decomposeable_module(module: torch.nn.Module) -> bool:
    """F:402-408 / D:540-546: any nn.Linear (subclasses included) or a 1x1, groups=1 Conv2d."""
    return isinstance(module, torch.nn.Linear) or (
        isinstance(module, torch.nn.Conv2d) and module.kernel_size[0] == 1
        and module.kernel_size[1] == 1 and module.groups == 1)


def is_num_params_reduced(proportion: float, in_features: int, out_features: int) -> bool:
    """F:273-281 / D:569-577."""
    baseline = in_features * out_features
    original_rank = min(in_features, out_features)
    proposed = (in_features + out_features) * proportion * original_rank
    return proposed < baseline


class PairState:
    """Paired rank trials. The reference evaluates a trial with two full
    forwards of the same batch, one with the decomposed weight copied into the layer and one with
    the original (F:211-233, D:247-278). The library owns the wrapped layer, so both variants can
    share ONE forward of the batch concatenated with itself: the wrapper sends the first half
    through the decomposed weight and the second half through the original. Same arithmetic per
    sample for any model that treats batch elements independently in eval mode; half the kernel
    launches (small models are launch-bound) and better-filled GEMMs.

    Safety: the first trial batch of a decompose call is evaluated BOTH ways; pairing is kept only
    if the two agree (relative Frobenius error of the logits within the dtype's rounding noise)
    and the paired forward is measurably faster, so a model with a batch-dependent forward, an
    unusual batch layout, or a forward that already fills the GPU (an LLM at 2048 tokens: measured
    slower paired) keeps the reference's two-forward path. PTDECO_B200_PAIRED_TRIALS=0 disables it."""

    def __init__(self) -> None:
        self.mode = "off" if os.environ.get("PTDECO_B200_PAIRED_TRIALS", "1") == "0" else "unknown"
        self.paired_forwards = 0
        self.probe: Optional[dict] = None  # what the verification measured

    @staticmethod
    def _double(inputs: Any) -> tuple[Optional[Any], int]:
        if isinstance(inputs, torch.Tensor):
            if inputs.dim() < 1 or inputs.shape[0] == 0:
                return None, 0
            return torch.cat([inputs, inputs], 0), inputs.shape[0]
        if isinstance(inputs, dict):
            bs = {v.shape[0] for v in inputs.values() if isinstance(v, torch.Tensor) and v.dim() >= 1}
            if len(bs) != 1 or any(isinstance(v, torch.Tensor) and v.dim() < 1 for v in inputs.values()):
                return None, 0
            return {k: (torch.cat([v, v], 0) if isinstance(v, torch.Tensor) else v)
                    for k, v in inputs.items()}, bs.pop()
        return None, 0

    def _paired(self, forward_fn, wrapper: WrappedModule, inputs: Any,
                deco_weight: torch.Tensor) -> Optional[tuple[torch.Tensor, torch.Tensor]]:
        doubled, b = self._double(inputs)
        if doubled is None:
            return None
        wrapper.pair_weight = deco_weight
        try:
            yy = forward_fn(doubled)
        finally:
            wrapper.pair_weight = None
        if not isinstance(yy, torch.Tensor) or yy.dim() < 1 or yy.shape[0] != 2 * b:
            return None
        self.paired_forwards += 1
        return yy[:b], yy[b:]

    def forward_pair(self, forward_fn, wrapper: WrappedModule, inputs: Any, orig_weight: torch.Tensor,
                     deco_weight: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """(y_deco, y_orig) of one trial batch. Leaves the original weight in the layer."""
        if self.mode == "on":
            out = self._paired(forward_fn, wrapper, inputs, deco_weight)
            if out is not None:
                return out
            self.mode = "off"
        probing = self.mode == "unknown" and deco_weight.is_cuda
        if probing:
            torch.cuda.synchronize(deco_weight.device)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
        wrapper.set_weight(deco_weight)
        y_deco = forward_fn(inputs)
        wrapper.set_weight(orig_weight)
        y_orig = forward_fn(inputs)
        if probing:
            # First trial batch of the call: evaluate it the paired way too (once to warm the
            # allocator up for the doubled shapes, once timed). Pairing is kept only if the results
            # agree AND the paired forward is faster -- it halves kernel launches, which pays for
            # launch-bound models and not for ones whose forward already fills the GPU.
            e[1].record()
            self.mode = "off"
            out = self._paired(forward_fn, wrapper, inputs, deco_weight)
            if out is not None and isinstance(y_orig, torch.Tensor) and out[0].shape == y_deco.shape:
                e[2].record()
                out = self._paired(forward_fn, wrapper, inputs, deco_weight)
                e[3].record()
                tol = 1e-4 if y_orig.dtype in (torch.float32, torch.float64) else 3e-2
                ref = torch.linalg.vector_norm(y_orig.float()).clamp_min(1e-30)
                err = max(float(torch.linalg.vector_norm(out[0].float() - y_deco.float()) / ref),
                          float(torch.linalg.vector_norm(out[1].float() - y_orig.float()) / ref))
                self.probe = {"rel_err": err, "two_forwards_ms": e[0].elapsed_time(e[1]),
                              "paired_ms": e[2].elapsed_time(e[3])}
                # NaN fails the comparisons and keeps pairing off
                if err <= tol and self.probe["paired_ms"] < 0.9 * self.probe["two_forwards_ms"]:
                    self.mode = "on"
                logging.getLogger("ptdeco.utils.common").info(
                    f"paired rank trials {self.mode}: rel_err={err:.2e} "
                    f"two_forwards={self.probe['two_forwards_ms']:.2f} ms paired={self.probe['paired_ms']:.2f} ms")
            self.paired_forwards = 0
        elif self.mode == "unknown":
            self.mode = "off"
        return y_deco, y_orig
