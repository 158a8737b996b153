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

from . import linalg


class StopForward(BaseException):
    """Raised by a wrapper in capture-only mode once it has seen its input (and produced its
    output): the calibration forward of F:189 / D:237 discards the model output, so the layers
    AFTER the target need not run. Caught by `calibration_forward`. A BaseException so that a broad
    `except Exception` inside a user model does not swallow it."""


class PairLayoutError(RuntimeError):
    """A paired rank trial reached the wrapper with a leading dimension that is not the doubled
    batch (seq-first layouts, reshapes that fold the batch): pairing is not valid for this layer."""


class WrappedModule(torch.nn.Module):
    """Method table of the reference's WrappedFALORModule / WrappedDWAINModule (F:29-48, D:19-38)."""

    def __init__(self) -> None:
        super().__init__()
        self.input = torch.zeros(size=(0,))
        self.output: Optional[torch.Tensor] = None
        self.capture_output = False
        self.capture_only = False  # raise StopForward after capturing (calibration forwards)
        self.calls = 0             # forward calls seen (a layer used twice per model forward: no early exit)
        # rank trial (see set_trial): rows go through the two-factor op W2 (W1 x) instead of a
        # materialised W2 W1 copied into the layer (F:347-348 / D:427-429 + set_weight)
        self.trial_factors: Optional[tuple[torch.Tensor, torch.Tensor]] = None
        self.trial_pair_batch = 0  # b > 0: batch is [2b, ...], first half decomposed, second original

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

    def output_covers_input_positions(self) -> bool:
        """Whether the layer output has one row per INPUT position, i.e. the captured output can
        stand in for the reference's y = x W^T over all input rows (F:125-126,159; D:115-116,239).
        False for strided / padded 1x1 convs: there the covariance is formed from the input."""
        return True

    def set_trial(self, w1: torch.Tensor, w2: torch.Tensor, pair_batch: int = 0) -> None:
        """Evaluate the layer as W2 (W1 x) + b (w1 [k, in], w2 [out, k]) until clear_trial():
        every row, or with pair_batch = b only the first b entries of a [2b, ...] batch."""
        dt = self.get_orig_module().weight.dtype
        dt = dt if dt in (torch.float32, torch.bfloat16) else torch.float32
        self.trial_factors = (w1.to(dt).contiguous(), w2.to(dt).contiguous())
        self.trial_pair_batch = int(pair_batch)

    def clear_trial(self) -> None:
        self.trial_factors = None
        self.trial_pair_batch = 0

    def _orig_forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()

    def _factored_forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.input = x
        self.calls += 1
        if self.trial_factors is not None:
            b = self.trial_pair_batch
            if b == 0:
                return self._factored_forward(x)
            if x.dim() < 1 or x.shape[0] != 2 * b:
                raise PairLayoutError(f"leading dim {tuple(x.shape)[:1]} is not the doubled batch {2 * b}")
            return torch.cat([self._factored_forward(x[:b]), self._orig_forward(x[b:])], 0)
        y = self._orig_forward(x)
        if self.capture_output:
            self.output = y
        if self.capture_only:
            raise StopForward()
        return y


class WrappedLinear(WrappedModule):
    def __init__(self, lin_orig: torch.nn.Module, name: Optional[str] = None):
        super().__init__()
        assert isinstance(lin_orig, torch.nn.Linear)
        self.lin_orig = lin_orig
        self.name = name

    def _orig_forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.lin_orig(x)

    def _factored_forward(self, x: torch.Tensor) -> torch.Tensor:
        w1, w2 = self.trial_factors
        rows = x.reshape(-1, self.lin_orig.in_features)
        y = linalg.lowrank_forward(rows, w1, w2, self.lin_orig.bias)
        return y.to(x.dtype).reshape(*x.shape[:-1], self.lin_orig.out_features)

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
        if not self.output_covers_input_positions():
            logging.getLogger("ptdeco.utils.common").warning(
                f"{name}: 1x1 conv with stride={conv_orig.stride} padding={conv_orig.padding}: like the "
                "reference (F:125-126,136-147), the covariance uses ALL input positions and the "
                "decomposed module is rebuilt WITHOUT stride / padding (its output shape changes); "
                "blacklist such layers unless that is intended")

    def output_covers_input_positions(self) -> bool:
        c = self.conv_orig
        return tuple(c.stride) == (1, 1) and c.padding in ((0, 0), 0, "valid")

    def _orig_forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.conv_orig(x)

    def _factored_forward(self, x: torch.Tensor) -> torch.Tensor:
        w1, w2 = self.trial_factors
        conv = self.conv_orig
        if not self.output_covers_input_positions():
            # strided / padded target: keep the layer's own geometry for the trial, as the
            # reference does by copying the effective weight into the original conv (F:211-233)
            h = conv._conv_forward(x, w1[:, :, None, None].to(x.dtype), None)
            return torch.nn.functional.conv2d(h, w2[:, :, None, None].to(x.dtype), conv.bias)
        n, c, hh, ww = x.shape
        rows = x.permute(0, 2, 3, 1).reshape(-1, c)
        y = linalg.lowrank_forward(rows, w1, w2, conv.bias)
        return y.to(x.dtype).reshape(n, hh, ww, conv.out_channels).permute(0, 3, 1, 2)

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


def calibration_forward(forward_fn, inputs: Any, wrapper: WrappedModule) -> None:
    """One calibration forward of the user model (F:189 / D:237), stopped right after the wrapped
    target produced its activation: the reference discards the model output of these calls, so the
    rest of the network is dead work. The first calibration forward of a layer always runs in
    full and counts the calls of the wrapper: a layer the model calls more than once per forward
    keeps full forwards (the reference's `self.input = x` keeps the LAST call's activations,
    F:63-65, which an exit at the first call would not see). PTDECO_B200_EARLY_EXIT=0 runs every
    forward in full."""
    if os.environ.get("PTDECO_B200_EARLY_EXIT", "1") == "0" or getattr(wrapper, "_calls_per_forward", None) != 1:
        before = wrapper.calls
        forward_fn(inputs)
        wrapper._calls_per_forward = wrapper.calls - before
        return
    wrapper.capture_only = True
    try:
        forward_fn(inputs)
    except StopForward:
        pass
    finally:
        wrapper.capture_only = False


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
    through the two-factor op and the second half through the original layer. Same arithmetic per
    sample for any model that treats batch elements independently in eval mode; half the kernel
    launches (small models are launch-bound) and better-filled GEMMs.

    Safety: for EVERY layer the first trial batch is evaluated both ways and pairing is used for
    the rest of that layer's search only if the two agree (relative Frobenius error of the logits
    within the dtype's rounding noise); the wrapper additionally checks that the tensor it receives
    still has the doubled batch as its leading dimension (PairLayoutError otherwise). A model with
    a batch-dependent forward or an unusual batch layout keeps the reference's two-forward path.

    Whether pairing is attempted at all is a fixed policy, never a timing measurement:
    PTDECO_B200_PAIRED_TRIALS=1 / 0 forces it on / off; the default ("auto") pairs models of at
    most 100 M parameters (launch-bound forwards) and leaves larger ones -- whose forward already
    fills the GPU: an 8B decoder at 2048 tokens measured slower paired -- on two forwards.
    The deterministic mode (_native.set_deterministic) never pairs: a layer's verification batch is
    evaluated unpaired and the rest paired, and which batch that is depends on how the trial jobs
    are dealt over GPUs."""

    AUTO_MAX_PARAMS = 100_000_000

    def __init__(self, root_module: Optional[torch.nn.Module] = None) -> None:
        env = os.environ.get("PTDECO_B200_PAIRED_TRIALS", "auto")
        if env == "0" or (linalg.nat.call_flags() & linalg.nat.FLAG_DETERMINISTIC):
            self.enabled = False
        elif env == "1":
            self.enabled = True
        else:
            n = sum(p.numel() for p in root_module.parameters()) if root_module is not None else 0
            self.enabled = root_module is not None and n <= self.AUTO_MAX_PARAMS
        self.mode = "unverified" if self.enabled else "off"  # state for the CURRENT layer
        self.paired_forwards = 0
        self.probe: Optional[dict] = None  # what the last verification measured

    def begin_layer(self) -> None:
        """Every layer re-verifies: its input layout can differ from the previous layer's."""
        self.mode = "unverified" if self.enabled else "off"

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
                factors: tuple[torch.Tensor, torch.Tensor]) -> Optional[tuple[torch.Tensor, torch.Tensor]]:
        doubled, b = self._double(inputs)
        if doubled is None:
            return None
        wrapper.set_trial(factors[0], factors[1], pair_batch=b)
        try:
            yy = forward_fn(doubled)
        except PairLayoutError:
            return None
        finally:
            wrapper.clear_trial()
        if not isinstance(yy, torch.Tensor) or yy.dim() < 1 or yy.shape[0] != 2 * b:
            return None
        self.paired_forwards += 1
        return yy[:b], yy[b:]

    def forward_pair(self, forward_fn, wrapper: WrappedModule, inputs: Any,
                     factors: tuple[torch.Tensor, torch.Tensor]) -> tuple[torch.Tensor, torch.Tensor]:
        """(y_deco, y_orig) of one trial batch; `factors` = (W1 [k, in], W2 [out, k]). The layer's
        own weight is never touched."""
        if self.mode == "on":
            out = self._paired(forward_fn, wrapper, inputs, factors)
            if out is not None:
                return out
            self.mode = "off"
        wrapper.set_trial(factors[0], factors[1])
        try:
            y_deco = forward_fn(inputs)
        finally:
            wrapper.clear_trial()
        y_orig = forward_fn(inputs)
        if self.mode == "unverified":
            # first trial batch of this layer: evaluate it the paired way too and compare
            self.mode = "off"
            out = self._paired(forward_fn, wrapper, inputs, factors)
            if out is not None and isinstance(y_orig, torch.Tensor) and out[0].shape == y_deco.shape:
                tol = 1e-4 if y_orig.dtype in (torch.float32, torch.float64) else 3e-2
                ref = torch.linalg.vector_norm(y_orig.float()).clamp_min(1e-30)
                err = max(float(torch.linalg.vector_norm(out[0].float() - y_deco.float()) / ref),
                          float(torch.linalg.vector_norm(out[1].float() - y_orig.float()) / ref))
                self.probe = {"rel_err": err, "layer": getattr(wrapper, "name", None)}
                if err <= tol:  # NaN fails the comparison and keeps pairing off
                    self.mode = "on"
                logging.getLogger("ptdeco.utils.common").debug(
                    f"paired rank trials {self.mode} for {self.probe['layer']}: rel_err={err:.2e}")
            self.paired_forwards = 0
        return y_deco, y_orig
