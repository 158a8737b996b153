"""CUDA-graph replay of the user model's forward during falor's rank search.

falor calls `root_module(x)` thousands of times on small vision models (DeiT-tiny: ~1750 forwards of
~150 kernels each); on a B200 those forwards are launch-bound, not compute-bound. The model is a
black box, but between two calls only parameter VALUES change (`set_weight` copies in place), so
after two eager warm-up calls the forward is captured once per input shape and replayed. Replay
runs exactly the kernels eager mode ran, so results are unchanged. Anything that cannot be captured
(host syncs or data-dependent control flow inside the user model, non-tensor inputs/outputs) makes
this object fall back to eager calls for good.

OFF by default (PTDECO_B200_CUDA_GRAPHS=1 enables it): measured on a B200 box, eager forwards of
DeiT-tiny / ConvNeXt-tiny already take ~1.5-2.5 ms and whole-run falor was faster eager (4.0 s / 6.0 s)
than with replay (8.3 s / 19.9 s), so the default stays with the plain calls.
"""
from __future__ import annotations

import logging
import os

import torch

logger = logging.getLogger("ptdeco.falor.decomposition")


class ActivationRecorder:
    """Forward hooks on the target layers that remember the last input / output tensor of each.
    In eager mode they see fresh tensors on every call; during graph capture they see the graph's
    static tensors, which every later replay refills in place -- so ONE captured graph serves the
    analysis of every layer (the per-layer wrappers only need to know where to look)."""

    def __init__(self, root: torch.nn.Module, names: list[str]):
        self.inputs: dict[str, torch.Tensor] = {}
        self.outputs: dict[str, torch.Tensor] = {}
        self._handles = []
        for name in names:
            mod = root.get_submodule(name)
            self._handles.append(mod.register_forward_hook(self._make_hook(name)))

    def _make_hook(self, name: str):
        def hook(_mod, inp, out):
            self.inputs[name] = inp[0]
            self.outputs[name] = out
        return hook

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()
        self.inputs.clear()
        self.outputs.clear()


class GraphedForward:
    WARMUP_CALLS = 2

    def __init__(self, module: torch.nn.Module, enabled: bool = True, recorder=None):
        self.module = module
        self.enabled = enabled and os.environ.get("PTDECO_B200_CUDA_GRAPHS", "0") == "1"
        self.recorder = recorder
        self._entries: dict = {}
        self._calls: dict = {}
        self.replays = 0

    def __call__(self, x):
        if not self.enabled or not isinstance(x, torch.Tensor) or not x.is_cuda or torch.is_grad_enabled():
            return self.module(x)
        key = (tuple(x.shape), x.dtype, x.device)
        entry = self._entries.get(key)
        if entry is None:
            seen = self._calls.get(key, 0)
            self._calls[key] = seen + 1
            if seen < self.WARMUP_CALLS:
                return self.module(x)
            entry = self._capture(x, key)
            if entry is None:
                return self.module(x)
        graph, static_x, static_y = entry
        static_x.copy_(x)
        graph.replay()
        self.replays += 1
        return static_y.clone()

    def _capture(self, x: torch.Tensor, key):
        try:
            static_x = x.clone()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_y = self.module(static_x)
            if not isinstance(static_y, torch.Tensor):
                raise TypeError("model output is not a tensor")
        except Exception as exc:  # the user model is not capturable: stay eager from now on
            logger.info(f"CUDA-graph capture of the model forward failed ({exc!r}); using eager calls")
            self.enabled = False
            torch.cuda.synchronize(x.device)
            return None
        self._entries[key] = (graph, static_x, static_y)
        return self._entries[key]
