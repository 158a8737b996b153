"""ctypes binding of the C-ABI in include/ptdeco_b200.h (the only way Python reaches the kernels).

There is deliberately no CPU or eager-torch fallback: if libptdeco_b200.so is missing, or a call
returns a non-zero code, this module raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libptdeco_b200.so")

F32 = 0
BF16 = 1

_lib: Optional[ctypes.CDLL] = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float
c_double = ctypes.c_double
c_size_t = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ptdeco_b200.h one to one
SIGNATURES = {
    "ptdeco_version": (c_int, []),
    "ptdeco_strerror": (ctypes.c_char_p, [c_int]),
    "ptdeco_syrk_workspace_bytes": (c_size_t, [c_int, c_ll, c_int]),
    "ptdeco_syrk_accumulate": (
        c_int,
        [c_void_p, c_int, c_ll, c_int, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_float, c_void_p,
         c_size_t, c_void_p],
    ),
    "ptdeco_cov_finalize": (
        c_int, [c_void_p, c_ll, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "ptdeco_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "ptdeco_gemm": (
        c_int,
        [c_void_p, c_int, c_int, c_ll, c_void_p, c_int, c_int, c_ll, c_int, c_int, c_int, c_float,
         c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p, c_size_t, c_void_p],
    ),
    "ptdeco_eigh_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ptdeco_eigh": (
        c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_ll, c_void_p, c_size_t, c_void_p]),
    "ptdeco_lowrank_workspace_bytes": (c_size_t, [c_int, c_ll, c_int, c_int, c_int]),
    "ptdeco_lowrank_forward": (
        c_int,
        [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_ll,
         c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ptdeco_nsr_workspace_bytes": (c_size_t, [c_ll]),
    "ptdeco_nsr_metric": (
        c_int,
        [c_void_p, c_void_p, c_int, c_ll, c_ll, c_double, c_void_p, c_size_t, c_void_p, c_void_p]),
    "ptdeco_kl_metric": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_ll, c_void_p, c_void_p]),
    "ptdeco_debug_set": (None, [c_int, c_ll]),
    "ptdeco_debug_get": (c_ll, [c_int]),
}


# ptdeco_debug_set keys of the low-rank forward's runtime switches <- environment variables
LOWRANK_KNOBS = {200: "PTDECO_B200_FORCE_DECODE", 201: "PTDECO_B200_NO_DECODE",
                 202: "PTDECO_B200_NO_FUSED", 203: "PTDECO_B200_NO_PERSISTENT",
                 204: "PTDECO_B200_NO_TMA_STORE", 205: "PTDECO_B200_FUSED_ROT",
                 206: "PTDECO_B200_FUSED_STAGES"}


class NativeError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        # Runtime switches of the forward kernels are read from the environment ONCE, here; the
        # kernels' host code never calls getenv (tests flip them through ptdeco_debug_set).
        for key, name in LOWRANK_KNOBS.items():
            if os.environ.get(name) is not None:
                v = os.environ[name]
                handle.ptdeco_debug_set(key, int(v) if v.lstrip("-").isdigit() else 1)
        if os.environ.get("PTDECO_B200_DETERMINISTIC", "0") == "1":
            # Reproducible mode: no split-K, so every output element is accumulated by exactly one
            # CTA in a fixed order (the default splits short-and-wide reductions over CTAs and
            # combines them with fp32 red.add, whose arrival order varies from run to run).
            handle.ptdeco_debug_set(1, 1)
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().ptdeco_strerror(code).decode()
        raise NativeError(f"{what} failed with code {code}: {msg}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise NativeError(f"unsupported dtype {t.dtype} (kernels take float32 or bfloat16)")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise NativeError(f"{name} must be a CUDA tensor: the ptdeco_b200 kernels only run on sm_100a "
                          "(there is no CPU path)")


class Workspace:
    """Grow-only per-device scratch buffer handed to the C-ABI calls."""

    def __init__(self) -> None:
        self._buf: dict[torch.device, torch.Tensor] = {}

    def get(self, device: torch.device, nbytes: int) -> torch.Tensor:
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        buf = self._buf.get(device)
        if buf is None or buf.numel() < nbytes:
            self._buf[device] = buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return buf

    def release(self) -> None:
        self._buf.clear()


WORKSPACE = Workspace()
