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
    "ptdeco_syrk_accumulate_ex": (
        c_int,
        [c_void_p, c_int, c_ll, c_int, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_float, c_void_p,
         c_size_t, c_void_p, ctypes.c_uint],
    ),
    "ptdeco_cov_finalize": (
        c_int, [c_void_p, c_ll, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "ptdeco_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "ptdeco_gemm": (
        c_int,
        [c_void_p, c_int, c_int, c_ll, c_void_p, c_int, c_int, c_ll, c_int, c_int, c_int, c_float,
         c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p, c_size_t, c_void_p],
    ),
    "ptdeco_gemm_ex": (
        c_int,
        [c_void_p, c_int, c_int, c_ll, c_void_p, c_int, c_int, c_ll, c_int, c_int, c_int, c_float,
         c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p, c_size_t, c_void_p, ctypes.c_uint],
    ),
    "ptdeco_eigh_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ptdeco_eigh_ex": (
        c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_ll, c_void_p, c_size_t, c_void_p,
                ctypes.c_uint]),
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
                 206: "PTDECO_B200_FUSED_STAGES", 208: "PTDECO_B200_FUSED_KSPLIT"}


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


FLAG_DETERMINISTIC = 1

_deterministic = os.environ.get("PTDECO_B200_DETERMINISTIC", "0") == "1"


def set_deterministic(on: bool) -> None:
    """Reproducible mode: no split-K, so every output element is accumulated by exactly one CTA in a
    fixed order (the default splits short-and-wide reductions over CTAs and combines them with fp32
    red.add, whose arrival order varies from run to run). A per-call flag of the C-ABI (`*_ex`
    entry points); this module only remembers what the Python layer should pass.
    PTDECO_B200_DETERMINISTIC=1 sets the initial value."""
    global _deterministic
    _deterministic = bool(on)


def call_flags() -> int:
    return FLAG_DETERMINISTIC if _deterministic else 0


def device_of(t: torch.Tensor):
    """Context manager making `t`'s device current for a C-ABI call: kernel attributes, SM counts
    and cooperative launches belong to the current device, which is not necessarily the tensor's."""
    if t.device.index is None or t.device.index == torch.cuda.current_device():
        return _NULL_CONTEXT
    return torch.cuda.device(t.device)


class _NullContext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_CONTEXT = _NullContext()


class Workspace:
    """Scratch buffers handed to the C-ABI calls, ONE PER (device, stream): calls on one stream are
    ordered, so they can share a buffer; calls on different streams (or threads using different
    streams) must not -- the decode kernel keeps its grid-barrier / ticket words and the rank-k
    intermediate in there. Buffers only grow; a replaced buffer goes back to torch's stream-aware
    caching allocator, which will not hand it to another stream while this one still uses it.
    During CUDA-graph capture every request gets its OWN buffer from the graph's pool and it is
    kept alive for the life of the process: a captured pointer must never be freed or reused."""

    def __init__(self) -> None:
        self._buf: dict[tuple[int, int], torch.Tensor] = {}
        self._captured: list[torch.Tensor] = []

    def get(self, device: torch.device, nbytes: int) -> torch.Tensor:
        device = torch.device(device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        device = torch.device("cuda", index)
        if nbytes <= 0:
            nbytes = 1
        if torch.cuda.is_current_stream_capturing():
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self._captured.append(buf)
            return buf
        key = (index, torch.cuda.current_stream(device).cuda_stream)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            self._buf[key] = buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return buf

    def release(self) -> None:
        self._buf.clear()


WORKSPACE = Workspace()
