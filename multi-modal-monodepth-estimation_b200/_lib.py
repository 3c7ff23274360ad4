"""ctypes binding of libb200swin.so (the C-ABI declared in include/b200swin.h).

There is no CPU fallback: if the library is missing or a call fails, the op raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libb200swin.so")

F32, BF16 = 0, 1
EPI_NONE, EPI_GELU, EPI_QKV, EPI_DGELU, EPI_ADD, EPI_RELU = 0, 1, 2, 3, 4, 5

_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_int64, c_float, c_size_t = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float,
                                               ctypes.c_size_t)
P, I, L, F, Z = c_void_p, c_int, c_int64, c_float, c_size_t

# name -> (restype, argtypes); must list every symbol of include/b200swin.h (tests/test_cabi.py checks)
SIGNATURES = {
    "b200swin_version": (c_int, []),
    "b200swin_last_error": (ctypes.c_char_p, []),
    "b200swin_silog_workspace_bytes": (Z, [L]),
    "b200swin_silog_fwd": (I, [P, I, P, L, F, P, P, P, Z, P]),
    "b200swin_silog_bwd": (I, [P, I, P, L, F, P, P, P, P]),
    "b200swin_window_gather": (I, [P, P, I, I, I, I, I, I, I, P]),
    "b200swin_window_scatter": (I, [P, P, I, I, I, I, I, I, I, P]),
    "b200swin_shift_mask": (I, [P, I, I, I, I, P]),
    "b200swin_patch_merge": (I, [P, P, I, I, I, I, I, I, P]),
    "b200swin_patchify": (I, [P, I, P, I, I, I, I, I, I, I, P]),
    "b200swin_dwconv3x3": (I, [P, P, P, I, I, I, I, I, I, P]),
    "b200swin_dwconv3x3_wgrad_workspace_bytes": (Z, [I, I, I, I]),
    "b200swin_dwconv3x3_wgrad": (I, [P, P, P, I, I, I, I, I, P, Z, P]),
    "b200swin_cpb_fwd": (I, [P, P, P, P, P, P, P, I, I, I, P]),
    "b200swin_cpb_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, P]),
    "b200swin_ln_fwd": (I, [P, P, P, P, P, L, P, P, P, L, I, F, I, P]),
    "b200swin_ln_fwd_stream32": (I, [P, P, P, P, P, L, P, P, P, P, L, I, F, P]),
    "b200swin_ln_fwd_sum": (I, [P, P, I, P, P, P, P, P, P, P, L, I, F, P]),
    "b200swin_ln_bwd_workspace_bytes": (Z, [L, I]),
    "b200swin_ln_bwd": (I, [P, P, P, P, P, P, L, P, P, P, P, L, I, I, P, Z, P]),
    "b200swin_attn_fwd": (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, P]),
    "b200swin_attn_bwd_workspace_bytes": (Z, [I, I, I, I, I, I, I]),
    "b200swin_attn_bwd_colsum_supported": (I, [I, I, I]),
    "b200swin_attn_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, I, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P, Z, P]),
    "b200swin_mha_fwd": (I, [P, P, P, L, L, L, P, L, P, I, I, I, I, I, F, I, P]),
    "b200swin_mha_bwd_workspace_bytes": (Z, [I, I, I]),
    "b200swin_mha_bwd": (I, [P, P, P, L, L, L, P, L, P, L, P, P, P, P, L, L, L, I, I, I, I, I, F, I, P, Z, P]),
    "b200swin_mha_avg_weights": (I, [P, P, L, L, P, P, I, I, I, I, I, F, I, P]),
    "b200swin_gemm_splits": (I, [L, L, L]),
    "b200swin_gemm_workspace_bytes": (Z, [L, L, I]),
    "b200swin_gemm_bf16": (I, [P, P, I, P, P, I, L, L, L, I, P, P, P, P, P, I, P, I, I, P, Z, P]),
    "b200swin_split_bf16": (I, [P, P, P, L, P]),
    "b200swin_colsum_workspace_bytes": (Z, [L, I]),
    "b200swin_colsum": (I, [P, I, L, L, L, I, P, P, P, Z, P]),
    "b200swin_adamw_chunk": (I, []),
    "b200swin_adamw_step": (I, [P, P, P, P, P, P, P, P, P, P, ctypes.c_double, ctypes.c_double, F, F, L, P]),
}


def load() -> ctypes.CDLL:
    """Load (once) and return the library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"b200swin: {LIB_PATH} is missing - build it with `python __graft_entry__.py build` "
                    "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for these ops.")
            cdll = ctypes.CDLL(LIB_PATH)
            lib = _Instrumented()
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(cdll, name)
                fn.restype = res
                fn.argtypes = args
                setattr(lib, name, _wrap(name, fn) if name in KERNELS_PER_CALL else fn)
            _lib = lib
    return _lib


class _Instrumented:
    """Namespace of the bound entry points (launching ones wrapped by the counters below)."""


# kernels launched per C-ABI call (for the bench's gpu_launches claim); split-K gemm adds its reduce
KERNELS_PER_CALL = {
    "b200swin_silog_fwd": 2, "b200swin_silog_bwd": 1, "b200swin_window_gather": 1, "b200swin_window_scatter": 1,
    "b200swin_shift_mask": 1, "b200swin_patch_merge": 1, "b200swin_patchify": 1, "b200swin_cpb_fwd": 1, "b200swin_cpb_bwd": 1, "b200swin_ln_fwd": 1, "b200swin_ln_fwd_stream32": 1, "b200swin_ln_fwd_sum": 1, "b200swin_ln_bwd": 2, "b200swin_attn_fwd": 1,
    "b200swin_attn_bwd": 2, "b200swin_gemm_bf16": 1, "b200swin_split_bf16": 1, "b200swin_colsum": 2,
    "b200swin_adamw_step": 1, "b200swin_mha_fwd": 1, "b200swin_mha_bwd": 3, "b200swin_mha_avg_weights": 1,
    "b200swin_dwconv3x3": 1, "b200swin_dwconv3x3_wgrad": 2,
}

COUNTERS = {"launches": 0, "calls": {}}
# optional live timing of the entry points with CUDA events on the launching stream (bench.py rooflines):
# TIMING = {"name": <symbol> | "*" | None, "events": [(start_event, end_event, symbol, args)]}
TIMING = {"name": None, "events": []}


def _wrap(name, fn):
    per_call = KERNELS_PER_CALL[name]

    def call(*args):
        n = per_call
        if name == "b200swin_gemm_bf16" and args[18] > 1:
            n += 1
        if name == "b200swin_attn_bwd" and args[25] in (1, 2) and (args[25] == 2 or args[22] not in (4, 6, 7, 8, 12)):
            n += 1                                     # KV-blocked backward: prep + dQ pass + dK/dV pass
        if name == "b200swin_mha_bwd" and args[23] == F32:
            n += 1                                     # CUDA-core backward: prep + dQ + dK + dV
        COUNTERS["launches"] += n
        COUNTERS["calls"][name] = COUNTERS["calls"].get(name, 0) + 1
        if TIMING["name"] == name or TIMING["name"] == "*":
            st = torch.cuda.current_stream()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = fn(*args)
            e1.record(st)
            TIMING["events"].append((e0, e1, name, tuple(a for a in args)))
            return rc
        return fn(*args)

    return call


def reset_counters():
    COUNTERS["launches"] = 0
    COUNTERS["calls"] = {}
    TIMING["events"] = []


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200swin_last_error()
        raise RuntimeError(f"b200swin {what} failed (code {rc}): {msg.decode() if msg else ''}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"b200swin supports float32 and bfloat16 tensors, got {t.dtype}")


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_of(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b200swin ops run on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")
