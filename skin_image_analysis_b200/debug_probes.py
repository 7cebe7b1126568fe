"""Python wrappers of the bring-up probes (include/sia_b200_debug.h group (2), libsia_b200_debug.so).

Test / tool infrastructure: nothing on the product path imports this module, and the product library
libsia_b200.so does not contain the probe kernels.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .ops import _need

# -------------------------------------------------------------------------------------------------
# bring-up probe
# -------------------------------------------------------------------------------------------------
def umma_probe(image: torch.Tensor, a_descs, b_descs, n: int, repeat: int = 1, want_cycles: bool = False):
    _need(image, torch.uint8, "image")
    k = len(a_descs)
    a = (ctypes.c_uint64 * k)(*[int(d) for d in a_descs])
    b = (ctypes.c_uint64 * k)(*[int(d) for d in b_descs])
    out = torch.zeros((128, n), dtype=torch.float32, device=image.device)
    cyc = ctypes.c_longlong(0)
    check(_lib.load_debug().sia_debug_umma_probe(ptr(image), image.numel(), a, b, k, n, ptr(out), repeat,
                                           ctypes.byref(cyc) if want_cycles else None, stream_ptr()),
          "sia_debug_umma_probe")
    torch.cuda.synchronize()
    return (out, cyc.value) if want_cycles else out


def umma_ts_probe(image: torch.Tensor, a_words: torch.Tensor, a_col_step: int, b_descs, n: int, idesc: int = 0):
    """tcgen05.mma with A in tensor memory (bring-up): a_words [128, a_cols] int32 = the words thread m stores to
    TMEM lane m; returns the 128 x n fp32 accumulator."""
    _need(image, torch.uint8, "image")
    _need(a_words, torch.int32, "a_words")
    k = len(b_descs)
    b = (ctypes.c_uint64 * k)(*[int(d) for d in b_descs])
    out = torch.zeros((128, n), dtype=torch.float32, device=image.device)
    check(_lib.load_debug().sia_debug_umma_ts_probe(ptr(image), image.numel(), ptr(a_words), a_words.shape[1], int(a_col_step),
                                              b, k, n, int(idesc), ptr(out), stream_ptr()), "sia_debug_umma_ts_probe")
    torch.cuda.synchronize()
    return out


def umma_probe_i8(image: torch.Tensor, a_descs, b_descs, n: int, idesc: int, repeat: int = 1,
                  want_cycles: bool = False):
    """kind::i8 variant of ``umma_probe``: returns the 128 x n int32 accumulator."""
    _need(image, torch.uint8, "image")
    k = len(a_descs)
    a = (ctypes.c_uint64 * k)(*[int(d) for d in a_descs])
    b = (ctypes.c_uint64 * k)(*[int(d) for d in b_descs])
    out = torch.zeros((128, n), dtype=torch.int32, device=image.device)
    cyc = ctypes.c_longlong(0)
    check(_lib.load_debug().sia_debug_umma_probe_ex(ptr(image), image.numel(), a, b, k, n, 1, int(idesc), ptr(out), repeat,
                                              ctypes.byref(cyc) if want_cycles else None, stream_ptr()),
          "sia_debug_umma_probe_ex")
    torch.cuda.synchronize()
    return (out, cyc.value) if want_cycles else out


def tma_probe(t: torch.Tensor, dims, strides_bytes, box, swizzle_bytes: int, coords, repeat: int = 1,
              step_dim: int = 0, step: int = 0):
    """One TMA box load of a bf16 tensor -> the shared-memory bytes as uint8 (bring-up tests).  With
    repeat > 1 returns (bytes, cycles): `repeat` loads in flight, coordinate step_dim advanced by step."""
    _need(t, torch.bfloat16, "t")
    rank = len(dims)
    nbytes = 2
    for b in box:
        nbytes *= int(b)
    out = torch.zeros(nbytes, dtype=torch.uint8, device=t.device)
    cyc = ctypes.c_longlong(0)
    check(_lib.load_debug().sia_debug_tma_probe(
        ptr(t), rank, (ctypes.c_uint64 * rank)(*[int(d) for d in dims]),
        (ctypes.c_uint64 * max(1, rank - 1))(*[int(s) for s in strides_bytes]),
        (ctypes.c_uint32 * rank)(*[int(b) for b in box]), int(swizzle_bytes),
        (ctypes.c_int * rank)(*[int(c) for c in coords]), ptr(out), int(repeat), int(step_dim), int(step),
        ctypes.byref(cyc) if repeat > 1 else None, stream_ptr()), "sia_debug_tma_probe")
    torch.cuda.synchronize()
    return (out, cyc.value) if repeat > 1 else out
