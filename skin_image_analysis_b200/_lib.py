"""ctypes binding of libsia_b200.so (C ABI in include/sia_b200.h).

There is no CPU fallback anywhere in this package: if the library is missing, or a call is made
without a CUDA device, the caller gets an exception -- never a silently slower path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_longlong, c_size_t, c_uint, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# SIA_LIB_PATH: A/B a differently built library (tools/stage_times.py); never a fallback -- it must exist
LIB_PATH = os.environ.get("SIA_LIB_PATH") or os.path.join(HERE, "libsia_b200.so")
DEBUG_LIB_PATH = os.path.join(HERE, "libsia_b200_debug.so")     # bring-up probes (tests / tools only)

LAYOUT_NCHW_F32 = 0
LAYOUT_NCHW_BF16 = 1
LAYOUT_NHWC4_BF16 = 2
NHWC4_PAD = 8            # SIA_NHWC4_PAD: extra pixels per row of the NHWC4 layout

_P = c_void_p

# name -> (restype, argtypes); mirrors include/sia_b200.h line by line
SIGNATURES = {
    "sia_version": (c_int, []),
    "sia_error_string": (c_char_p, [c_int]),
    "sia_device_info": (c_int, [_P, _P, _P]),
    "sia_watchdog_status": (c_uint, [c_int]),
    "sia_preprocess_u8hwc": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P, c_int, c_int,
                                     ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_int, c_int, _P, _P]),
    "sia_preprocess_tc_u8hwc": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, c_int, c_int, c_int,
                                        c_int, c_int, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float), _P, _P]),
    "sia_preprocess_tc2_u8hwc": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, c_int, c_int,
                                         c_int, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float), _P, _P]),
    "sia_preprocess_mma_u8hwc": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, _P, _P, c_int, c_int, c_int,
                                         ctypes.POINTER(c_int32), ctypes.POINTER(c_float), ctypes.POINTER(c_float),
                                         c_int, c_int, _P, _P]),
    "sia_debug_set_mma_warps": (c_int, [c_int]),
    "sia_debug_set_tail_impl": (c_int, [c_int]),
    "sia_debug_set_programmatic_launch": (c_int, [c_int]),
    "sia_preprocess_tv_u8hwc": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, _P, c_int, c_int, _P, c_int,
                                        c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "sia_debug_tv_force_generic": (c_int, [c_int]),
    "sia_nchw_f32_to_nhwc4_bf16": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "sia_chw_u8_to_hwc_u8": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "sia_pad_nhwc_bf16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P]),
    "sia_pack_conv7x7_c3": (c_int, [_P, _P, _P]),
    "sia_pack_conv7x7_c3_bytes": (c_size_t, []),
    "sia_pack_conv3x3": (c_int, [_P, c_int, c_int, _P, _P]),
    "sia_pack_conv3x3_bytes": (c_size_t, [c_int, c_int]),
    "sia_pack_linear_chw_to_hwc": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "sia_pack_linear_chw_to_hwc_padded": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "sia_pack_conv3x3_padded": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "sia_conv7x7_c3_relu_pool2_strided": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P]),
    "sia_head_tail_chain": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, ctypes.POINTER(_P), ctypes.POINTER(_P),
                                    ctypes.POINTER(c_int), _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "sia_conv7x7_c3_relu_pool2": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "sia_conv3x3_relu_pool2": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "sia_linear_splitk": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "sia_linear_splitk_tiled": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "sia_retile_linear_w": (c_int, [_P, c_int, c_int, _P, _P]),
    "sia_head_tail": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int,
                              c_int, _P, _P]),
    "sia_confusion_counts": (c_int, [_P, _P, _P, c_longlong, c_longlong, c_int, c_int, _P, _P]),
    "sia_debug_set_stats": (c_int, [_P]),
    "sia_debug_set_trace": (c_int, [_P]),
}

# include/sia_b200_debug.h group (2): exported by libsia_b200_debug.so
DEBUG_SIGNATURES = {
    "sia_debug_umma_probe": (c_int, [_P, c_int, ctypes.POINTER(c_uint64), ctypes.POINTER(c_uint64), c_int, c_int,
                                     _P, c_int, ctypes.POINTER(c_longlong), _P]),
    "sia_debug_umma_probe_ex": (c_int, [_P, c_int, ctypes.POINTER(c_uint64), ctypes.POINTER(c_uint64), c_int, c_int,
                                        c_int, ctypes.c_uint32, _P, c_int, ctypes.POINTER(c_longlong), _P]),
    "sia_debug_umma_probe_switch": (c_int, [c_int, c_int]),
    "sia_debug_umma_ts_probe": (c_int, [_P, c_int, _P, c_int, c_int, ctypes.POINTER(c_uint64), c_int, c_int,
                                        ctypes.c_uint32, _P, _P]),
    "sia_debug_tma_probe": (c_int, [_P, c_int, ctypes.POINTER(c_uint64), ctypes.POINTER(c_uint64),
                                    ctypes.POINTER(ctypes.c_uint32), c_int, ctypes.POINTER(c_int), _P, c_int, c_int,
                                    c_int, ctypes.POINTER(c_longlong), _P]),
    "sia_debug_tmem_ld_rates": (c_int, [ctypes.POINTER(ctypes.c_double), c_int]),
    "sia_debug_alu_rates": (c_int, [ctypes.POINTER(ctypes.c_double), c_int]),
}

_lib = None
_debug_lib = None


class SiaError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SiaError(
            f"{LIB_PATH} is missing: build it with `python -m skin_image_analysis_b200.build` "
            "(needs nvcc; there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_debug() -> ctypes.CDLL:
    """The probe library (tests/test_umma_probe.py, tools/): never needed by the product path."""
    global _debug_lib
    if _debug_lib is not None:
        return _debug_lib
    if not os.path.exists(DEBUG_LIB_PATH):
        raise SiaError(f"{DEBUG_LIB_PATH} is missing: build it with `python -m skin_image_analysis_b200.build`")
    lib = ctypes.CDLL(DEBUG_LIB_PATH)
    for name, (res, args) in DEBUG_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _debug_lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sia_error_string(rc).decode()
        wd = load().sia_watchdog_status(0)
        extra = f" [watchdog 0x{wd:08x}]" if wd not in (0, 0xFFFFFFFF) else ""
        raise SiaError(f"{what or 'sia call'} failed: {msg} (code {rc}){extra}")


def require_cuda(*tensors) -> None:
    import torch
    if not torch.cuda.is_available():
        raise SiaError("no CUDA device: the sm_100a evaluation path has no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SiaError("expected CUDA tensors (the sm_100a evaluation path has no CPU fallback)")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
