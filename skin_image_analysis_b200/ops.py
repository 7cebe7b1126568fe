"""Tensor-level wrappers over the C ABI (include/sia_b200.h).

Every function takes / returns CUDA tensors, checks dtype / contiguity / device on the host and
launches on the current torch stream.  PyTorch is used for allocation and streams only.
"""
from __future__ import annotations

import ctypes
import os
from functools import lru_cache

import numpy as np
import torch

from . import _lib
from . import resize_weights as _rw
from ._lib import (LAYOUT_NCHW_BF16, LAYOUT_NCHW_F32, LAYOUT_NHWC4_BF16, NHWC4_PAD, SiaError, check, ptr,
                   stream_ptr)

__all__ = [
    "preprocess_u8hwc", "nchw_f32_to_nhwc4", "pack_conv7x7_c3", "pack_conv3x3", "pack_linear_chw_to_hwc",
    "conv7x7_c3_relu_pool2", "conv3x3_relu_pool2", "linear_splitk", "retile_linear_w", "TiledLinearWeight", "head_tail", "head_tail_chain", "confusion_counts",
    "pad_nhwc", "chw_to_hwc_u8", "LAYOUT_NCHW_F32", "LAYOUT_NCHW_BF16", "LAYOUT_NHWC4_BF16", "NHWC4_PAD",
]


def _need(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise SiaError(f"{name}: expected a CUDA tensor (there is no CPU fallback)")
    if t.device.index != torch.cuda.current_device():
        # launches go to the CURRENT device's stream: one process per GPU, or wrap the call in torch.cuda.device(...)
        raise SiaError(f"{name}: tensor lives on {t.device} but the current CUDA device is "
                       f"cuda:{torch.cuda.current_device()}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


# -------------------------------------------------------------------------------------------------
# K1-K3 preprocess
# -------------------------------------------------------------------------------------------------
class _DeviceTables:
    def __init__(self, t: _rw.ResizeTables, device):
        self.host = t
        self.x_off = torch.from_numpy(t.x_off).to(device)
        self.x_w = torch.from_numpy(np.ascontiguousarray(t.x_w)).to(device)
        self.x_wq = None if t.x_wq is None else torch.from_numpy(t.x_wq.view(np.int32).copy()).to(device)
        self.row_w = torch.from_numpy(np.ascontiguousarray(t.row_w)).to(device)
        self.row_emit = torch.from_numpy(np.ascontiguousarray(t.row_emit)).to(device)
        self.y_first_last = torch.from_numpy(np.ascontiguousarray(t.y_first_last)).to(device)


@lru_cache(maxsize=64)
def _tables(device_index: int, src_h: int, src_w: int, out_h: int, out_w: int, scale: float, antialias):
    t = _rw.build_tables(src_h, src_w, out_h, out_w, scale=scale, antialias=antialias)
    return _DeviceTables(t, torch.device("cuda", device_index))


class _DeviceTcTables:
    def __init__(self, t: _rw.TcTables, device):
        self.host = t
        self.a_packed = torch.from_numpy(t.a_packed).to(device)
        self.lane_scale = torch.from_numpy(t.lane_scale).to(device)
        self.tile_row0 = torch.from_numpy(t.tile_row0).to(device)
        self.items = torch.from_numpy(t.items).to(device)


@lru_cache(maxsize=64)
def _tc_tables(device_index: int, src_h: int, src_w: int, out_h: int, out_w: int, antialias):
    """Tables of the tensor-core kernel, or None when the geometry does not fit it."""
    try:
        t = _rw.build_tc_tables(src_h, src_w, out_h, out_w, antialias=antialias)
    except ValueError:
        return None
    return _DeviceTcTables(t, torch.device("cuda", device_index))


class _DeviceTc2Tables:
    def __init__(self, t: _rw.Tc2Tables, device):
        self.host = t
        self.a_packed = torch.from_numpy(t.vert.a_packed).to(device)
        self.lane_scale = torch.from_numpy(t.vert.lane_scale).to(device)
        self.tile_row0 = torch.from_numpy(t.vert.tile_row0).to(device)
        self.b2 = torch.from_numpy(t.b2).to(device)
        self.block_meta = torch.from_numpy(np.ascontiguousarray(t.block_meta)).to(device)
        self.slot_scale = torch.from_numpy(np.ascontiguousarray(t.slot_scale)).to(device)


@lru_cache(maxsize=64)
def _tc2_tables(device_index: int, src_h: int, src_w: int, out_h: int, out_w: int, antialias):
    """Tables of the two-product tensor-core kernel, or None when the geometry does not fit it."""
    try:
        t = _rw.build_tc2_tables(src_h, src_w, out_h, out_w, antialias=antialias)
    except ValueError:
        return None
    return _DeviceTc2Tables(t, torch.device("cuda", device_index))


class _DeviceMmaTables:
    def __init__(self, t: _rw.MmaTables, device):
        self.host = t
        self.wy_frag = torch.from_numpy(np.ascontiguousarray(t.wy_frag).view(np.int32)).to(device)
        self.r0 = torch.from_numpy(np.ascontiguousarray(t.r0)).to(device)
        self.wx_frag = torch.from_numpy(np.ascontiguousarray(t.wx_frag).view(np.int32)).to(device)
        self.wx_mask = torch.from_numpy(np.ascontiguousarray(t.wx_mask).view(np.int32)).to(device)
        self.tile_begin = torch.from_numpy(np.ascontiguousarray(t.tile_begin)).to(device)
        qs, cs = _rw.MMA_ROW_MAPS[t.row_map]
        self.q_stride, self.c_row = qs, (ctypes.c_int32 * 4)(*cs)


@lru_cache(maxsize=64)
def _mma_tables(device_index: int, src_h: int, src_w: int, out_h: int, out_w: int, antialias):
    """Tables of the warp-MMA kernel, or None when the geometry does not fit it."""
    try:
        t = _rw.build_mma_tables(src_h, src_w, out_h, out_w, antialias=antialias)
    except ValueError:
        return None
    return _DeviceMmaTables(t, torch.device("cuda", device_index))


_LAYOUT_DTYPE = {LAYOUT_NCHW_F32: torch.float32, LAYOUT_NCHW_BF16: torch.bfloat16, LAYOUT_NHWC4_BF16: torch.bfloat16}


def preprocess_u8hwc(src: torch.Tensor, size, layout: int = LAYOUT_NCHW_F32, mean=(0.0, 0.0, 0.0),
                     std=(1.0, 1.0, 1.0), scale: float = 1.0 / 255.0, antialias="skimage",
                     rows_per_cta: int = 32, out: torch.Tensor | None = None,
                     fixed_point: bool = True, impl: str = "auto") -> torch.Tensor:
    """[B,H,W,3] uint8 -> resized / scaled / normalised batch in ``layout``.

    Defaults reproduce the reference transform exactly: ``float32(u8)/255`` (tone_bias_dataset.py:335),
    ``skimage.transform.resize`` (:425), no mean/std, CHW (:470).  ``fixed_point`` lets the bf16 layouts
    use the 15-bit integer-dot-product horizontal pass (<= 0.03 bf16 ulp from the fp32 pass); the fp32
    layout always uses fp32 arithmetic.  ``impl``: "mma" (csrc/preprocess_mma.cu: both passes on mma.sync with
    register-built operands; NHWC4 layout, src_w % 8 == 0, even src_h, out_w % 8 == 0 -- what "auto" picks first),
    "cuda_core" (csrc/preprocess.cu), "tensor_core"
    (csrc/preprocess_tc.cu: vertical pass as a tcgen05 GEMM; NHWC4 layout, src_w % 8 == 0, <= 256 source rows
    per 128 output rows), "tensor_core2" (csrc/preprocess_tc2.cu: the horizontal pass is a second tcgen05 product
    too; additionally <= 4 + 13 + 4 output columns per 40-pixel block, out_w % 4 == 0) or "auto" = tensor_core2, else
    tensor_core, whenever they apply and ``fixed_point`` allows a reduced-precision pass (0.102 / 0.109 ms vs
    0.137 ms at the bench shape), else cuda_core.
    """
    _need(src, torch.uint8, "src")
    if src.dim() != 4 or src.shape[3] != 3:
        raise ValueError("src must be [B,H,W,3] uint8")
    b, sh, sw, _ = src.shape
    oh, ow = int(size[0]), int(size[1])
    tab = _tables(src.device.index, sh, sw, oh, ow, float(scale), antialias)
    shape = (b, oh, ow + NHWC4_PAD, 4) if layout == LAYOUT_NHWC4_BF16 else (b, 3, oh, ow)
    if out is None:
        out = torch.empty(shape, dtype=_LAYOUT_DTYPE[layout], device=src.device)
    else:
        _need(out, _LAYOUT_DTYPE[layout], "out")
        if tuple(out.shape) != shape:
            raise ValueError(f"out must have shape {shape}")
    osc = (ctypes.c_float * 3)(*[1.0 / float(s) for s in std])
    obi = (ctypes.c_float * 3)(*[-float(m) / float(s) for m, s in zip(mean, std)])
    if impl == "auto" and os.environ.get("SIA_PREPROCESS_IMPL"):      # A/B switch for whole-pipeline timing runs
        impl = os.environ["SIA_PREPROCESS_IMPL"]
    if impl not in ("auto", "cuda_core", "tensor_core", "tensor_core2", "mma"):
        raise ValueError("impl must be 'auto', 'mma', 'cuda_core', 'tensor_core' or 'tensor_core2'")
    if impl == "mma" or (impl == "auto" and fixed_point and layout == LAYOUT_NHWC4_BF16):
        tm = _mma_tables(src.device.index, sh, sw, oh, ow, antialias) if layout == LAYOUT_NHWC4_BF16 else None
        if tm is not None and src.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0:
            h = tm.host
            mul = (ctypes.c_float * 3)(*[_rw.MMA_OUT_SCALE * float(scale) / float(s) for s in std])
            check(_lib.load().sia_preprocess_mma_u8hwc(
                ptr(src), b, sh, sw, ptr(tm.wy_frag), ptr(tm.r0), h.n_msteps, h.kv, ptr(tm.wx_frag), ptr(tm.wx_mask),
                ptr(tm.tile_begin), h.n_groups, h.n_tiles, tm.q_stride, tm.c_row, mul, obi, oh, ow, ptr(out),
                stream_ptr()), "sia_preprocess_mma_u8hwc")
            return out
        if impl == "mma":
            raise SiaError("warp-MMA preprocess does not support this geometry / layout")
    auto_tc = impl == "auto" and fixed_point and layout == LAYOUT_NHWC4_BF16 and src.data_ptr() % 16 == 0
    t2 = None
    if (impl == "tensor_core2" or auto_tc) and layout == LAYOUT_NHWC4_BF16 and out.data_ptr() % 32 == 0:
        t2 = _tc2_tables(src.device.index, sh, sw, oh, ow, antialias)
    if impl == "tensor_core2" and (t2 is None or src.data_ptr() % 16 != 0):
        raise SiaError("two-product tensor-core preprocess does not support this geometry / layout")
    if t2 is not None:
        tsc = (ctypes.c_float * 3)(*[float(scale) / float(s) for s in std])
        v = t2.host.vert
        check(_lib.load().sia_preprocess_tc2_u8hwc(
            ptr(src), b, sh, sw, ptr(t2.a_packed), ptr(t2.lane_scale), ptr(t2.tile_row0), v.n_tiles, v.tile_rows,
            ptr(t2.b2), ptr(t2.block_meta), ptr(t2.slot_scale), t2.host.n_blocks, t2.host.last_block_cols, oh, ow,
            tsc, obi, ptr(out), stream_ptr()), "sia_preprocess_tc2_u8hwc")
        return out
    if impl == "tensor_core" or (impl == "auto" and fixed_point and layout == LAYOUT_NHWC4_BF16):
        tc = _tc_tables(src.device.index, sh, sw, oh, ow, antialias) if layout == LAYOUT_NHWC4_BF16 else None
        if tc is not None and src.data_ptr() % 16 == 0 and out.data_ptr() % 32 == 0:
            tsc = (ctypes.c_float * 3)(*[float(scale) / float(s) for s in std])
            check(_lib.load().sia_preprocess_tc_u8hwc(
                ptr(src), b, sh, sw, ptr(tc.a_packed), ptr(tc.lane_scale), ptr(tc.tile_row0), tc.host.n_tiles,
                tc.host.tile_rows, ptr(tc.items), tc.host.n_items, tc.host.n_blocks, tc.host.last_block_cols,
                int(tc.host.pads_in_schedule), oh, ow,
                tsc, obi, ptr(out), stream_ptr()), "sia_preprocess_tc_u8hwc")
            return out
        if impl == "tensor_core":                   # asked for explicitly: refuse loudly; "auto" falls through to the
            raise SiaError("tensor-core preprocess does not support this geometry / layout")   # CUDA-core kernel
    check(_lib.load().sia_preprocess_u8hwc(
        ptr(src), b, sh, sw, ptr(tab.x_off), ptr(tab.x_w), ptr(tab.x_wq) if fixed_point else 0, tab.host.x_taps,
        ptr(tab.row_w), ptr(tab.row_emit),
        ptr(tab.y_first_last), oh, ow, osc, obi, layout, int(rows_per_cta), ptr(out), stream_ptr()),
        "sia_preprocess_u8hwc")
    return out


class _DeviceTvTables:
    def __init__(self, src_h, src_w, out_h, out_w, mean, std, device):
        self.x = _rw.tv_axis(src_w, out_w)
        self.y = _rw.tv_axis(src_h, out_h)
        self.tile_rows, self.max_rows = _rw.tv_tile_plan(self.y, src_w * 3, out_w, self.x.taps)
        self.x_min = torch.from_numpy(self.x.xmin).to(device)
        self.x_w = torch.from_numpy(np.ascontiguousarray(self.x.w.T)).to(device)      # tap-major
        self.y_min = torch.from_numpy(self.y.xmin).to(device)
        self.y_w = torch.from_numpy(self.y.w).to(device)
        self.lut = torch.from_numpy(_rw.tv_normalise_lut(mean, std)).to(device)


@lru_cache(maxsize=64)
def _tv_tables(device_index: int, src_h: int, src_w: int, out_h: int, out_w: int, mean: tuple, std: tuple):
    return _DeviceTvTables(src_h, src_w, out_h, out_w, mean, std, torch.device("cuda", device_index))


IMAGENET_MEAN = (0.485, 0.456, 0.406)      # notebooks/ToneClassifier/CNNTrialDataset.py:73
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_tv_u8hwc(src: torch.Tensor, size=(224, 224), layout: int = LAYOUT_NCHW_F32, mean=IMAGENET_MEAN,
                        std=IMAGENET_STD, out: torch.Tensor | None = None, planar: bool = False) -> torch.Tensor:
    """[B,H,W,3] uint8 decode buffers (``planar=True``: [B,3,H,W], what ``torchvision.io.read_image`` returns,
    CNNTrialDataset.py:93) -> the ToneClassifier test transform (CNNTrialDataset.py:71-76):
    ``v2.Resize(size)`` on uint8 (bilinear, antialias: ATen's fixed-point resampler, uint8 after each axis) ->
    ``v2.ToDtype(float32, scale=True)`` -> ``v2.Normalize(mean, std)``.  Bit-exact with torchvision for the float32
    layout (integer taps + a 3 x 256 table of the float32 normalisation); the bf16 layouts round that value once."""
    _need(src, torch.uint8, "src")
    if src.dim() != 4 or src.shape[1 if planar else 3] != 3:
        raise ValueError("src must be [B,H,W,3] uint8 ([B,3,H,W] with planar=True)")
    b, sh, sw = (src.shape[0], src.shape[2], src.shape[3]) if planar else src.shape[:3]
    oh, ow = int(size[0]), int(size[1])
    tab = _tv_tables(src.device.index, sh, sw, oh, ow, tuple(float(m) for m in mean), tuple(float(v) for v in std))
    shape = (b, oh, ow + NHWC4_PAD, 4) if layout == LAYOUT_NHWC4_BF16 else (b, 3, oh, ow)
    if out is None:
        out = torch.empty(shape, dtype=_LAYOUT_DTYPE[layout], device=src.device)
    else:
        _need(out, _LAYOUT_DTYPE[layout], "out")
        if tuple(out.shape) != shape:
            raise ValueError(f"out must have shape {shape}")
    check(_lib.load().sia_preprocess_tv_u8hwc(
        ptr(src), b, sh, sw, ptr(tab.x_min), ptr(tab.x_w), tab.x.taps, tab.x.precision, ptr(tab.y_min), ptr(tab.y_w),
        tab.y.taps, tab.y.precision, ptr(tab.lut), oh, ow, tab.tile_rows, tab.max_rows, layout, int(bool(planar)),
        ptr(out), stream_ptr()), "sia_preprocess_tv_u8hwc")
    return out


def nchw_f32_to_nhwc4(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _need(x, torch.float32, "x")
    b, c, h, w = x.shape
    if c != 3:
        raise ValueError("expected [B,3,H,W]")
    if out is None:
        out = torch.empty((b, h, w + NHWC4_PAD, 4), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().sia_nchw_f32_to_nhwc4_bf16(ptr(x), b, h, w, ptr(out), stream_ptr()), "sia_nchw_f32_to_nhwc4_bf16")
    return out


# -------------------------------------------------------------------------------------------------
# weight packing
# -------------------------------------------------------------------------------------------------
def pack_conv7x7_c3(w: torch.Tensor) -> torch.Tensor:
    _need(w, torch.float32, "w")
    if tuple(w.shape) != (32, 3, 7, 7):
        raise ValueError("conv1 weight must be [32,3,7,7]")
    lib = _lib.load()
    out = torch.empty(lib.sia_pack_conv7x7_c3_bytes(), dtype=torch.uint8, device=w.device)
    check(lib.sia_pack_conv7x7_c3(ptr(w), ptr(out), stream_ptr()), "sia_pack_conv7x7_c3")
    return out


def pack_conv3x3(w: torch.Tensor, cin_pad: int | None = None, cout_pad: int | None = None) -> torch.Tensor:
    """[cout,cin,3,3] fp32 -> the packed bf16 operand for buffers of cin_pad x cout_pad channels (zero padded)."""
    _need(w, torch.float32, "w")
    cout, cin, kh, kw = w.shape
    if (kh, kw) != (3, 3):
        raise ValueError("expected a 3x3 kernel")
    cin_pad, cout_pad = cin_pad or cin, cout_pad or cout
    lib = _lib.load()
    out = torch.empty(lib.sia_pack_conv3x3_bytes(cin_pad, cout_pad), dtype=torch.uint8, device=w.device)
    check(lib.sia_pack_conv3x3_padded(ptr(w), cin, cout, cin_pad, cout_pad, ptr(out), stream_ptr()),
          "sia_pack_conv3x3_padded")
    return out


def pack_linear_chw_to_hwc(w: torch.Tensor, c: int, hw: int, n_pad: int | None = None,
                           c_pad: int | None = None) -> torch.Tensor:
    """[n, c*hw] fp32 (CHW-flattened columns) -> bf16 [n_pad, hw*c_pad] (HWC-flattened, zero padded)."""
    _need(w, torch.float32, "w")
    n, k = w.shape
    if k != c * hw:
        raise ValueError("in_features != c*hw")
    n_pad, c_pad = n_pad or n, c_pad or c
    out = torch.empty((n_pad, hw * c_pad), dtype=torch.bfloat16, device=w.device)
    check(_lib.load().sia_pack_linear_chw_to_hwc_padded(ptr(w), n, c, hw, n_pad, c_pad, ptr(out), stream_ptr()),
          "sia_pack_linear_chw_to_hwc_padded")
    return out


# -------------------------------------------------------------------------------------------------
# K4 conv blocks
# -------------------------------------------------------------------------------------------------
def conv7x7_c3_relu_pool2(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor,
                          out: torch.Tensor | None = None, c_offset: int = 0) -> torch.Tensor:
    """One 32-output-channel slice of the first block, written at channel ``c_offset`` of ``out`` [B,H/2,W/2,C]."""
    _need(x, torch.bfloat16, "x")
    _need(w_packed, torch.uint8, "w_packed")
    _need(bias, torch.float32, "bias")
    b, h, wp, c = x.shape
    w = wp - NHWC4_PAD
    if c != 4 or w < 16:
        raise ValueError("expected padded NHWC4 input [B,H,W+8,4]")
    if out is None:
        out = torch.empty((b, h // 2, w // 2, 32), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().sia_conv7x7_c3_relu_pool2_strided(ptr(x), b, h, w, ptr(w_packed), ptr(bias), ptr(out),
                                                        out.shape[3], int(c_offset), stream_ptr()),
          "sia_conv7x7_c3_relu_pool2_strided")
    return out


def conv3x3_relu_pool2(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, cout: int,
                       out: torch.Tensor | None = None) -> torch.Tensor:
    _need(x, torch.bfloat16, "x")
    _need(w_packed, torch.uint8, "w_packed")
    _need(bias, torch.float32, "bias")
    b, h, w, cin = x.shape
    if out is None:
        out = torch.empty((b, h // 2, w // 2, cout), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().sia_conv3x3_relu_pool2(ptr(x), b, h, w, cin, cout, ptr(w_packed), ptr(bias), ptr(out),
                                             stream_ptr()), "sia_conv3x3_relu_pool2")
    return out


def chw_to_hwc_u8(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[B,3,H,W] uint8 (a GPU JPEG decoder's planar output) -> [B,H,W,3] uint8 decode buffers for the transform kernels."""
    _need(x, torch.uint8, "x")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("x must be [B,3,H,W] uint8")
    b, _, h, w = x.shape
    if out is None:
        out = torch.empty((b, h, w, 3), dtype=torch.uint8, device=x.device)
    else:
        _need(out, torch.uint8, "out")
    check(_lib.load().sia_chw_u8_to_hwc_u8(ptr(x), b, h, w, ptr(out), stream_ptr()), "sia_chw_u8_to_hwc_u8")
    return out


def pad_nhwc(x: torch.Tensor, valid_hw, out_hw, out: torch.Tensor | None = None) -> torch.Tensor:
    """[B,h,w,C] bf16 -> [B,out_h,out_w,C] bf16 holding the valid corner of ``x``, zero elsewhere (floor pooling on
    odd sizes: the next conv kernel wants even sizes and zero padding beyond the valid corner)."""
    _need(x, torch.bfloat16, "x")
    b, h, w, c = x.shape
    if out is None:
        out = torch.empty((b, out_hw[0], out_hw[1], c), dtype=torch.bfloat16, device=x.device)
    else:
        _need(out, torch.bfloat16, "out")
    check(_lib.load().sia_pad_nhwc_bf16(ptr(x), b, h, w, c, int(valid_hw[0]), int(valid_hw[1]), ptr(out),
                                        int(out_hw[0]), int(out_hw[1]), stream_ptr()), "sia_pad_nhwc_bf16")
    return out


# -------------------------------------------------------------------------------------------------
# K5 / K6 linear part
# -------------------------------------------------------------------------------------------------
class TiledLinearWeight:
    """fc1 weights re-laid tile by tile (``sia_retile_linear_w``): ``tiles`` is a flat bf16 tensor of n*k elements."""

    def __init__(self, tiles: torch.Tensor, n: int, k: int):
        self.tiles, self.n, self.k = tiles, n, k
        self.shape = (n, k)
        self.device = tiles.device


def retile_linear_w(w: torch.Tensor) -> TiledLinearWeight:
    """bf16 [n, k] row-major -> TiledLinearWeight (every 128 x 64 tile contiguous, pre-swizzled)."""
    _need(w, torch.bfloat16, "w")
    n, k = w.shape
    tiles = torch.empty(n * k, dtype=torch.bfloat16, device=w.device)
    check(_lib.load().sia_retile_linear_w(ptr(w), n, k, ptr(tiles), stream_ptr()), "sia_retile_linear_w")
    return TiledLinearWeight(tiles, n, k)


def linear_splitk(a: torch.Tensor, w, splits: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """``w``: bf16 [n, k] tensor, or a TiledLinearWeight (one bulk copy per weight tile)."""
    _need(a, torch.bfloat16, "a")
    tiled = isinstance(w, TiledLinearWeight)
    _need(w.tiles if tiled else w, torch.bfloat16, "w")
    m, k = a.shape
    n, k2 = w.shape
    if k != k2:
        raise ValueError("inner dimensions differ")
    if out is None:
        out = torch.empty((splits, m, n), dtype=torch.float32, device=a.device)
    if tiled:
        check(_lib.load().sia_linear_splitk_tiled(ptr(a), ptr(w.tiles), m, n, k, splits, ptr(out), stream_ptr()),
              "sia_linear_splitk_tiled")
    else:
        check(_lib.load().sia_linear_splitk(ptr(a), ptr(w), m, n, k, splits, ptr(out), stream_ptr()),
              "sia_linear_splitk")
    return out


def head_tail(partial, b1, w2t, b2, w3, b3, label=None, groups=None, n_groups: int = 0, counts=None,
              logp=None, pred=None):
    """Returns (logp [M,2] f32, pred [M] u8); accumulates into ``counts`` when given."""
    _need(partial, torch.float32, "partial")
    splits, m, n1 = partial.shape
    n2 = w2t.shape[1]
    for t, nm in ((b1, "b1"), (w2t, "w2t"), (b2, "b2"), (w3, "w3"), (b3, "b3")):
        _need(t, torch.float32, nm)
    if logp is None:
        logp = torch.empty((m, 2), dtype=torch.float32, device=partial.device)
    if pred is None:
        pred = torch.empty((m,), dtype=torch.uint8, device=partial.device)
    n_attr, stride = 0, 0
    if counts is not None:
        _need(counts, torch.int64, "counts")
        _need(label, torch.uint8, "label")
        _need(groups, torch.uint8, "groups")
        n_attr, stride = groups.shape
        if tuple(counts.shape) != (n_attr, n_groups, 2, 2):
            raise ValueError("counts must be [n_attr, n_groups, 2, 2]")
    check(_lib.load().sia_head_tail(ptr(partial), splits, m, n1, n2, ptr(b1), ptr(w2t), ptr(b2), ptr(w3), ptr(b3),
                                    ptr(logp), ptr(pred), ptr(label), ptr(groups), stride, n_attr, n_groups,
                                    ptr(counts), stream_ptr()), "sia_head_tail")
    return logp, pred


def head_tail_chain(partial, n1: int, b1, layers, label=None, groups=None, n_groups: int = 0, counts=None,
                    logp=None, pred=None):
    """Tail for any number of Linear layers after fc1: ``layers`` = [(w_t [n_in, n_out] f32, b [n_out] f32), ...],
    ReLU after all but the last (2 classes).  ``partial`` is [splits, M, n1_stride] with n1 <= n1_stride."""
    _need(partial, torch.float32, "partial")
    splits, m, n1_stride = partial.shape
    _need(b1, torch.float32, "b1")
    for wt, b in layers:
        _need(wt, torch.float32, "w_t")
        _need(b, torch.float32, "b")
    if logp is None:
        logp = torch.empty((m, 2), dtype=torch.float32, device=partial.device)
    if pred is None:
        pred = torch.empty((m,), dtype=torch.uint8, device=partial.device)
    n_attr, stride = 0, 0
    if counts is not None:
        _need(counts, torch.int64, "counts")
        _need(label, torch.uint8, "label")
        _need(groups, torch.uint8, "groups")
        n_attr, stride = groups.shape
        if tuple(counts.shape) != (n_attr, n_groups, 2, 2):
            raise ValueError("counts must be [n_attr, n_groups, 2, 2]")
    k = len(layers)
    wt_arr = (ctypes.c_void_p * k)(*[wt.data_ptr() for wt, _ in layers])
    b_arr = (ctypes.c_void_p * k)(*[b.data_ptr() for _, b in layers])
    n_arr = (ctypes.c_int * k)(*[int(wt.shape[1]) for wt, _ in layers])
    check(_lib.load().sia_head_tail_chain(ptr(partial), splits, m, int(n1), n1_stride, ptr(b1), k, wt_arr, b_arr, n_arr,
                                          ptr(logp), ptr(pred), ptr(label), ptr(groups), stride, n_attr, n_groups,
                                          ptr(counts), stream_ptr()), "sia_head_tail_chain")
    return logp, pred


# -------------------------------------------------------------------------------------------------
# K7 confusion counts
# -------------------------------------------------------------------------------------------------
def confusion_counts(pred: torch.Tensor, label: torch.Tensor, groups: torch.Tensor, n_groups: int,
                     counts: torch.Tensor | None = None) -> torch.Tensor:
    """counts[a][g][label][pred] (int64), accumulated into ``counts`` if given.

    pred, label: uint8 [N] in {0,1} (1 = 'malignant'); groups: uint8 [A,N], values >= n_groups mean
    "in no group of this attribute" (the reference's filter() semantics, tone_bias_test.py:283-289).
    """
    _need(pred, torch.uint8, "pred")
    _need(label, torch.uint8, "label")
    _need(groups, torch.uint8, "groups")
    if groups.dim() != 2 or pred.dim() != 1 or label.shape != pred.shape or groups.shape[1] != pred.shape[0]:
        raise ValueError("expected pred [N], label [N], groups [A,N]")
    n_attr, n = groups.shape
    if counts is None:
        counts = torch.zeros((n_attr, n_groups, 2, 2), dtype=torch.int64, device=pred.device)
    else:
        _need(counts, torch.int64, "counts")
        if tuple(counts.shape) != (n_attr, n_groups, 2, 2):
            raise ValueError("counts must be [n_attr, n_groups, 2, 2]")
    check(_lib.load().sia_confusion_counts(ptr(pred), ptr(label), ptr(groups), n, n, n_attr, n_groups, ptr(counts),
                                           stream_ptr()), "sia_confusion_counts")
    return counts
