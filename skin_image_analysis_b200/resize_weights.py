"""Host-side construction of the banded resize operators consumed by ``sia_preprocess_u8hwc``.

``skimage.transform.resize(image, (h, w))`` with the defaults the reference uses
(tone_bias_dataset.py:425: order 1, mode 'reflect' -> ndimage 'mirror', anti-aliasing Gaussian with
sigma = max(0, (in/out - 1)/2) truncated at 4 sigma, half-pixel-centred zoom) is linear and
separable, so per axis it is a banded matrix  W[out, in]  = Zoom . Gauss  with the mirror boundary
folded into the band.  This module builds those bands in float64 and turns the vertical one into
the "row schedule" the streaming kernel executes (csrc/preprocess.cu).

Pure host arithmetic on a few hundred numbers; it runs once per (source size, output size).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

N_SLOTS = 4          # accumulator slots of the kernel's vertical pass (PRE_SLOTS)
SUPPORTED_X_TAPS = (8, 16)


def _mirror(idx: int, n: int) -> int:
    if n == 1:
        return 0
    period = 2 * (n - 1)
    idx = abs(idx) % period
    return period - idx if idx >= n else idx


def _gaussian(sigma: float):
    radius = int(4.0 * sigma + 0.5)
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    return radius, k / k.sum()


def axis_operator_rows(n_in: int, n_out: int, antialias: bool):
    """For each output sample: dict {source index: float64 weight} of the composed operator."""
    sigma = max(0.0, (n_in / n_out - 1.0) / 2.0) if antialias else 0.0
    if sigma > 1e-15:
        radius, kern = _gaussian(sigma)
    else:
        radius, kern = 0, np.ones(1)
    rows = []
    for i in range(n_out):
        cc = (i + 0.5) * (n_in / n_out) - 0.5 if n_in != n_out else float(i)
        if n_in > 1:                                  # fold the continuous coordinate (mirror)
            period = 2.0 * (n_in - 1)
            cc = abs(cc) % period
            if cc > n_in - 1:
                cc = period - cc
        else:
            cc = 0.0
        i0 = int(np.floor(cc))
        t = cc - i0
        row: dict[int, float] = {}
        for src, wz in ((i0, 1.0 - t), (_mirror(i0 + 1, n_in), t)):
            if wz == 0.0:
                continue
            for j in range(-radius, radius + 1):
                col = _mirror(src + j, n_in)
                row[col] = row.get(col, 0.0) + wz * kern[j + radius]
        rows.append(row)
    return rows


def axis_bands(n_in: int, n_out: int, antialias: bool, min_taps: int = 1):
    """(offset[n_out] int32, weights[n_out, taps] float64) with offset + taps <= n_in."""
    rows = axis_operator_rows(n_in, n_out, antialias)
    width = max(max(r) - min(r) + 1 for r in rows)
    taps = max(width, min_taps)
    if taps > n_in:
        raise ValueError(f"source axis of {n_in} samples is shorter than the {taps}-tap window")
    off = np.zeros(n_out, np.int32)
    w = np.zeros((n_out, taps), np.float64)
    for i, r in enumerate(rows):
        o = min(min(r), n_in - taps)
        off[i] = o
        for col, val in r.items():
            w[i, col - o] = val
    return off, w


X_FIXED_ONE = 1 << 15


def quantise_x_weights(x_w: np.ndarray) -> np.ndarray | None:
    """[out_w, 8] float64 -> [out_w, 4] uint32 of packed 15-bit fixed-point pairs (w[2i] | w[2i+1] << 16)
    for the IDP.2A horizontal pass.  Each row is renormalised to sum to exactly 2^15 (the largest weight
    absorbs the rounding residue), so a flat image stays exactly flat.  None if a weight is negative."""
    if x_w.shape[1] != 8 or np.any(x_w < 0):
        return None
    q = np.rint(x_w * X_FIXED_ONE).astype(np.int64)
    target = np.rint(x_w.sum(1) * X_FIXED_ONE).astype(np.int64)
    rows = np.arange(q.shape[0])
    q[rows, q.argmax(1)] += target - q.sum(1)
    if np.any(q < 0) or np.any(q > 0xFFFF):
        return None
    q = q.astype(np.uint32)
    return (q[:, 0::2] | (q[:, 1::2] << np.uint32(16))).astype(np.uint32)


@dataclass
class ResizeTables:
    src_h: int
    src_w: int
    out_h: int
    out_w: int
    x_taps: int
    x_off: np.ndarray          # int32 [out_w]
    x_w: np.ndarray            # float32 [out_w, x_taps]
    x_wq: np.ndarray | None    # uint32 [out_w, 4]: packed pairs of 15-bit weights (8-tap windows only)
    row_w: np.ndarray          # float32 [src_h, 4]
    row_emit: np.ndarray       # int32 [src_h, 4]
    y_first_last: np.ndarray   # int32 [out_h, 2]
    wy_dense: np.ndarray | None = None   # float64 [out_h, src_h]  (kept for tests)
    wx_dense: np.ndarray | None = None


def build_tables(src_h: int, src_w: int, out_h: int, out_w: int, scale: float = 1.0 / 255.0,
                 antialias: str | bool = "skimage", keep_dense: bool = False) -> ResizeTables:
    """``antialias='skimage'``: Gaussian pre-filter iff any axis shrinks (the reference's behaviour);
    ``False``: plain half-pixel bilinear."""
    if antialias == "skimage":
        aa = out_h < src_h or out_w < src_w
    else:
        aa = bool(antialias)
    # ---- horizontal band, padded to a tap count the kernel is instantiated for --------------------
    x_off, x_w = axis_bands(src_w, out_w, aa)
    x_taps = next((t for t in SUPPORTED_X_TAPS if t >= x_w.shape[1]), None)
    if x_taps is None:
        raise ValueError(f"horizontal window of {x_w.shape[1]} taps exceeds {SUPPORTED_X_TAPS[-1]} "
                         f"(scale factor {src_w / out_w:.2f} too large)")
    if x_taps > src_w:
        raise ValueError(f"source width {src_w} is smaller than the {x_taps}-tap window")
    x_off, x_w = axis_bands(src_w, out_w, aa, min_taps=x_taps)
    # ---- vertical band -> per-source-row schedule ---------------------------------------------------
    rows = axis_operator_rows(src_h, out_h, aa)
    row_w = np.zeros((src_h, N_SLOTS), np.float64)
    row_emit = np.full((src_h, N_SLOTS), -1, np.int32)
    owner = np.full((src_h, N_SLOTS), -1, np.int64)
    y_first_last = np.zeros((out_h, 2), np.int32)
    for i, r in enumerate(rows):
        first, last = min(r), max(r)
        y_first_last[i] = (first, last)
        slot = i % N_SLOTS
        if np.any(owner[first:last + 1, slot] >= 0):
            raise ValueError("more than 4 output rows in flight per source row: unsupported vertical scale "
                             f"({src_h} -> {out_h})")
        owner[first:last + 1, slot] = i
        for src, val in r.items():
            row_w[src, slot] = val * scale
        row_emit[last, slot] = i
    if np.any(np.diff(y_first_last[:, 0]) < 0) or np.any(np.diff(y_first_last[:, 1]) < 0):
        raise ValueError("vertical windows are not monotonic: unsupported geometry")
    t = ResizeTables(src_h, src_w, out_h, out_w, x_taps, x_off.astype(np.int32), x_w.astype(np.float32),
                     quantise_x_weights(x_w) if x_taps == 8 else None,
                     row_w.astype(np.float32), row_emit, y_first_last)
    if keep_dense:
        wy = np.zeros((out_h, src_h))
        for i, r in enumerate(rows):
            for c, v in r.items():
                wy[i, c] = v
        wx = np.zeros((out_w, src_w))
        for i in range(out_w):
            wx[i, x_off[i]:x_off[i] + x_w.shape[1]] = x_w[i]
        t.wy_dense, t.wx_dense = wy, wx
    return t


def rescale_size(h: int, w: int, output_size) -> tuple[int, int]:
    """Output size logic of ``Rescale`` (tone_bias_dataset.py:414-423)."""
    if isinstance(output_size, int):
        if h > w:
            new_h, new_w = output_size * h / w, output_size
        else:
            new_h, new_w = output_size, output_size * w / h
    else:
        new_h, new_w = output_size
    return int(new_h), int(new_w)
