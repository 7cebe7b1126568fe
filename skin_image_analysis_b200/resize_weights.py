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

import math

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


# -------------------------------------------------------------------------------------------------
# Tables of the tensor-core kernel (csrc/preprocess_tc.cu)
# -------------------------------------------------------------------------------------------------
TC_LANES = 128            # output rows per tile = TMEM lanes of one accumulator
TC_KWIN = 256             # source rows one tile may read (K of the vertical GEMM)
TC_BLOCK_STRIDE = 120     # image-row bytes between column blocks (40 pixels; multiple of 8)
TC_BLOCK_COLS = 128       # accumulator columns per block (the 8-byte overlap keeps every item in one block)
TC_ITEM_PX = 3            # source pixels one schedule item may consume
TC_ITEM_LOAD = 9          # accumulator columns one item loads (= 3 * TC_ITEM_PX)
TC_A_LBO, TC_A_SBO = 128, (TC_KWIN // 8) * 128     # K-major, no swizzle: 8 x 16-byte core matrices
NHWC4_PAD = 8             # SIA_NHWC4_PAD: extra pixel columns of the padded NHWC4 row


@dataclass
class TcTables:
    """Vertical operator as per-tile fp16 A operands (already in shared-memory order) + the horizontal
    operator as a static schedule of items.  Item i consumes up to 3 consecutive source pixels, adds them
    into the 4 accumulator slots with the item's weights, then emits output column ``emit`` (or nothing)
    from slot i % 4 and clears that slot."""
    n_tiles: int
    tile_rows: int
    tile_row0: np.ndarray      # int32 [n_tiles]: first source row of the tile's K window
    a_packed: np.ndarray       # uint8 [n_tiles, 128*256*2]: fp16 weights in UMMA K-major core-matrix order
    a_dense: np.ndarray        # float64 [n_tiles, 128, 256]: the same weights (fp16-rounded) for tests
    lane_scale: np.ndarray     # float32 [n_tiles, 128]: sum(w) / sum(fp16(w)) per output row (1 for unused lanes)
    items: np.ndarray          # float32-viewed [n_items, 16]: (col, emit, block, vector-store column of the group or -1) as int32, then 3 x 4 weights
    item_px0: np.ndarray       # int32 [n_items]: first source pixel of each item (tests / emulation)
    n_blocks: int
    last_block_cols: int       # MMA N of the last block (multiple of 16)
    pads_in_schedule: bool = False   # the vector-store groups also write the zero pad columns of the NHWC4 row

    @property
    def n_items(self) -> int:
        return self.items.shape[0]


def build_tc_tables(src_h: int, src_w: int, out_h: int, out_w: int, antialias: str | bool = "skimage") -> TcTables:
    """Raises ValueError when the geometry does not fit the tensor-core kernel (the caller then uses the
    CUDA-core kernel): K window > 256 rows per tile, an output column that is live for more than 4
    schedule items, a row length that is not a multiple of 8 bytes, or an odd number of source rows (the
    kernel's TMA view pairs rows)."""
    if antialias == "skimage":
        aa = out_h < src_h or out_w < src_w
    else:
        aa = bool(antialias)
    if (src_w * 3) % 8 != 0 or src_h % 2 != 0:
        raise ValueError("tensor-core preprocess needs src_w % 8 == 0 and an even src_h")
    n_tiles = -(-out_h // TC_LANES)
    tile_rows = -(-out_h // n_tiles)
    rows_y = axis_operator_rows(src_h, out_h, aa)
    tile_row0 = np.zeros(n_tiles, np.int32)
    a_dense = np.zeros((n_tiles, TC_LANES, TC_KWIN), np.float64)
    lane_scale = np.ones((n_tiles, TC_LANES), np.float32)
    for t in range(n_tiles):
        lo_i, hi_i = t * tile_rows, min(out_h, (t + 1) * tile_rows)
        r_lo = min(min(rows_y[i]) for i in range(lo_i, hi_i))
        r_hi = max(max(rows_y[i]) for i in range(lo_i, hi_i))
        if r_hi - r_lo + 1 > TC_KWIN:
            raise ValueError(f"vertical window of {r_hi - r_lo + 1} source rows per {tile_rows}-row tile exceeds {TC_KWIN}")
        tile_row0[t] = r_lo
        for i in range(lo_i, hi_i):
            exact = 0.0
            for r, v in rows_y[i].items():
                a_dense[t, i - lo_i, r - r_lo] = np.float64(np.float16(v))
                exact += v
            lane_scale[t, i - lo_i] = exact / a_dense[t, i - lo_i].sum()
    # shared-memory image of each A operand: (l, k) at (l/8)*SBO + (k/8)*LBO + (l%8)*16 + (k%8)*2
    a16 = a_dense.astype(np.float16)
    a_packed = np.zeros((n_tiles, TC_LANES * TC_KWIN), np.float16)
    l = np.arange(TC_LANES)[:, None]
    k = np.arange(TC_KWIN)[None, :]
    dst = ((l // 8) * TC_A_SBO + (k // 8) * TC_A_LBO + (l % 8) * 16 + (k % 8) * 2) // 2
    for t in range(n_tiles):
        a_packed[t, dst.reshape(-1)] = a16[t].reshape(-1)

    # ---- horizontal schedule ------------------------------------------------------------------------
    rows_x = axis_operator_rows(src_w, out_w, aa)
    first = [min(r) for r in rows_x]
    last = [max(r) for r in rows_x]
    if any(first[j] < first[j - 1] or last[j] < last[j - 1] for j in range(1, out_w)):
        raise ValueError("horizontal windows are not monotonic")
    px0, npx, emit = [], [], []
    cur = 0
    for j in range(out_w):
        n = min(last[j], src_w - 1) - cur + 1
        n = max(n, 0)                      # columns ending on an already consumed pixel (mirror folding): 0-px item
        while n > TC_ITEM_PX:
            px0.append(cur); npx.append(TC_ITEM_PX); emit.append(-1)
            cur += TC_ITEM_PX
            n -= TC_ITEM_PX
        px0.append(cur); npx.append(n); emit.append(j)
        cur += n
    # align the schedule so that output column j (padded column j + 1) is emitted by an item with index = j + 1
    # (mod 4): groups of 4 items then cover 4 aligned padded columns and can be stored as one 32-byte sector
    lead = next(i for i, e in enumerate(emit) if e >= 0)
    for _ in range((1 - lead) % 4):
        px0.insert(0, 0); npx.insert(0, 0); emit.insert(0, -1)
    while len(px0) % 8 != 0:                          # the kernel works on groups of 4 + 4 items: pad with no-ops
        px0.append(cur); npx.append(0); emit.append(-1)
    n_items = len(px0)
    item_of_col = {e: i for i, e in enumerate(emit) if e >= 0}
    item_of_px = np.zeros(src_w, np.int64)
    for i in range(n_items):
        item_of_px[px0[i]:px0[i] + npx[i]] = i
    w = np.zeros((n_items, TC_ITEM_PX, N_SLOTS), np.float64)
    for j, r in enumerate(rows_x):
        ij = item_of_col[j]
        for x, v in r.items():
            it = int(item_of_px[x])
            if x >= cur or not (ij - (N_SLOTS - 1) <= it <= ij):
                raise ValueError("an output column is live for more than 4 schedule items")
            w[it, x - px0[it], ij % N_SLOTS] += v
    row_bytes = src_w * 3
    items = np.zeros((n_items, 16), np.float32)
    info = items.view(np.int32)
    block = 0
    for i in range(n_items):
        if npx[i] > 0:
            block = (3 * px0[i]) // TC_BLOCK_STRIDE
            col = 3 * px0[i] - block * TC_BLOCK_STRIDE
        else:
            col = 0                        # loads nothing it uses; stay in the current block
        assert 0 <= col and col + TC_ITEM_LOAD <= TC_BLOCK_COLS
        info[i, 0:4] = (col, emit[i], block, -1)
        items[i, 4:16] = w[i].reshape(-1)
    # groups of 4 items whose pixels are 4 consecutive, 32-byte aligned columns of the padded NHWC4 row (pixel j
    # lives in column j + 1; an item that emits nothing stands for a zero pad column) are stored as one sector
    pad_cols = set([0] + list(range(out_w + 1, out_w + NHWC4_PAD)))
    covered = set()
    for g0 in range(0, n_items, 4):
        cols = [emit[g0 + u] + 1 if emit[g0 + u] >= 0 else None for u in range(4)]
        known = [(u, c) for u, c in enumerate(cols) if c is not None]
        if known:
            c0 = known[0][1] - known[0][0]
        else:                              # an all-pad group continues where the previous vector group ended
            c0 = info[g0 - 4, 3] + 4 if g0 >= 4 and info[g0 - 4, 3] >= 0 else -1
        ok = c0 >= 0 and c0 % 4 == 0 and c0 + 4 <= out_w + NHWC4_PAD
        for u in range(4):
            ok = ok and ((cols[u] == c0 + u) if cols[u] is not None else (c0 + u in pad_cols))
        if ok:
            info[g0, 3] = c0
            covered.update(c for c in range(c0, c0 + 4) if c in pad_cols)
    pads_in_schedule = covered == pad_cols
    n_blocks = block + 1
    last_cols = min(TC_BLOCK_COLS, -(-(row_bytes - (n_blocks - 1) * TC_BLOCK_STRIDE) // 16) * 16)
    return TcTables(n_tiles, tile_rows, tile_row0, a_packed.view(np.uint8).reshape(n_tiles, -1), a_dense, lane_scale,
                    items, np.asarray(px0, np.int32), n_blocks, last_cols, bool(pads_in_schedule))


def tc_emulate(u8: np.ndarray, t: TcTables, out_h: int, out_w: int, scale: float = 1.0 / 255.0) -> np.ndarray:
    """numpy model of the tensor-core kernel's arithmetic (fp16 vertical weights, exact products, fp32
    horizontal pass driven by the item schedule) -> [out_h, out_w, 3] float32.  Pins the tables on the CPU."""
    src_h, src_w, _ = u8.shape
    s = u8.reshape(src_h, src_w * 3).astype(np.float64)
    out = np.full((out_h, out_w, 3), np.nan, np.float32)
    info = t.items.view(np.int32)
    for tile in range(t.n_tiles):
        rows = np.clip(t.tile_row0[tile] + np.arange(TC_KWIN), 0, src_h - 1)
        v = (t.a_dense[tile] @ s[rows]).astype(np.float32)              # [128, row_bytes]
        v = np.concatenate([v, np.zeros((TC_LANES, 3 * TC_ITEM_PX), np.float32)], axis=1)
        acc = np.zeros((TC_LANES, N_SLOTS, 3), np.float32)
        lanes = min(t.tile_rows, out_h - tile * t.tile_rows)
        for i in range(t.n_items):
            col, emit, block, _n = (int(z) for z in info[i, :4])
            base = block * TC_BLOCK_STRIDE + col
            wts = t.items[i, 4:16].reshape(TC_ITEM_PX, N_SLOTS)
            for kk in range(TC_ITEM_PX):
                px = v[:, base + 3 * kk: base + 3 * kk + 3]
                acc += wts[kk][None, :, None] * px[:, None, :]
            if emit >= 0:
                o = acc[:lanes, i % N_SLOTS] * (t.lane_scale[tile, :lanes, None] * np.float32(scale))
                out[tile * t.tile_rows: tile * t.tile_rows + lanes, emit] = o
            acc[:, i % N_SLOTS] = 0
    return out


# -------------------------------------------------------------------------------------------------
# Tables of the torchvision-variant kernel (csrc/preprocess_tv.cu): ATen's fixed-point antialias resampler
# -------------------------------------------------------------------------------------------------
@dataclass
class TvAxis:
    xmin: np.ndarray        # int32 [n_out]: first source index of every output sample's window
    w: np.ndarray           # int16 [n_out, taps]: taps scaled by 2^precision (zero padded)
    taps: int
    precision: int


def tv_axis(n_in: int, n_out: int) -> TvAxis:
    """One axis of ``torch.nn.functional.interpolate(mode="bilinear", antialias=True)`` on uint8 -- the
    resampler behind ``v2.Resize`` (notebooks/ToneClassifier/CNNTrialDataset.py:71).  Formulas of ATen's
    ``_compute_indices_min_size_weights_aa`` / ``_compute_index_ranges_int16_weights``
    (aten/src/ATen/native/cpu/UpSampleKernel.cpp): triangle filter of half width max(scale, 1) centred on
    scale * (i + 0.5), weights normalised in float64, then rounded half away from zero to int16 at the largest
    precision that keeps the biggest weight below 2^15.  An axis that keeps its size is the identity (ATen skips
    the pass).  Windows that would run past the last sample are shifted left and zero padded in front so that
    xmin + taps <= n_in always holds (the kernel then needs no bounds checks)."""
    if n_in < 1 or n_out < 1:
        raise ValueError("sizes must be positive")
    if n_in == n_out:
        return TvAxis(np.arange(n_out, dtype=np.int32), np.full((n_out, 1), 1 << 14, np.int16), 1, 14)
    scale = n_in / n_out
    support = scale if scale >= 1.0 else 1.0
    taps = int(math.ceil(support)) * 2 + 1
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    xmin = np.zeros(n_out, np.int32)
    w = np.zeros((n_out, taps), np.float64)
    sizes = np.zeros(n_out, np.int64)
    wt_max = 0.0
    for i in range(n_out):
        centre = scale * (i + 0.5)
        lo = max(int(centre - support + 0.5), 0)
        size = min(max(min(int(centre + support + 0.5), n_in) - lo, 0), taps)
        vals = [max(0.0, 1.0 - abs((j + lo - centre + 0.5) * invscale)) for j in range(size)]
        total = 0.0
        for v in vals:
            total += v
        if total != 0.0:
            for j in range(size):
                w[i, j] = vals[j] / total
            wt_max = max(wt_max, max(w[i, :size]))
        xmin[i], sizes[i] = lo, size
    prec = 0
    while prec < 22 and int(0.5 + wt_max * (1 << (prec + 1))) < (1 << 15):
        prec += 1
    v = w * float(1 << prec)
    wi = np.where(v < 0, np.trunc(v - 0.5), np.trunc(v + 0.5)).astype(np.int16)
    if taps > n_in:                                   # tiny sources: the window is the whole axis
        wi = wi[:, :n_in].copy()
        taps = n_in
    for i in range(n_out):
        over = int(xmin[i]) + taps - n_in
        if over > 0:
            if np.any(wi[i, taps - over:] != 0):
                raise ValueError("internal: non-zero taps past the edge")
            wi[i] = np.concatenate([np.zeros(over, np.int16), wi[i, :taps - over]])
            xmin[i] -= over
    return TvAxis(xmin, np.ascontiguousarray(wi), taps, prec)


def tv_normalise_lut(mean, std) -> np.ndarray:
    """float32 [3, 256]: ``v2.ToDtype(float32, scale=True)`` then ``v2.Normalize(mean, std)`` of byte b in channel
    c, with the same float32 operations (CNNTrialDataset.py:72-73), so a table lookup is bit-identical."""
    b = np.arange(256, dtype=np.float32) * np.float32(1.0 / 255.0)
    m = np.asarray(mean, np.float32).reshape(3, 1)
    s = np.asarray(std, np.float32).reshape(3, 1)
    return ((b[None, :] - m) / s).astype(np.float32)


def tv_tile_plan(y: TvAxis, row_bytes: int, out_w: int, x_taps: int, budget: int = 100 * 1024) -> tuple[int, int]:
    """(tile_rows, max_window_rows) of the kernel launch: the largest tile of 32/16/8/4/2/1 output rows whose
    source window (plus the horizontal-pass buffer and tables) fits ``budget`` bytes of shared memory, so that
    two CTAs share an SM; falls back to the 227 KB limit before giving up."""
    n_out = y.xmin.shape[0]
    hpitch = (out_w * 3 + 3) & ~3
    for limit in (budget, 227 * 1024):
        for tile in (32, 16, 8, 4, 2, 1):
            rows = 0
            for i0 in range(0, n_out, tile):
                i1 = min(i0 + tile, n_out)
                rows = max(rows, int(y.xmin[i1 - 1]) + y.taps - int(y.xmin[i0]))
            # source window: the larger of the interleaved and the planar (three separately aligned planes) layouts
            window = max((rows * row_bytes + 32 + 15) & ~15, 3 * ((rows * (row_bytes // 3) + 31) & ~15))
            smem = window + ((rows * hpitch + 15) & ~15) + 768 * 4 + out_w * 4 + out_w * x_taps * 2 + 16
            if smem <= limit:
                return tile, rows
    raise ValueError("source rows too long for the shared-memory window of the torchvision-variant kernel")


# -------------------------------------------------------------------------------------------------
# Tables of the two-product tensor-core kernel (csrc/preprocess_tc2.cu): horizontal pass as a second UMMA
# -------------------------------------------------------------------------------------------------
TC2_SLOTS_IN, TC2_SLOTS_FULL, TC2_SLOTS_OUT = 4, 13, 4      # output-pixel slots of one column block
TC2_SLOTS = TC2_SLOTS_IN + TC2_SLOTS_FULL + TC2_SLOTS_OUT   # 21 pixels x 3 channels = 63 of the 64 UMMA columns
TC2_N = 64
TC2_B_BYTES = TC2_N * TC_BLOCK_COLS * 2                     # one block's slice of Wx, fp16, K-major core matrices


@dataclass
class Tc2Tables:
    """Vertical operator exactly as in TcTables; horizontal operator as one fp16 matrix per 120-byte column block.

    Block b owns the source bytes [120 b, 120 b + 120) of the image row (the last block: the rest) = whole pixels.
    Every block has 21 output-pixel slots (x 3 channels = 63 of the 64 UMMA columns).  An output pixel whose taps
    straddle blocks b and b + 1 sits in slot 17 + t of block b and in slot t of block b + 1 (t = 0..3, both sets
    right-aligned: the kernel carries slots 17..20 into slots 0..3 of the next block in registers); one whose taps
    all lie in block b sits after the block's incoming pixels.  The pixels a block COMPLETES are therefore the
    contiguous slots [s_lo, s_hi), in output-column order starting at column j_lo -- which lets the kernel store them
    as aligned 32-byte groups.  ``b2[b]`` is the [64 x 128] matrix (row = 3 * slot + channel, column = byte inside
    the block) in UMMA K-major core-matrix order; ``slot_scale[b, s]`` = sum(w) / sum(fp16(w)) of the column slot s
    completes."""
    vert: TcTables
    b2: np.ndarray            # uint8 [n_blocks, TC2_B_BYTES]
    b2_dense: np.ndarray      # float64 [n_blocks, 64, 128] (fp16-rounded), for tests / emulation
    block_meta: np.ndarray    # int32 [n_blocks, 4]: s_lo, s_hi, j_lo, 0
    slot_col: np.ndarray      # int32 [n_blocks, 21]: output column completed by the slot, -1 = none (tests / emulation)
    slot_scale: np.ndarray    # float32 [n_blocks, 21]
    n_blocks: int
    last_block_cols: int


def build_tc2_tables(src_h: int, src_w: int, out_h: int, out_w: int, antialias: str | bool = "skimage") -> Tc2Tables:
    """Raises ValueError when the geometry does not fit (the caller falls back to the one-product kernel)."""
    vert = build_tc_tables(src_h, src_w, out_h, out_w, antialias)
    if out_w % 4 != 0:
        raise ValueError("the two-product kernel stores whole groups of four output columns: out_w % 4 must be 0")
    if antialias == "skimage":
        aa = out_h < src_h or out_w < src_w
    else:
        aa = bool(antialias)
    rows_x = axis_operator_rows(src_w, out_w, aa)
    n_blocks = vert.n_blocks
    px_per_block = TC_BLOCK_STRIDE // 3

    def owner(px: int) -> int:
        return min(px // px_per_block, n_blocks - 1)

    blocks_of = [sorted({owner(x) for x in r}) for r in rows_x]
    if any(len(b) > 2 or (len(b) == 2 and b[1] != b[0] + 1) for b in blocks_of):
        raise ValueError("an output column spans more than two column blocks")
    b2_dense = np.zeros((n_blocks, TC2_N, TC_BLOCK_COLS), np.float64)
    block_meta = np.zeros((n_blocks, 4), np.int32)
    slot_col = np.full((n_blocks, TC2_SLOTS), -1, np.int32)
    slot_scale = np.ones((n_blocks, TC2_SLOTS), np.float32)
    slot_of = {}                                           # (block, output column) -> slot
    first_out = TC2_SLOTS_IN + TC2_SLOTS_FULL              # 17
    for b in range(n_blocks):
        ins = [j for j in range(out_w) if blocks_of[j] == [b - 1, b]]
        full = [j for j in range(out_w) if blocks_of[j] == [b]]
        outs = [j for j in range(out_w) if blocks_of[j] == [b, b + 1]]
        if len(ins) > TC2_SLOTS_IN or len(outs) > TC2_SLOTS_OUT:
            raise ValueError("too many straddling output columns per column block")
        done = ins + full                                  # completed here, in column order
        if done != list(range(done[0], done[0] + len(done))) if done else False:
            raise ValueError("the columns a block completes are not contiguous")
        s_lo = TC2_SLOTS_IN - len(ins)
        if s_lo + len(done) > first_out:                   # more than 13 complete columns
            extra = s_lo + len(done) - first_out
            if not ins and extra <= s_lo:                  # no incoming pixels: start further left
                s_lo -= extra
            elif not outs and b == n_blocks - 1 and s_lo + len(done) <= TC2_SLOTS:
                pass                                       # last block: run into the unused "out" slots
            else:
                raise ValueError("too many output columns per column block for the slot layout")
        for q, j in enumerate(done):
            slot_of[(b, j)] = s_lo + q
            slot_col[b, s_lo + q] = j
        for q, j in enumerate(outs):
            slot_of[(b, j)] = TC2_SLOTS - len(outs) + q
        block_meta[b] = (s_lo, s_lo + len(done), done[0] if done else 0, 0)
    # slot 17 + t of block b feeds slot t of block b + 1
    for b in range(n_blocks - 1):
        for j in [j for j in range(out_w) if blocks_of[j] == [b, b + 1]]:
            assert slot_of[(b, j)] - first_out == slot_of[(b + 1, j)]
    for j, r in enumerate(rows_x):
        exact = sum(r.values())
        s16 = sum(float(np.float16(v)) for v in r.values())
        for x, v in r.items():
            b = owner(x)
            s = slot_of[(b, j)]
            k0 = 3 * x - b * TC_BLOCK_STRIDE
            if k0 + 3 > TC_BLOCK_COLS:
                raise ValueError("a source pixel falls outside its block's accumulator columns")
            for c in range(3):
                b2_dense[b, 3 * s + c, k0 + c] += float(np.float16(v))
        b_done = blocks_of[j][-1]
        slot_scale[b_done, slot_of[(b_done, j)]] = exact / s16
    # shared-memory image: (n, k) at (n/8)*SBO + (k/8)*LBO + (n%8)*16 + (k%8)*2, LBO = 128, SBO = 16 * 128
    n = np.arange(TC2_N)[:, None]
    k = np.arange(TC_BLOCK_COLS)[None, :]
    dst = ((n // 8) * (TC_BLOCK_COLS // 8) * 128 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2) // 2
    b2 = np.zeros((n_blocks, TC2_N * TC_BLOCK_COLS), np.float16)
    for b in range(n_blocks):
        b2[b, dst.reshape(-1)] = b2_dense[b].astype(np.float16).reshape(-1)
    return Tc2Tables(vert, b2.view(np.uint8).reshape(n_blocks, -1), b2_dense, block_meta, slot_col, slot_scale,
                     n_blocks, vert.last_block_cols)


def tc2_emulate(u8: np.ndarray, t: Tc2Tables, out_h: int, out_w: int, scale: float = 1.0 / 255.0) -> np.ndarray:
    """numpy model of the two-product kernel's arithmetic: fp16 vertical weights, fp32 accumulate, lane scale,
    V rounded to fp16, fp16 horizontal weights, fp32 accumulate, carry across blocks -> [out_h, out_w, 3] float32."""
    v = t.vert
    src_h, src_w, _ = u8.shape
    s = u8.reshape(src_h, src_w * 3).astype(np.float64)
    out = np.full((out_h, out_w, 3), np.nan, np.float32)
    for tile in range(v.n_tiles):
        rows = np.clip(v.tile_row0[tile] + np.arange(TC_KWIN), 0, src_h - 1)
        vv = (v.a_dense[tile] @ s[rows]).astype(np.float32) * v.lane_scale[tile][:, None]
        vv = vv.astype(np.float16).astype(np.float64)
        vv = np.concatenate([vv, np.zeros((TC_LANES, TC_BLOCK_COLS), np.float64)], axis=1)
        lanes = min(v.tile_rows, out_h - tile * v.tile_rows)
        carry = np.zeros((TC_LANES, TC2_SLOTS_OUT * 3), np.float32)
        for b in range(t.n_blocks):
            blk = vv[:, b * TC_BLOCK_STRIDE: b * TC_BLOCK_STRIDE + TC_BLOCK_COLS]
            d2 = (blk @ t.b2_dense[b].T).astype(np.float32)                      # [128, 64]
            d2[:, :TC2_SLOTS_IN * 3] += carry
            carry = d2[:, (TC2_SLOTS_IN + TC2_SLOTS_FULL) * 3: TC2_SLOTS * 3].copy()
            for sl in range(TC2_SLOTS):
                j = int(t.slot_col[b, sl])
                if j >= 0:
                    o = d2[:lanes, 3 * sl: 3 * sl + 3] * (t.slot_scale[b, sl] * np.float32(scale))
                    out[tile * v.tile_rows: tile * v.tile_rows + lanes, j] = o
    return out


# -------------------------------------------------------------------------------------------------
# Tables of the warp-MMA kernel (csrc/preprocess_mma.cu)
# -------------------------------------------------------------------------------------------------
MMA_ROWS = 16             # output rows of one m-step = M of mma.sync.m16n8k16
MMA_GROUP_PX = 32         # source pixels of one column group: 8 lane groups x 4 pixels (12 bytes) each
MMA_WY_SHIFT = 15         # vertical weights are stored as fp16(w * 2^15); the image bytes enter as fp16 subnormals b * 2^-24
MMA_OUT_SCALE = 2.0 ** (24 - MMA_WY_SHIFT)        # what the kernel multiplies its accumulators by (besides scale / std)
# which row of a 16-row chunk lane quad-index q reads for the four K slots c = 0..3 (k = 2q, 2q+1, 2q+8, 2q+9):
# "natural": 2q + (0, 1, 8, 9);  "spread": 4q + (0, 1, 2, 3) -- chosen per row pitch so that the 32 lanes of one
# shared-memory load hit 32 different banks (600-pixel rows: 450 words = 2 (mod 32) -> "spread" is conflict free)
MMA_ROW_MAPS = {"natural": (2, (0, 1, 8, 9)), "spread": (4, (0, 1, 2, 3))}


def mma_row_map(row_bytes: int) -> str:
    """The mapping with fewer shared-memory bank conflicts for rows of ``row_bytes`` bytes stored back to back."""
    words = row_bytes // 4
    best, best_cost = "natural", None
    for name, (qs, cs) in MMA_ROW_MAPS.items():
        cost = 0
        for c in cs:
            banks = [((qs * q + c) * words + 3 * g) % 32 for q in range(4) for g in range(8)]
            cost += len(banks) - len(set(banks))
        if best_cost is None or cost < best_cost:
            best, best_cost = name, cost
    return best


@dataclass
class MmaTables:
    """Everything ``sia_preprocess_mma_u8hwc`` needs for one (source size, output size).

    The resize is two banded products run on the warp-level tensor-core path (mma.sync.m16n8k16, fp16 in / fp32
    accumulate), chained in registers:

        V[i, k]    = sum_r Wy[i, r] S[r, k]        M = 16 output rows, N = 8 image-byte columns, K = 16*kv source rows
        out[i, j]  = sum_x Wx[j, x] V[i, x, c]     M = the same 16 rows, N = 8 output pixels of channel c, K = 16 pixels

    * one m-step = 16 output rows; it reads the source rows [r0[m], r0[m] + 16 kv), r0 a multiple of 8.
    * the source row is cut into groups of 32 pixels (96 bytes); lane group g (= lane / 4) of a warp owns the 12 bytes
      [96 grp + 12 g, +12): V n-tile b (0..11) holds byte b of every lane group, i.e. its 8 columns are the pixels
      4 g + b / 3 of channel b % 3 -- which makes the accumulators of the first product, two n-tiles at a time,
      exactly the A operand of the second one (K index k < 8: pixel 4 k + 2 X, k >= 8: pixel 4 (k - 8) + 2 X + 1 for
      the K chunk X in {0, 1}); the de-interleaving of the channels costs nothing.
    * output n-tile t = the padded columns [8 t, 8 t + 8) of the NHWC4 row (pixel j sits in column j + 1, pad columns
      carry zero weights and come out as zeros).  A tile reads at most two adjacent groups; it is computed when its
      LAST group has been processed, from the V fragments of that group ("cur") and of the one before ("prev"):
      the tiles finished by group grp are [tile_begin[grp], tile_begin[grp + 1]).
    * the weights are rounded to fp16 individually (every weight keeps 11 significant bits, so sparse / dark images
      keep their relative accuracy); the row and column sums are then 1 +- 3e-4, which is NOT renormalised: a flat
      image comes out within 6e-4 (relative) of flat before the bf16 rounding, i.e. within one bf16 ulp.  The only
      scale left in the kernel is the constant ``MMA_OUT_SCALE``.
    """
    kv: int
    n_msteps: int
    n_groups: int
    n_tiles: int
    row_map: str
    r0: np.ndarray            # int32 [n_msteps]
    wy_frag: np.ndarray       # uint32 [n_msteps, kv, 32, 4]     A fragments of the vertical operator
    wx_frag: np.ndarray       # uint32 [n_tiles, 2, 2, 32, 2]    B fragments of the horizontal operator [tile][prev/cur][X]
    wx_mask: np.ndarray       # uint32 [n_tiles]                 bit (rel * 2 + X): that fragment has a non-zero weight
    tile_begin: np.ndarray    # int32 [n_groups + 1]
    wy16: np.ndarray          # float64 [n_msteps * 16, src_h]   fp16-rounded vertical weights (unshifted), for emulation
    wx16: np.ndarray          # float64 [out_w + 8, src_w]       fp16-rounded horizontal weights per padded column


def _f16_bits(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, np.float16).view(np.uint16).astype(np.uint32)


def build_mma_tables(src_h: int, src_w: int, out_h: int, out_w: int, antialias: str | bool = "skimage") -> MmaTables:
    """Raises ValueError when the geometry does not fit the kernel (the caller falls back to another kernel)."""
    if src_w % 8 != 0 or src_h % 2 != 0:
        raise ValueError("the warp-MMA kernel copies 16-byte aligned octets of rows: src_w % 8 and src_h % 2 must be 0")
    if out_w % 8 != 0:
        raise ValueError("the warp-MMA kernel writes n-tiles of 8 padded columns: out_w % 8 must be 0")
    if antialias == "skimage":
        aa = out_h < src_h or out_w < src_w
    else:
        aa = bool(antialias)
    rows_y = axis_operator_rows(src_h, out_h, aa)
    rows_x = axis_operator_rows(src_w, out_w, aa)
    # ---------------- vertical operator: windows and A fragments ------------------------------------------------------
    n_msteps = -(-out_h // MMA_ROWS)
    r0 = np.zeros(n_msteps, np.int32)
    span = 0
    for m in range(n_msteps):
        rs = rows_y[m * MMA_ROWS: (m + 1) * MMA_ROWS]
        first, last = min(min(r) for r in rs), max(max(r) for r in rs)
        r0[m] = first // 8 * 8
        span = max(span, last - r0[m] + 1)
    kv = -(-span // 16)
    if kv > 3:
        raise ValueError(f"16 output rows read {span} source rows: more than the 48-row window of the warp-MMA kernel")
    kv = max(kv, 2)
    if np.any(np.diff(r0) < 0):
        raise ValueError("vertical windows are not monotonic")
    wy16 = np.zeros((n_msteps * MMA_ROWS, src_h), np.float64)
    a_dense = np.zeros((n_msteps, MMA_ROWS, 16 * kv), np.float64)           # fp16(w * 2^15), window-relative columns
    for i, r in enumerate(rows_y):
        m = i // MMA_ROWS
        if max(r.values()) * 2.0 ** MMA_WY_SHIFT > 65504.0:
            raise ValueError("vertical weight overflows fp16")
        q = {c: float(np.float16(v * 2.0 ** MMA_WY_SHIFT)) for c, v in r.items()}     # round to nearest: every weight
        for c, v in q.items():                                                      # keeps 11 significant bits
            a_dense[m, i % MMA_ROWS, c - r0[m]] = v
            wy16[i, c] = v * 2.0 ** -MMA_WY_SHIFT
    row_map = mma_row_map(3 * src_w)
    qs, cs = MMA_ROW_MAPS[row_map]
    lane = np.arange(32)
    g, q4 = lane // 4, lane % 4
    wy_frag = np.zeros((n_msteps, kv, 32, 4), np.uint32)
    for kc in range(kv):
        blk = a_dense[:, :, 16 * kc: 16 * kc + 16]                             # [m, row, chunk row]
        # A fragment register reg holds (row, logical k) = a0 (g, 2q / 2q+1), a1 (g+8, same), a2 (g, 2q+8 / 2q+9), a3;
        # logical k = 2q + (c & 1) + 8 (c >> 1) reads chunk row qs * q + cs[c]
        for reg, (rr, c_lo) in enumerate(((g, 0), (g + 8, 0), (g, 2), (g + 8, 2))):
            lo = _f16_bits(blk[:, rr, qs * q4 + cs[c_lo]])
            hi = _f16_bits(blk[:, rr, qs * q4 + cs[c_lo + 1]])
            wy_frag[:, kc, :, reg] = lo | (hi << np.uint32(16))
    # ---------------- horizontal operator: n-tiles and B fragments ----------------------------------------------------
    n_groups = -(-src_w // MMA_GROUP_PX)
    n_tiles = (out_w + NHWC4_PAD) // 8
    wx16 = np.zeros((out_w + NHWC4_PAD, src_w), np.float64)
    for j, r in enumerate(rows_x):
        for c, v in r.items():
            wx16[j + 1, c] = float(np.float16(v))
    tile_first, tile_last = np.zeros(n_tiles, np.int32), np.zeros(n_tiles, np.int32)
    for t in range(n_tiles):
        cols = np.nonzero(wx16[8 * t: 8 * t + 8].any(0))[0]
        if len(cols) == 0:
            raise ValueError("an output n-tile without any real pixel")
        tile_first[t], tile_last[t] = int(cols.min()) // MMA_GROUP_PX, int(cols.max()) // MMA_GROUP_PX
    if np.any(tile_last - tile_first > 1):
        raise ValueError("an output n-tile reads more than two column groups")
    if np.any(np.diff(tile_last) < 0):
        raise ValueError("horizontal windows are not monotonic")
    tile_begin = np.array([int(np.searchsorted(tile_last, grp, side="left")) for grp in range(n_groups + 1)], np.int32)
    wx_frag = np.zeros((n_tiles, 2, 2, 32, 2), np.uint32)
    wx_mask = np.zeros(n_tiles, np.uint32)
    k_idx = np.stack([2 * q4, 2 * q4 + 1, 2 * q4 + 8, 2 * q4 + 9], 1)         # [lane, 4] logical k of b0.lo, b0.hi, b1.lo, b1.hi
    for t in range(n_tiles):
        pc = 8 * t + g                                                        # padded column of n = g
        for rel in range(2):
            grp = int(tile_last[t]) - 1 + rel
            if grp < 0:
                continue
            for x_chunk in range(2):
                px = 32 * grp + 4 * (k_idx % 8) + 2 * x_chunk + (k_idx // 8)     # source pixel of logical k
                w = np.where(px < src_w, wx16[pc[:, None], np.minimum(px, src_w - 1)], 0.0)
                if np.any(w != 0.0):
                    wx_mask[t] |= np.uint32(1 << (rel * 2 + x_chunk))
                bits = _f16_bits(w)
                wx_frag[t, rel, x_chunk, :, 0] = bits[:, 0] | (bits[:, 1] << np.uint32(16))
                wx_frag[t, rel, x_chunk, :, 1] = bits[:, 2] | (bits[:, 3] << np.uint32(16))
    return MmaTables(kv, n_msteps, n_groups, n_tiles, row_map, r0, wy_frag, wx_frag, wx_mask, tile_begin, wy16, wx16)


def mma_emulate(u8: np.ndarray, t: MmaTables, out_h: int, out_w: int, scale: float = 1.0 / 255.0) -> np.ndarray:
    """numpy model of the warp-MMA kernel's arithmetic (fp16 weights, fp32 accumulate, V rounded once to fp16) -> [out_h, out_w, 3] float32, before mean / std and the bf16 rounding."""
    src_h, src_w, _ = u8.shape
    s = u8.reshape(src_h, src_w * 3).astype(np.float64)
    v = (t.wy16[:out_h] * 2.0 ** MMA_WY_SHIFT) @ (s * 2.0 ** -24)              # what the first product accumulates
    v16 = v.astype(np.float32).astype(np.float16).astype(np.float64).reshape(out_h, src_w, 3)
    h = np.einsum("jx,ixc->ijc", t.wx16[1: out_w + 1], v16).astype(np.float32)
    return h * (np.float32(MMA_OUT_SCALE) * np.float32(scale))
