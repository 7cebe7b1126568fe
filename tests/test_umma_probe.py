"""Hardware bring-up of the tcgen05 descriptor conventions the conv / linear kernels rely on.

Each experiment writes a shared-memory image exactly as TMA (or the weight packer) would, issues
tcgen05.mma through ``sia_debug_umma_probe`` and compares the 128 x N accumulator with numpy.
Inputs are small integers, so the fp32 result is exact and the comparison is ``==``.

REQUIRED experiments are the layouts the shipped kernels use; the others are exploratory (future
single-halo-copy conv design) and only recorded in gpurun_out/probe_report.json.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SW_NONE, SW_128, SW_64 = 0, 2, 4


def desc(addr, lbo, sbo, layout, base_offset=0):
    return ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) \
        | ((base_offset & 7) << 49) | ((layout & 7) << 61)


def bf16_bits(a):
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def rand_int(rng, shape, lo=-3, hi=4):
    return rng.integers(lo, hi, shape).astype(np.float32)


class Image:
    def __init__(self, nbytes):
        self.buf = np.zeros(nbytes // 2, np.uint16)

    def put_rows_swizzled(self, base, mat, row_bytes, first_row=0):
        """mat [R, row_bytes/2]: row r goes to smem row (first_row + r); 16-byte units XOR-swizzled by the
        ABSOLUTE smem row index, as TMA SWIZZLE_128B / 64B writes them (base aligned to 1024)."""
        bits = bf16_bits(mat)
        units = row_bytes // 16
        for r in range(mat.shape[0]):
            row = first_row + r
            x = (row % 8) if row_bytes == 128 else ((row >> 1) & 3)
            for u in range(units):
                dst = (base + row * row_bytes + ((u ^ x) * 16)) // 2
                self.buf[dst:dst + 8] = bits[r, u * 8:(u + 1) * 8]

    def put_core_matrices(self, base, mat, lbo, sbo):
        """no-swizzle K-major: (r,k) at (r//8)*sbo + (k//8)*lbo + (r%8)*16 + (k%8)*2."""
        bits = bf16_bits(mat)
        for r in range(mat.shape[0]):
            for kc in range(mat.shape[1] // 8):
                dst = (base + (r // 8) * sbo + kc * lbo + (r % 8) * 16) // 2
                self.buf[dst:dst + 8] = bits[r, kc * 8:(kc + 1) * 8]

    def put_raw(self, base, mat):
        bits = bf16_bits(mat).reshape(-1)
        self.buf[base // 2: base // 2 + bits.size] = bits

    def tensor(self):
        return torch.from_numpy(self.buf.view(np.uint8).copy()).cuda()


def run(img, a_descs, b_descs, n):
    from skin_image_analysis_b200 import debug_probes as probes
    return probes.umma_probe(img.tensor(), a_descs, b_descs, n).cpu().numpy()


REPORT = {}


def record(name, ok, required, extra=None):
    """Merged into gpurun_out/probe_report.json (each experiment may run in its own process)."""
    os.makedirs("gpurun_out", exist_ok=True)
    path = "gpurun_out/probe_report.json"
    try:
        with open(path) as f:
            REPORT.update(json.load(f))
    except Exception:
        pass
    REPORT[name] = {"ok": bool(ok), "required": required, **(extra or {})}
    with open(path, "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


@pytest.mark.parametrize("n", [64, 128, 256, 32])
def test_sw128_canonical(n):
    rng = np.random.default_rng(n)
    a, b = rand_int(rng, (128, 64)), rand_int(rng, (n, 64))
    img = Image(16384 + n * 128)
    img.put_rows_swizzled(0, a, 128)
    img.put_rows_swizzled(16384, b, 128)
    ad = [desc(kk * 32, 0, 1024, SW_128) for kk in range(4)]
    bd = [desc(16384 + kk * 32, 0, 1024, SW_128) for kk in range(4)]
    got = run(img, ad, bd, n)
    ok = np.array_equal(got, a @ b.T)
    record(f"sw128_canonical_n{n}", ok, True, {"maxerr": float(np.abs(got - a @ b.T).max())})
    assert ok


def test_sw64_canonical():
    rng = np.random.default_rng(1)
    a, b = rand_int(rng, (128, 32)), rand_int(rng, (64, 32))
    img = Image(8192 + 4096)
    img.put_rows_swizzled(0, a, 64)
    img.put_rows_swizzled(8192, b, 64)
    ad = [desc(kk * 32, 0, 512, SW_64) for kk in range(2)]
    bd = [desc(8192 + kk * 32, 0, 512, SW_64) for kk in range(2)]
    got = run(img, ad, bd, 64)
    ok = np.array_equal(got, a @ b.T)
    record("sw64_canonical", ok, True, {"maxerr": float(np.abs(got - a @ b.T).max())})
    assert ok


@pytest.mark.parametrize("r", [1, 2])
def test_swizzled_vertical_tap_shift(r):
    """conv3x3: the halo buffer is 144 smem rows; tap r starts r*8 rows (whole atoms) further."""
    rng = np.random.default_rng(10 + r)
    halo, b = rand_int(rng, (144, 64)), rand_int(rng, (128, 64))
    img = Image(18432 + 16384)
    img.put_rows_swizzled(0, halo, 128)
    img.put_rows_swizzled(18432, b, 128)
    ad = [desc(r * 1024 + kk * 32, 0, 1024, SW_128) for kk in range(4)]
    bd = [desc(18432 + kk * 32, 0, 1024, SW_128) for kk in range(4)]
    got = run(img, ad, bd, 128)
    want = halo[r * 8: r * 8 + 128] @ b.T
    ok = np.array_equal(got, want)
    record(f"sw128_tap_shift_r{r}", ok, True)
    # the same for 64-byte rows (conv 32 -> 64)
    halo, b = rand_int(rng, (144, 32)), rand_int(rng, (64, 32))
    img = Image(9216 + 4096)
    img.put_rows_swizzled(0, halo, 64)
    img.put_rows_swizzled(9216, b, 64)
    ad = [desc(r * 512 + kk * 32, 0, 512, SW_64) for kk in range(2)]
    bd = [desc(9216 + kk * 32, 0, 512, SW_64) for kk in range(2)]
    got2 = run(img, ad, bd, 64)
    ok2 = np.array_equal(got2, halo[r * 8: r * 8 + 128] @ b.T)
    record(f"sw64_tap_shift_r{r}", ok2, True)
    assert ok and ok2


def test_noswizzle_canonical_and_field_meaning():
    """K = 32 (4 chunks of 8): which of LBO / SBO is the K-direction stride?  conv1's B operand assumes
    LBO = K-adjacent core matrices, SBO = next 8 rows."""
    rng = np.random.default_rng(3)
    a, b = rand_int(rng, (128, 32)), rand_int(rng, (64, 32))
    res = {}
    for name, (k_stride_field) in (("lbo_is_k", "lbo"), ("sbo_is_k", "sbo")):
        img = Image(8192 + 4096)
        # physical layout: chunks along K contiguous (128 B apart), 8-row groups 512 B apart
        img.put_core_matrices(0, a, 128, 512)
        img.put_core_matrices(8192, b, 128, 512)
        if k_stride_field == "lbo":
            mk = lambda addr: desc(addr, 128, 512, SW_NONE)          # noqa: E731
        else:
            mk = lambda addr: desc(addr, 512, 128, SW_NONE)          # noqa: E731
        ad = [mk(kk * 256) for kk in range(2)]
        bd = [mk(8192 + kk * 256) for kk in range(2)]
        got = run(img, ad, bd, 64)
        res[name] = bool(np.array_equal(got, a @ b.T))
    record("nosw_lbo_is_k", res["lbo_is_k"], True, res)
    assert res["lbo_is_k"], res


def test_noswizzle_overlapping_windows_conv1():
    """conv1: A rows are pixel PAIRS 16 B apart in a raw [38][192 B] image patch, K-adjacent core matrix
    = next two pixels (LBO 16 B), next 8 rows = next output-row pair = two image rows down (SBO 384 B).
    All 16 UMMAs of a tile, N = 128 (2x2 output pixels per GEMM row)."""
    rng = np.random.default_rng(4)
    patch = rand_int(rng, (38, 96))                # 24 px * 4 ch per row
    bmat = rand_int(rng, (128, 256), -2, 3)
    a_bytes = 7424
    img = Image(a_bytes + 65536)
    img.put_raw(0, patch)
    img.put_core_matrices(a_bytes, bmat, 128, 4096)
    ad, bd = [], []
    for r in range(8):
        for kk in range(2):
            ad.append(desc(r * 192 + kk * 32, 16, 384, SW_NONE))
            bd.append(desc(a_bytes + (r * 4 + kk * 2) * 128, 128, 4096, SW_NONE))
    got = run(img, ad, bd, 128)
    # expected: A[m = yp*8 + xp][r*32 + j] = patch[2*yp + r][xp*8 + j]
    a = np.zeros((128, 256), np.float32)
    for yp in range(16):
        for xp in range(8):
            for r in range(8):
                a[yp * 8 + xp, r * 32:(r + 1) * 32] = patch[2 * yp + r, xp * 8: xp * 8 + 32]
    ok = np.array_equal(got, a @ bmat.T)
    record("nosw_overlap_conv1", ok, True, {"maxerr": float(np.abs(got - a @ bmat.T).max())})
    assert ok


def test_exploratory_row_shifts_inside_swizzle_atoms():
    """Not used by shipped kernels: start addresses shifted by single 128 B rows (an x tap shift on one
    halo copy) and 8-row groups that are not 1024-byte multiples apart.  Recorded, not asserted."""
    rng = np.random.default_rng(5)
    halo, b = rand_int(rng, (200, 64)), rand_int(rng, (128, 64))
    img = Image(25600 + 16384)
    img.put_rows_swizzled(0, halo, 128)
    img.put_rows_swizzled(25600, b, 128)
    bd = [desc(25600 + kk * 32, 0, 1024, SW_128) for kk in range(4)]
    for s in (1, 2, 3):
        for bo in sorted({0, s}):
            ad = [desc(s * 128 + kk * 32, 0, 1024, SW_128, bo) for kk in range(4)]
            got = run(img, ad, bd, 128)
            ok = np.array_equal(got, halo[s: s + 128] @ b.T)
            record(f"explore_sw128_rowshift{s}_baseoff{bo}", ok, False)
    # groups 10 rows apart (a [18][10]-pixel halo tile): row m -> smem row (m//8)*10 + m%8 + shift
    for shift in (0, 1, 11):
        idx = np.array([(m // 8) * 10 + m % 8 + shift for m in range(128)])
        for bo in sorted({0, shift % 8}):
            ad = [desc(shift * 128 + kk * 32, 0, 1280, SW_128, bo) for kk in range(4)]
            got = run(img, ad, bd, 128)
            ok = np.array_equal(got, halo[idx] @ b.T)
            record(f"explore_sw128_sbo1280_shift{shift}_baseoff{bo}", ok, False)


def test_exploratory_issue_rate():
    """SM-clock cycles per UMMA (M=128, K=16) for the operand layouts in use -- recorded only."""
    from skin_image_analysis_b200 import debug_probes as probes
    rng = np.random.default_rng(6)
    out = {}
    for n in (32, 64, 128, 256):
        a, b = rand_int(rng, (128, 64)), rand_int(rng, (n, 64))
        img = Image(16384 + n * 128)
        img.put_rows_swizzled(0, a, 128)
        img.put_rows_swizzled(16384, b, 128)
        ad = [desc(kk * 32, 0, 1024, SW_128) for kk in range(4)] * 8
        bd = [desc(16384 + kk * 32, 0, 1024, SW_128) for kk in range(4)] * 8
        _, cyc = probes.umma_probe(img.tensor(), ad, bd, n, repeat=64, want_cycles=True)
        out[f"sw128_n{n}_cycles_per_mma"] = cyc / (64 * 32)
    # 64-byte rows (conv 32->64) and the single-halo addressing (8-row groups one halo row apart)
    for name, rb, lay, n, shift, sbo in (("sw64_n64_aligned", 64, SW_64, 64, 0, 8 * 64),
                                         ("sw64_n64_halo_shift11", 64, SW_64, 64, 11, 10 * 64),
                                         ("sw64_n64_halo_shift0", 64, SW_64, 64, 0, 10 * 64),
                                         ("sw128_n128_halo_shift11", 128, SW_128, 128, 11, 10 * 128),
                                         ("sw128_n64_halo_shift11", 128, SW_128, 64, 11, 10 * 128)):
        k = rb // 2
        halo, b = rand_int(rng, (200, k)), rand_int(rng, (n, k))
        a_bytes = (200 * rb + 1023) // 1024 * 1024
        img = Image(a_bytes + n * rb)
        img.put_rows_swizzled(0, halo, rb)
        img.put_rows_swizzled(a_bytes, b, rb)
        ad = [desc(shift * rb + kk * 32, 0, sbo, lay) for kk in range(k // 16)] * (32 // (k // 16))
        bd = [desc(a_bytes + kk * 32, 0, 8 * rb, lay) for kk in range(k // 16)] * (32 // (k // 16))
        _, cyc = probes.umma_probe(img.tensor(), ad, bd, n, repeat=64, want_cycles=True)
        out[name + "_cycles_per_mma"] = cyc / (64 * 32)
    patch, bmat = rand_int(rng, (38, 96)), rand_int(rng, (128, 256))
    img = Image(7424 + 65536)
    img.put_raw(0, patch)
    img.put_core_matrices(7424, bmat, 128, 4096)
    ad = [desc(r * 192 + kk * 32, 16, 384, SW_NONE) for r in range(8) for kk in range(2)]
    bd = [desc(7424 + (r * 4 + kk * 2) * 128, 128, 4096, SW_NONE) for r in range(8) for kk in range(2)]
    _, cyc = probes.umma_probe(img.tensor(), ad, bd, 128, repeat=128, want_cycles=True)
    out["conv1_nosw_n128_cycles_per_mma"] = cyc / (128 * 16)
    record("issue_rate", True, False, out)


# ----------------------------------------------------------------------------------------------------
# TMA bring-up: the exact shared-memory image of the boxes the conv / linear kernels request
# ----------------------------------------------------------------------------------------------------
def _coded(shape, seed):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(-120, 121, shape).astype(np.float32)).to(torch.bfloat16).cuda()


def _swizzle_rows(rows_u16, row_bytes, swizzle):
    """rows_u16 [R, row_bytes/2] -> bytes as TMA lays them out (16-byte units XORed by the row index)."""
    out = np.zeros_like(rows_u16)
    units = row_bytes // 16
    for r in range(rows_u16.shape[0]):
        x = 0 if swizzle == 0 else (r % 8) if swizzle == 128 else ((r >> 1) & 3) if swizzle == 64 else ((r >> 2) & 1)
        for u in range(units):
            out[r, (u ^ x) * 8:(u ^ x) * 8 + 8] = rows_u16[r, u * 8:u * 8 + 8]
    return out.reshape(-1)


def _bits(t):
    return t.cpu().view(torch.int16).numpy().view(np.uint16)


@pytest.mark.parametrize("cin,swz", [(64, 128), (32, 64)])
@pytest.mark.parametrize("coords", [(0, -1, -1, 0), (0, 8, 15, 1), (0, 3, 2, 1)])
def test_tma_conv3x3_halo_box(cin, swz, coords):
    """box [C<=64, 8 x, 18 y, 1 n] of an NHWC tensor: 144 smem rows of one pixel, zero outside the image."""
    from skin_image_analysis_b200 import debug_probes as probes
    b, h, w = 2, 20, 16
    t = _coded((b, h, w, cin), 7)
    got = probes.tma_probe(t, (cin, w, h, b), (cin * 2, w * cin * 2, h * w * cin * 2), (cin, 8, 18, 1), swz, coords)
    src = _bits(t)
    rows = np.zeros((144, cin), np.uint16)
    for yy in range(18):
        for xx in range(8):
            y, x = coords[2] + yy, coords[1] + xx
            if 0 <= y < h and 0 <= x < w:
                rows[yy * 8 + xx] = src[coords[3], y, x]
    want = _swizzle_rows(rows, cin * 2, swz)
    ok = np.array_equal(got.cpu().numpy().view(np.uint16), want)
    record(f"tma_conv3x3_c{cin}_{coords}", ok, True)
    assert ok


@pytest.mark.parametrize("coords", [(-8, -3, 0), (56, 13, 1), (120, 29, 1)])
def test_tma_conv1_patch_box(coords, rows=38):
    """box [24 px * 4 ch, 38 rows, 1] of the padded NHWC4 image seen as [B, H, (W+8)*4].  The innermost
    start must be a multiple of 8 elements (16 bytes) -- (x0-2)*4 with x0 % 16 == 0 -- an unaligned start
    raises an illegal-instruction fault (found on the first bring-up run)."""
    from skin_image_analysis_b200 import debug_probes as probes
    b, h, w = 2, 32, 56
    t = _coded((b, h, w * 4), 8)
    n_rows = rows
    got = probes.tma_probe(t, (w * 4, h, b), (w * 8, h * w * 8), (96, n_rows, 1), 0, coords)
    src = _bits(t)
    rows = np.zeros((n_rows, 96), np.uint16)
    for yy in range(n_rows):
        y = coords[1] + yy
        for e in range(96):
            x = coords[0] + e
            if 0 <= y < h and 0 <= x < w * 4:
                rows[yy, e] = src[coords[2], y, x]
    ok = np.array_equal(got.cpu().numpy().view(np.uint16), rows.reshape(-1))
    record(f"tma_conv1_{coords}", ok, True)
    assert ok


def test_tma_linear_tile_box():
    from skin_image_analysis_b200 import debug_probes as probes
    m, k = 200, 256
    t = _coded((m, k), 9)
    got = probes.tma_probe(t, (k, m), (k * 2,), (64, 128), 128, (64, 128))
    src = _bits(t)
    rows = np.zeros((128, 64), np.uint16)
    rows[:72] = src[128:200, 64:128]
    ok = np.array_equal(got.cpu().numpy().view(np.uint16), _swizzle_rows(rows, 128, 128))
    record("tma_linear_tile", ok, True)
    assert ok


def test_exploratory_alu_rates():
    """Lane-ops per SM clock of the instruction kinds the preprocess kernel can be built from."""
    import ctypes
    from skin_image_analysis_b200 import _lib
    out = (ctypes.c_double * 8)()
    _lib.check(_lib.load_debug().sia_debug_alu_rates(out, 8))
    names = ["FFMA", "PRMT", "I2F_U8_plus_IADD", "DP4A", "DP2A", "IMAD", "SHF", "FFMA2"]
    record("alu_rates_lane_ops_per_clk_per_sm", True, False, dict(zip(names, [round(v, 1) for v in out])))


@pytest.mark.parametrize("row_bytes,layout", [(128, SW_128), (64, SW_64)])
def test_single_halo_copy_taps(row_bytes, layout):
    """conv3x3 A operand: ONE [18][10]-pixel halo copy in the swizzled layout; tap (r,s) reads rows
    (y+r)*10 + (x+s): start shifted by (r*10+s) rows, 8-row groups one halo row (10 rows) apart.  Works
    because the swizzle XOR is taken from the absolute shared-memory address (base_offset stays 0)."""
    rng = np.random.default_rng(row_bytes)
    k = row_bytes // 2
    n = 128 if row_bytes == 128 else 64
    halo, b = rand_int(rng, (180, k)), rand_int(rng, (n, k))
    a_bytes = (180 * row_bytes + 1023) // 1024 * 1024
    img = Image(a_bytes + n * row_bytes)
    img.put_rows_swizzled(0, halo, row_bytes)
    img.put_rows_swizzled(a_bytes, b, row_bytes)
    all_ok = True
    for r in range(3):
        for s in range(3):
            shift = r * 10 + s
            ad = [desc(shift * row_bytes + kk * 32, 0, 10 * row_bytes, layout) for kk in range(k // 16)]
            bd = [desc(a_bytes + kk * 32, 0, 8 * row_bytes, layout) for kk in range(k // 16)]
            got = run(img, ad, bd, n)
            idx = np.array([(m // 8) * 10 + m % 8 + shift for m in range(128)])
            all_ok &= bool(np.array_equal(got, halo[idx] @ b.T))
    record(f"single_halo_taps_rowbytes{row_bytes}", all_ok, True)
    assert all_ok


def test_exploratory_tma_box_throughput():
    """SM-clock cycles per TMA box (32 boxes in flight from one SM) for the shapes in use and candidates."""
    from skin_image_analysis_b200 import debug_probes as probes
    out = {}
    rep = 32
    # conv1 patch boxes on a [B, H, WP*4] view of the padded NHWC4 image
    for name, wp, box_px, rows, c0 in (("conv1_192B_rows_pitch232_start-16B", 232, 24, 38, -8),
                                       ("conv1_192B_rows_pitch240_start_aligned128", 240, 24, 38, 0),
                                       ("conv1_256B_rows_pitch256_start_aligned128", 256, 32, 38, 0),
                                       ("conv1_128B_rows_pitch240_start_aligned128", 240, 16, 38, 0)):
        b, h = 4, 224
        t = _coded((b, h, wp * 4), 1)
        _, cyc = probes.tma_probe(t, (wp * 4, h, b), (wp * 8, h * wp * 8), (box_px * 4, rows, 1), 0, (c0, 0, 0),
                               repeat=rep, step_dim=1, step=4)
        out[name] = cyc / rep
    # conv3x3 halo boxes
    for name, cin, swz, box in (("conv3_sw128_10x18", 64, 128, (64, 10, 18, 1)), ("conv2_sw64_10x18", 32, 64, (32, 10, 18, 1)),
                                ("conv3_sw128_8x18", 64, 128, (64, 8, 18, 1)), ("conv3_sw128_16x12", 64, 128, (64, 16, 12, 1))):
        b, h, w = 2, 112, 112
        t = _coded((b, h, w, cin), 2)
        _, cyc = probes.tma_probe(t, (cin, w, h, b), (cin * 2, w * cin * 2, h * w * cin * 2), box, swz, (0, 7, 3, 0),
                               repeat=rep, step_dim=2, step=2)
        out[name] = cyc / rep
    # plain 2-D GEMM tile
    t = _coded((1024, 4096), 3)
    _, cyc = probes.tma_probe(t, (4096, 1024), (8192,), (64, 128), 128, (0, 0), repeat=rep, step_dim=0, step=64)
    out["linear_sw128_64x128"] = cyc / rep
    record("tma_cycles_per_box", True, False, {k: round(v, 1) for k, v in out.items()})


def test_exploratory_a_operand_in_tensor_memory():
    """tcgen05.mma with A read from TMEM (what a second product on an accumulator needs).
    Hypothesis: for 16-bit kinds, lane m / 32-bit column c of the A region holds A[m, 2c] (low half) and A[m, 2c+1]
    (high half), and one K = 16 instruction consumes 8 columns.  B = [I | 0] picks A's columns back out, so the
    accumulator reveals the K index the hardware assigns to every stored half-word; a random product then checks
    the arithmetic.  Exploratory: recorded in gpurun_out/probe_report.json."""
    from skin_image_analysis_b200 import debug_probes as probes
    k, n = 64, 64
    a = (np.arange(128)[:, None] * 0 + np.arange(k)[None, :]).astype(np.float32)          # A[m, kk] = kk
    a[:, 0] = np.arange(128) % 64                                                         # column 0 carries the row
    bits = bf16_bits(a).astype(np.uint32)
    words = (bits[:, 0::2] | (bits[:, 1::2] << 16)).astype(np.uint32).view(np.int32)      # [128, k/2]
    img = Image(n * k * 2)
    img.put_core_matrices(0, np.eye(n, k, dtype=np.float32), 128, (k // 8) * 128)         # B[n, kk] = delta
    bd = [desc(kk * 256, 128, (k // 8) * 128, SW_NONE) for kk in range(k // 16)]
    got = probes.umma_ts_probe(img.tensor(), torch.from_numpy(words.copy()).cuda(), 8, bd, n).cpu().numpy()
    ok_map = np.array_equal(got, a[:, :n])
    rng = np.random.default_rng(3)
    a2, b2 = rand_int(rng, (128, k)), rand_int(rng, (n, k))
    bits = bf16_bits(a2).astype(np.uint32)
    words = (bits[:, 0::2] | (bits[:, 1::2] << 16)).astype(np.uint32).view(np.int32)
    img2 = Image(n * k * 2)
    img2.put_core_matrices(0, b2, 128, (k // 8) * 128)
    got2 = probes.umma_ts_probe(img2.tensor(), torch.from_numpy(words.copy()).cuda(), 8, bd, n).cpu().numpy()
    ok_rand = np.array_equal(got2, a2 @ b2.T)
    record("explore_a_in_tmem_f16", ok_map and ok_rand, False,
           {"mapping_ok": bool(ok_map), "random_ok": bool(ok_rand), "row0": got[0, :16].tolist(), "row5": got[5, :16].tolist(),
            "col0": got[:8, 0].tolist()})
    assert got.shape == (128, n)
