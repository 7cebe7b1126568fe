"""Hardware bring-up of tcgen05.mma kind::i8 as the tensor-core preprocess kernel uses it
(csrc/preprocess_tc.cu): A = u8 weight digits, K-major, no swizzle; B = the raw u8 image bytes, MN-major
(the byte index along the image row is the contiguous one), no swizzle, padded group stride; D = s32.

Integer arithmetic, so every comparison is ``==``.
"""
import numpy as np
import pytest
import torch

from tests.test_umma_probe import desc, record

pytestmark = pytest.mark.gpu

SW_NONE = 0


def idesc_i8(m, n, a_signed=0, b_signed=0, a_mn=0, b_mn=0):
    return (2 << 4) | (a_signed << 7) | (b_signed << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def put_k_major(buf, base, mat, lbo, sbo):
    """u8 [rows, K]: (r,k) at (r//8)*sbo + (k//16)*lbo + (r%8)*16 + k%16."""
    rows, kk = mat.shape
    for r in range(rows):
        for kc in range(kk // 16):
            dst = base + (r // 8) * sbo + kc * lbo + (r % 8) * 16
            buf[dst:dst + 16] = mat[r, kc * 16:(kc + 1) * 16]


def put_mn_major(buf, base, mat_kn, lbo, sbo):
    """u8 [K, N]: (k,n) at (n//16)*sbo + (k//8)*lbo + (k%8)*16 + n%16  (16 contiguous bytes of N per k)."""
    kk, n = mat_kn.shape
    for k in range(kk):
        for nu in range(n // 16):
            dst = base + nu * sbo + (k // 8) * lbo + (k % 8) * 16
            buf[dst:dst + 16] = mat_kn[k, nu * 16:(nu + 1) * 16]


@pytest.mark.parametrize("n", [96, 128, 256])
def test_i8_k_major_both(n):
    from skin_image_analysis_b200 import debug_probes as probes
    rng = np.random.default_rng(n)
    k = 64
    a = rng.integers(0, 256, (128, k), dtype=np.uint8)
    b = rng.integers(0, 256, (n, k), dtype=np.uint8)
    a_bytes, b_base = 128 * k, 128 * k
    buf = np.zeros(b_base + n * k, np.uint8)
    put_k_major(buf, 0, a, 128, (k // 16) * 128)
    put_k_major(buf, b_base, b, 128, (k // 16) * 128)
    ad = [desc(s * 256, 128, (k // 16) * 128, SW_NONE) for s in range(k // 32)]
    bd = [desc(b_base + s * 256, 128, (k // 16) * 128, SW_NONE) for s in range(k // 32)]
    got = probes.umma_probe_i8(torch.from_numpy(buf).cuda(), ad, bd, n, idesc_i8(128, n)).cpu().numpy()
    want = a.astype(np.int64) @ b.astype(np.int64).T
    ok = np.array_equal(got, want)
    record(f"i8_k_major_n{n}", ok, True, {"maxerr": float(np.abs(got - want).max())})
    assert ok, a_bytes


@pytest.mark.parametrize("n,k,pad", [(96, 64, 16), (96, 256, 16), (128, 64, 0), (256, 96, 16)])
def test_i8_b_mn_major(n, k, pad):
    """B = [K source rows, N bytes of the image row] exactly as the preprocess loader lays it out."""
    from skin_image_analysis_b200 import debug_probes as probes
    rng = np.random.default_rng(n + k)
    a = rng.integers(0, 256, (128, k), dtype=np.uint8)
    s = rng.integers(0, 256, (k, n), dtype=np.uint8)          # s[r, byte]
    a_sbo = (k // 16) * 128
    b_base = 128 * k
    b_lbo, b_sbo = 128, (k // 8) * 128 + pad
    buf = np.zeros(b_base + (n // 16) * b_sbo + 1024, np.uint8)
    put_k_major(buf, 0, a, 128, a_sbo)
    put_mn_major(buf, b_base, s, b_lbo, b_sbo)
    ad = [desc(st * 256, 128, a_sbo, SW_NONE) for st in range(k // 32)]
    bd = [desc(b_base + st * 4 * b_lbo, b_lbo, b_sbo, SW_NONE) for st in range(k // 32)]
    got = probes.umma_probe_i8(torch.from_numpy(buf).cuda(), ad, bd, n, idesc_i8(128, n, b_mn=1)).cpu().numpy()
    want = a.astype(np.int64) @ s.astype(np.int64)
    ok = np.array_equal(got, want)
    if not ok:      # the other reading of LBO / SBO for MN-major operands
        bd2 = [desc(b_base + st * 4 * b_lbo, b_sbo, b_lbo, SW_NONE) for st in range(k // 32)]
        got2 = probes.umma_probe_i8(torch.from_numpy(buf).cuda(), ad, bd2, n, idesc_i8(128, n, b_mn=1)).cpu().numpy()
        record(f"i8_b_mn_major_n{n}_k{k}_swapped", np.array_equal(got2, want), False)
    record(f"i8_b_mn_major_n{n}_k{k}", ok, True, {"maxerr": float(np.abs(got - want).max())})
    assert ok


def test_i8_issue_rate():
    """SM-clock cycles per kind::i8 MMA (M=128, K=32) for the N values the kernel may use."""
    from skin_image_analysis_b200 import debug_probes as probes
    out = {}
    for n in (96, 128, 192, 256):
        k = 256
        a_sbo, b_base = (k // 16) * 128, 128 * k
        b_lbo, b_sbo = 128, (k // 8) * 128 + 16
        buf = np.zeros(b_base + (n // 16) * b_sbo + 1024, np.uint8)
        ad = [desc(st * 256, 128, a_sbo, SW_NONE) for st in range(k // 32)]
        bd = [desc(b_base + st * 4 * b_lbo, b_lbo, b_sbo, SW_NONE) for st in range(k // 32)]
        img = torch.from_numpy(buf).cuda()
        _, c1 = probes.umma_probe_i8(img, ad, bd, n, idesc_i8(128, n, b_mn=1), repeat=4, want_cycles=True)
        _, c2 = probes.umma_probe_i8(img, ad, bd, n, idesc_i8(128, n, b_mn=1), repeat=36, want_cycles=True)
        out[f"n{n}"] = (c2 - c1) / (32 * len(ad))
    record("i8_cycles_per_mma_b_mn_major", True, False, out)
    print(out)
