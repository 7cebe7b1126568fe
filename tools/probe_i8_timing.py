"""Cycles per tcgen05.mma for operand layout variants (timing only; operands are zeros)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skin_image_analysis_b200 import debug_probes as probes
from tests.test_umma_probe import desc
from tests.test_umma_i8_probe import idesc_i8

SW_NONE, SW_128 = 0, 2
img = torch.zeros(190 * 1024, dtype=torch.uint8, device="cuda")

def cyc(kind, ad, bd, n, idesc):
    f = probes.umma_probe_i8 if kind == 1 else None
    if kind == 1:
        _, c1 = probes.umma_probe_i8(img, ad, bd, n, idesc, repeat=4, want_cycles=True)
        _, c2 = probes.umma_probe_i8(img, ad, bd, n, idesc, repeat=36, want_cycles=True)
    else:
        _, c1 = probes.umma_probe(img, ad, bd, n, repeat=4, want_cycles=True)
        _, c2 = probes.umma_probe(img, ad, bd, n, repeat=36, want_cycles=True)
    return (c2 - c1) / (32 * len(ad))

B0 = 64 * 1024
for n in (64, 96, 128, 256):
    r = {}
    # bf16 SW128 K-major both, 4 k-steps of 16 elements inside one 128-byte row
    ad = [desc(kk * 32, 0, 1024, SW_128) for kk in range(4)]
    bd = [desc(B0 + kk * 32, 0, 1024, SW_128) for kk in range(4)]
    r["bf16_sw128"] = cyc(0, ad, bd, n, 0)
    # bf16 no-swizzle K-major both (K = 64: 8 cores of 8 elements)
    ad = [desc(kk * 256, 128, 1024, SW_NONE) for kk in range(4)]
    bd = [desc(B0 + kk * 256, 128, 1024, SW_NONE) for kk in range(4)]
    r["bf16_none"] = cyc(0, ad, bd, n, 0)
    # i8 no-swizzle K-major both, K = 256
    k = 256
    ad = [desc(s * 256, 128, (k // 16) * 128, SW_NONE) for s in range(k // 32)]
    bd = [desc(B0 + s * 256, 128, (k // 16) * 128, SW_NONE) for s in range(k // 32)]
    r["i8_none_kk"] = cyc(1, ad, bd, n, idesc_i8(128, n))
    # i8 SW128 K-major both (rows of 128 bytes = 4 k-steps)
    ad = [desc(s * 32, 0, 1024, SW_128) for s in range(4)]
    bd = [desc(B0 + s * 32, 0, 1024, SW_128) for s in range(4)]
    r["i8_sw128_kk"] = cyc(1, ad, bd, n, idesc_i8(128, n))
    # i8 A no-swizzle K-major, B MN-major no-swizzle (padded SBO)
    ad = [desc(s * 256, 128, (k // 16) * 128, SW_NONE) for s in range(k // 32)]
    bd = [desc(B0 + s * 512, 128, (k // 8) * 128 + 16, SW_NONE) for s in range(k // 32)]
    r["i8_none_b_mn"] = cyc(1, ad, bd, n, idesc_i8(128, n, b_mn=1))
    # same with unpadded SBO
    bd = [desc(B0 + s * 512, 128, (k // 8) * 128, SW_NONE) for s in range(k // 32)]
    r["i8_none_b_mn_nopad"] = cyc(1, ad, bd, n, idesc_i8(128, n, b_mn=1))
    # i8 A SW128 K-major, B MN-major SW128: 128 bytes of N per k-row, k-groups 1024 apart, N-blocks at LBO
    kb = 128
    ad = [desc(s * 32, 0, 1024, SW_128) for s in range(4)]
    bd = [desc(B0 + s * 4096, (kb // 8) * 1024, 1024, SW_128) for s in range(4)]
    r["i8_sw128_b_mn"] = cyc(1, ad, bd, n, idesc_i8(128, n, b_mn=1))
    # A no-swizzle, B MN-major SW128
    ad = [desc(s * 256, 128, (kb // 16) * 128, SW_NONE) for s in range(4)]
    r["i8_anone_b_mn_sw128"] = cyc(1, ad, bd, n, idesc_i8(128, n, b_mn=1))
    print(n, {a: round(b, 1) for a, b in r.items()}, flush=True)
