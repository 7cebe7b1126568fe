"""Cycles per bf16 UMMA (N=128, K=16, no-swizzle operands) when the accumulator changes every k MMAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skin_image_analysis_b200 import _lib, debug_probes as probes
from tests.test_umma_probe import desc
lib = _lib.load_debug()
img = torch.zeros(160 * 1024, dtype=torch.uint8, device="cuda")
B0 = 64 * 1024
ad = [desc(kk * 256, 128, 1024, 0) for kk in range(16)]
bd = [desc(B0 + kk * 256, 128, 1024, 0) for kk in range(16)]
def cyc(n):
    _, c1 = probes.umma_probe(img, ad, bd, n, repeat=8, want_cycles=True)
    _, c2 = probes.umma_probe(img, ad, bd, n, repeat=72, want_cycles=True)
    return (c2 - c1) / (64 * 16)
for n in (128,):
    for k, commit in ((0, 0), (16, 0), (16, 1), (8, 0), (8, 1), (4, 0), (2, 0)):
        lib.sia_debug_umma_probe_switch(k, commit)
        print(f"n={n} switch_every={k} commit={commit}: {cyc(n):.1f} clk/MMA", flush=True)
lib.sia_debug_umma_probe_switch(0, 0)
