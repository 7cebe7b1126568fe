#!/usr/bin/env python
"""Times the preprocess kernels alone (CUDA events, inputs rotating over 4 distinct batches > L2):
    python tools/time_pre.py [batch] [impl ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from skin_image_analysis_b200 import ops  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
impls = sys.argv[2:] or ["mma", "tensor_core2", "tensor_core", "cuda_core"]
size = int(os.environ.get("OUT", "224"))
g = torch.Generator(device="cuda").manual_seed(0)
bufs = [torch.randint(0, 256, (batch, 450, 600, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(4)]
out = torch.empty((batch, size, size + 8, 4), dtype=torch.bfloat16, device="cuda")
bytes_alg = batch * (450 * 600 * 3 + 3 * size * size * 2)
for impl in impls:
    try:
        for i in range(4):
            ops.preprocess_u8hwc(bufs[i], (size, size), ops.LAYOUT_NHWC4_BF16, out=out, impl=impl)
    except Exception as exc:
        print(f"{impl:14s} unsupported: {exc}")
        continue
    torch.cuda.synchronize()
    n = 40
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        ops.preprocess_u8hwc(bufs[i % 4], (size, size), ops.LAYOUT_NHWC4_BF16, out=out, impl=impl)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    med = ts[n // 2]
    print(f"{impl:14s} batch {batch} out {size}: median {med:.4f} ms  min {ts[0]:.4f}  "
          f"{bytes_alg / med / 1e6:.0f} GB/s algorithmic = {bytes_alg / med / 1e6 / 6541.8:.3f} of the HBM copy peak")
