#!/usr/bin/env python
"""CUDA-event time of the two preprocess kernels at the bench shape (run under gpurun)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skin_image_analysis_b200 import ops

def timed(fn, iters=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
ring = [torch.randint(0, 256, (b, 450, 600, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(4)]
x4 = torch.empty((b, 224, 232, 4), dtype=torch.bfloat16, device="cuda")
res = {"batch": b}
for impl in ("cuda_core", "tensor_core", "tensor_core2"):
    k = [0]
    def fn():
        k[0] += 1
        ops.preprocess_u8hwc(ring[k[0] % 4], (224, 224), ops.LAYOUT_NHWC4_BF16, out=x4, impl=impl)
    ms = timed(fn)
    res[impl] = {"ms": round(ms, 4), "GBps": round(b * 1111056 / ms / 1e6, 1)}
print(json.dumps(res))
