#!/usr/bin/env python
"""Event timeline of CTA 0 of the tensor-core preprocess kernel (needs a -DSIA_INSTRUMENT build; run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skin_image_analysis_b200 import _lib, ops
NAMES = ["P:raw_free", "P:tma_issued", "C:raw_landed", "C:converted", "M:acc_free", "M:issued", "E:acc_seen", "E:released"]
b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
u8 = torch.randint(0, 256, (b, 450, 600, 3), dtype=torch.uint8, device="cuda", generator=g)
x4 = torch.empty((b, 224, 232, 4), dtype=torch.bfloat16, device="cuda")
fn = lambda: ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16, out=x4, impl="tensor_core")
buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
fn(); torch.cuda.synchronize()
_lib.check(lib.sia_debug_set_trace(buf.data_ptr()))
fn(); torch.cuda.synchronize()
_lib.check(lib.sia_debug_set_trace(0))
t = buf.view(64, 8).cpu().numpy().astype(np.int64)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
print("seq " + " ".join(f"{n:>13s}" for n in NAMES))
for i in range(40):
    print(f"{i:4d} " + " ".join(f"{v:13d}" for v in rel[i]))
