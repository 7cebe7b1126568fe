#!/usr/bin/env python
"""CUDA-event time of every kernel of the step at the bench shape (run under gpurun).

    [SIA_LIB_PATH=other.so] python tools/stage_times.py [batch] [tag]

Prints one JSON line {stage: ms}.  Symbols missing from an older library are tolerated so that two
builds can be compared (A/B) with the same script."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from skin_image_analysis_b200 import _lib  # noqa: E402

if os.environ.get("SIA_LIB_PATH"):          # older builds lack the newest debug symbols
    import ctypes
    probe = ctypes.CDLL(os.environ["SIA_LIB_PATH"])
    for name in list(_lib.SIGNATURES):
        if not hasattr(probe, name):
            del _lib.SIGNATURES[name]

from skin_image_analysis_b200 import ops  # noqa: E402


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


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tag = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("SIA_LIB_PATH", "head")
    g = torch.Generator(device="cuda").manual_seed(0)
    u8 = torch.randint(0, 256, (b, 450, 600, 3), dtype=torch.uint8, device="cuda", generator=g)
    x4 = ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16)
    w1 = ops.pack_conv7x7_c3(torch.randn(32, 3, 7, 7, device="cuda", generator=g) * 0.1)
    w2 = ops.pack_conv3x3(torch.randn(64, 32, 3, 3, device="cuda", generator=g) * 0.05)
    w3 = ops.pack_conv3x3(torch.randn(128, 64, 3, 3, device="cuda", generator=g) * 0.05)
    b1, b2, b3 = (torch.zeros(n, device="cuda") for n in (32, 64, 128))
    a1 = ops.conv7x7_c3_relu_pool2(x4, w1, b1)
    a2 = ops.conv3x3_relu_pool2(a1, w2, b2, 64)
    a3 = ops.conv3x3_relu_pool2(a2, w3, b3, 128)
    res = {"tag": tag, "batch": b}
    res["preprocess"] = timed(lambda: ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16, out=x4))
    res["conv1"] = timed(lambda: ops.conv7x7_c3_relu_pool2(x4, w1, b1, out=a1))
    res["conv2"] = timed(lambda: ops.conv3x3_relu_pool2(a1, w2, b2, 64, out=a2))
    res["conv3"] = timed(lambda: ops.conv3x3_relu_pool2(a2, w3, b3, 128, out=a3))
    w_fc = torch.randn(512, 100352, device="cuda", generator=g).to(torch.bfloat16)
    splits = 18 if b >= 256 else 37
    part = ops.linear_splitk(a3.view(b, -1), w_fc, splits)
    b1f, b2f, b3f = torch.zeros(512, device="cuda"), torch.zeros(256, device="cuda"), torch.zeros(2, device="cuda")
    w2t = torch.randn(512, 256, device="cuda", generator=g) * 0.05
    w3f = torch.randn(2, 256, device="cuda", generator=g) * 0.05
    res["fc1"] = timed(lambda: ops.linear_splitk(a3.view(b, -1), w_fc, splits, out=part))
    res["tail"] = timed(lambda: ops.head_tail(part, b1f, w2t, b2f, w3f, b3f))
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))


if __name__ == "__main__":
    main()
