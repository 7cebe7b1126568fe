#!/usr/bin/env python
"""Small, fixed workload for ncu: two batches of the BASELINE configs[1] step, kernels launched
directly (no CUDA graph) so every launch is visible by name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import OUT, SRC_H, SRC_W, build_state  # noqa: E402
from skin_image_analysis_b200.engine import EvalEngine  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda", 0)
    state = build_state(dev, centre=False)          # no calibration forward: every launch below is the step
    eng = EvalEngine(state, batch, (SRC_H, SRC_W), OUT, device=dev, use_graph=False, n_slots=1)
    g = torch.Generator(device=dev).manual_seed(0)
    eng.u8[0].copy_(torch.randint(0, 256, eng.u8[0].shape, dtype=torch.uint8, device=dev, generator=g))
    eng.groups[0].zero_()
    for _ in range(steps):
        eng.step_resident(0)
    eng.synchronize()
    # the ToneClassifier transform variant on the same decode buffers (one launch)
    from skin_image_analysis_b200 import ops
    ops.preprocess_tv_u8hwc(eng.u8[0], (OUT, OUT), ops.LAYOUT_NHWC4_BF16, out=eng.x4)
    # the round-1 default, the two-product tcgen05 preprocess, one launch for comparison
    ops.preprocess_u8hwc(eng.u8[0], (OUT, OUT), ops.LAYOUT_NHWC4_BF16, out=eng.x4, impl="tensor_core2")
    torch.cuda.synchronize()
    print("counted", int(eng.read_counts()[0].sum()))


if __name__ == "__main__":
    main()
