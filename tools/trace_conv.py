#!/usr/bin/env python
"""Event timeline of CTA 0 of the conv kernels at the bench shape (run under gpurun).

Prints, per tile, SM-clock cycles relative to the first event: when the producer saw a free stage and
issued the TMA load, when the MMA warp got its accumulator / operands and finished issuing, when the
epilogue saw the finished accumulator, drained it and stored the outputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from skin_image_analysis_b200 import _lib, ops  # noqa: E402

NAMES = ["P:stage_free", "P:tma_issued", "M:acc_free", "M:ops_landed", "M:issued", "E:acc_done", "E:drained", "E:stored"]


def run(name, fn, tiles=40):
    lib = _lib.load()
    buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    fn()
    torch.cuda.synchronize()
    _lib.check(lib.sia_debug_set_trace(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    _lib.check(lib.sia_debug_set_trace(0))
    t = buf.view(64, 8).cpu().numpy().astype(np.int64)
    t0 = t[t > 0].min()
    rel = np.where(t > 0, t - t0, -1)
    lines = [f"== {name}  (cycles since first event)", "tile " + " ".join(f"{n:>13s}" for n in NAMES)]
    for i in range(tiles):
        lines.append(f"{i:4d} " + " ".join(f"{v:13d}" for v in rel[i]))
    d = np.diff(rel[8:tiles, 4])
    lines.append(f"steady-state tile period (MMA issue-to-issue): mean {d.mean():.0f}  min {d.min()}  max {d.max()}")
    lat = rel[8:tiles, 3] - rel[8:tiles, 1]
    lines.append(f"TMA issue -> operands seen by MMA warp: mean {lat.mean():.0f} (includes queueing behind earlier tiles)")
    out = "\n".join(lines)
    print(out)
    return out


def main():
    b = 256
    g = torch.Generator(device="cuda").manual_seed(0)
    x4 = torch.zeros(b, 224, 232, 4, dtype=torch.bfloat16, device="cuda")
    x4[:, :, 1:225, :3] = torch.rand(b, 224, 224, 3, device="cuda", generator=g).to(torch.bfloat16)
    w1 = ops.pack_conv7x7_c3(torch.randn(32, 3, 7, 7, device="cuda", generator=g) * 0.1)
    w2 = ops.pack_conv3x3(torch.randn(64, 32, 3, 3, device="cuda", generator=g) * 0.05)
    w3 = ops.pack_conv3x3(torch.randn(128, 64, 3, 3, device="cuda", generator=g) * 0.05)
    b1, b2, b3 = (torch.zeros(n, device="cuda") for n in (32, 64, 128))
    a1 = ops.conv7x7_c3_relu_pool2(x4, w1, b1)
    a2 = ops.conv3x3_relu_pool2(a1, w2, b2, 64)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/trace_conv.txt", "w") as f:
        f.write(run("conv1", lambda: ops.conv7x7_c3_relu_pool2(x4, w1, b1, out=a1)) + "\n")
        f.write(run("conv2", lambda: ops.conv3x3_relu_pool2(a1, w2, b2, 64, out=a2)) + "\n")
        f.write(run("conv3", lambda: ops.conv3x3_relu_pool2(a2, w3, b3, 128)) + "\n")


if __name__ == "__main__":
    main()
