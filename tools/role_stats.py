#!/usr/bin/env python
"""Per-role wait breakdown of the conv kernels at the bench shape (needs a GPU; run under gpurun).

For each kernel: average SM-clock cycles per tile that the TMA producer waits for a free stage, the MMA
warp waits for a free accumulator / for operands, and the epilogue waits for a finished accumulator --
i.e. which of TMA, tensor pipe and epilogue is the bottleneck."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from skin_image_analysis_b200 import _lib, ops  # noqa: E402


def run(name, fn, grid=148):
    buf = torch.zeros(grid * 8, dtype=torch.int64, device="cuda")
    lib = _lib.load()
    fn()
    torch.cuda.synchronize()
    _lib.check(lib.sia_debug_set_stats(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    _lib.check(lib.sia_debug_set_stats(0))
    s = buf.view(grid, 8).double().cpu()
    tiles = s[:, 6].clamp(min=1)
    keys = ["producer_wait_stage", "mma_wait_accumulator", "mma_wait_operands", "mma_loop", "epi_wait_accumulator",
            "epi_loop"]
    out = {k: float((s[:, i] / tiles).mean()) for i, k in enumerate(keys)}
    out["tiles_per_cta"] = float(tiles.mean())
    print(name, json.dumps({k: round(v, 1) for k, v in out.items()}))
    return out


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    g = torch.Generator(device="cuda").manual_seed(0)
    x4 = torch.zeros(b, 224, 232, 4, dtype=torch.bfloat16, device="cuda")
    x4[:, :, 1:225, :3] = torch.rand(b, 224, 224, 3, device="cuda", generator=g).to(torch.bfloat16)
    w1 = ops.pack_conv7x7_c3(torch.randn(32, 3, 7, 7, device="cuda", generator=g) * 0.1)
    w2 = ops.pack_conv3x3(torch.randn(64, 32, 3, 3, device="cuda", generator=g) * 0.05)
    w3 = ops.pack_conv3x3(torch.randn(128, 64, 3, 3, device="cuda", generator=g) * 0.05)
    b1, b2, b3 = (torch.zeros(n, device="cuda") for n in (32, 64, 128))
    a1 = ops.conv7x7_c3_relu_pool2(x4, w1, b1)
    a2 = ops.conv3x3_relu_pool2(a1, w2, b2, 64)
    res = {
        "conv1": run("conv1", lambda: ops.conv7x7_c3_relu_pool2(x4, w1, b1, out=a1)),
        "conv2": run("conv2", lambda: ops.conv3x3_relu_pool2(a1, w2, b2, 64, out=a2)),
        "conv3": run("conv3", lambda: ops.conv3x3_relu_pool2(a2, w3, b3, 128)),
    }
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/role_stats.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
