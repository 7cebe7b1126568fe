"""``torch.ops.sia_b200.*`` (TORCH_LIBRARY shim over the C ABI, csrc_torch/torch_ops.cpp): registered schemas, loud
refusal of CPU tensors / wrong dtypes (TORCH_CHECK), and -- on the GPU -- bit-identical results to the ctypes path."""
import numpy as np
import pytest
import torch

from tests import helpers


@pytest.fixture(scope="module")
def sia():
    from skin_image_analysis_b200 import torch_ops
    return torch_ops.load()


def test_ops_are_registered_and_refuse_cpu_tensors(sia):
    assert sia.version() == 100
    names = ["nchw_f32_to_nhwc4", "conv7x7_c3_relu_pool2", "conv3x3_relu_pool2", "linear_splitk", "head_tail",
             "confusion_counts", "preprocess_mma"]
    for n in names:
        schema = str(getattr(sia, n).default._schema)
        assert schema.startswith(f"sia_b200::{n}("), schema
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        sia.nchw_f32_to_nhwc4(torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        sia.confusion_counts(torch.zeros(4, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8),
                             torch.zeros((1, 4), dtype=torch.uint8), 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        sia.linear_splitk(torch.zeros(2, 64, dtype=torch.bfloat16), torch.zeros(2, 64, dtype=torch.bfloat16), 1)


@pytest.mark.gpu
def test_torch_ops_equal_the_ctypes_path(sia):
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200 import resize_weights as rw
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(2, 3, 64, 64, device="cuda", generator=g)
    x4 = sia.nchw_f32_to_nhwc4(x)
    assert torch.equal(x4, ops.nchw_f32_to_nhwc4(x))
    w1 = torch.randn(32, 3, 7, 7, device="cuda", generator=g) * 0.1
    b1 = torch.randn(32, device="cuda", generator=g) * 0.1
    p1 = ops.pack_conv7x7_c3(w1)
    a1 = sia.conv7x7_c3_relu_pool2(x4, p1, b1)
    assert torch.equal(a1, ops.conv7x7_c3_relu_pool2(x4, p1, b1))
    w2 = torch.randn(64, 32, 3, 3, device="cuda", generator=g) * 0.05
    b2 = torch.randn(64, device="cuda", generator=g) * 0.1
    p2 = ops.pack_conv3x3(w2)
    a2 = sia.conv3x3_relu_pool2(a1, p2, b2, 64)
    assert torch.equal(a2, ops.conv3x3_relu_pool2(a1, p2, b2, 64)) and a2.shape == (2, 16, 16, 64)
    a = torch.randn(8, 1280, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(512, 1280, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    part = sia.linear_splitk(a, w, 5)
    assert torch.equal(part, ops.linear_splitk(a, w, 5))
    hb1 = torch.randn(512, device="cuda", generator=g)
    hw2 = (torch.randn(256, 512, device="cuda", generator=g) * 0.05).t().contiguous()
    hb2 = torch.randn(256, device="cuda", generator=g)
    hw3 = torch.randn(2, 256, device="cuda", generator=g) * 0.1
    hb3 = torch.randn(2, device="cuda", generator=g)
    logp, pred = sia.head_tail(part, hb1, hw2, hb2, hw3, hb3)
    logp2, pred2 = ops.head_tail(part, hb1, hw2, hb2, hw3, hb3)
    assert torch.equal(logp, logp2) and torch.equal(pred, pred2)
    label = torch.randint(0, 2, (8,), device="cuda", dtype=torch.uint8)
    groups = torch.randint(0, 7, (3, 8), device="cuda", dtype=torch.uint8)
    assert torch.equal(sia.confusion_counts(pred, label, groups, 6), ops.confusion_counts(pred, label, groups, 6))
    # the fused transform
    u8 = torch.from_numpy(np.stack([helpers.synthetic_u8_image(450, 600, 60 + i, "smooth") for i in range(2)])).cuda()
    t = rw.build_mma_tables(450, 600, 224, 224)
    qs, cs = rw.MMA_ROW_MAPS[t.row_map]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()      # noqa: E731
    out = sia.preprocess_mma(u8, dev(t.wy_frag), dev(t.r0), dev(t.wx_frag), dev(t.wx_mask), dev(t.tile_begin), t.kv, qs,
                             list(cs), [rw.MMA_OUT_SCALE / 255.0] * 3, [0.0, 0.0, 0.0], 224, 224)
    assert torch.equal(out, ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="mma"))
    with pytest.raises(RuntimeError, match="expected"):
        sia.conv3x3_relu_pool2(a1.float(), p2, b2, 64)            # wrong dtype: TORCH_CHECK
