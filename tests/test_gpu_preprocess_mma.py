"""K1-K3 parity of the warp-MMA preprocess kernel (csrc/preprocess_mma.cu, impl="mma" -- what impl="auto" picks for
the NHWC4 layout): both banded products of the resize on mma.sync with register-built operands."""
import numpy as np
import pytest
import torch

from oracle import resize as R
from tests import helpers

pytestmark = pytest.mark.gpu


def _gpu(u8_list, size, **kw):
    from skin_image_analysis_b200 import ops
    x = torch.from_numpy(np.stack(u8_list)).cuda()
    return ops.preprocess_u8hwc(x, size, ops.LAYOUT_NHWC4_BF16, **kw)


def _bf(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("src_hw,size,batch", [((450, 600), (224, 224), 5), ((480, 640), (224, 224), 3),
                                               ((450, 600), (512, 512), 2), ((300, 400), (224, 224), 3),
                                               ((96, 128), (64, 88), 4), ((600, 448), (296, 224), 2),
                                               ((450, 600), (200, 224), 2), ((200, 200), (224, 224), 2)])
def test_mma_pass_matches_oracle(src_hw, size, batch):
    """Within one bf16 ulp of the bf16-rounded oracle, <= 8e-4 of full scale before rounding (V is rounded once to
    fp16), equal to the numpy model of its arithmetic up to fp32 summation order, zero pads, batch independent.
    Geometries: the bench shape; 640-pixel rows (the "natural" row map); 512 x 512 (up-sampling in y, kv = 2, four
    tiles per group); 300 x 400 (kv = 2); a small one; portrait; out_h not a multiple of 16; up-sampling both ways."""
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200 import resize_weights as rw
    kinds = ["noise", "smooth", "extremes", "noise", "smooth"]
    imgs = [helpers.synthetic_u8_image(src_hw[0], src_hw[1], 300 + i, kinds[i]) for i in range(batch)]
    oh, ow = size
    got = _gpu(imgs, size, impl="mma").float().cpu().numpy()
    assert got.shape == (batch, oh, ow + ops.NHWC4_PAD, 4)
    assert np.all(got[..., 3] == 0) and np.all(got[:, :, 0] == 0) and np.all(got[:, :, ow + 1:] == 0)
    want = np.stack([R.transform_u8(im, size) for im in imgs]).transpose(0, 2, 3, 1)
    want_bf = _bf(want)
    px = got[:, :, 1:ow + 1, :3]
    ulp = np.maximum(np.abs(want_bf), 2.0 ** -126) * 2.0 ** -7
    assert np.all(np.abs(px - want_bf) <= ulp)
    assert np.abs(px - want).max() <= 2.0 ** -8 + 8e-4
    assert (px != want_bf).mean() < 0.06                      # V is rounded to fp16: 2-4 % of the outputs round the other way
    t = rw.build_mma_tables(src_hw[0], src_hw[1], oh, ow)
    model_bf = _bf(rw.mma_emulate(imgs[0], t, oh, ow))
    assert (px[0] != model_bf).mean() < 1e-3                  # fp32 summation order only
    single = _gpu(imgs[-1:], size, impl="mma").float().cpu().numpy()
    assert np.array_equal(single[0], got[-1])
    assert torch.equal(_gpu(imgs, size), _gpu(imgs, size, impl="mma"))           # "auto" == this kernel


def test_mma_pass_large_batches_flat_images_mean_std():
    """Batches that do not divide evenly over the CTAs (ranges start mid-image), one image, flat images,
    mean / std, and a loud refusal for a geometry the kernel does not take."""
    from skin_image_analysis_b200 import _lib, ops
    rng = np.random.default_rng(5)
    base = [helpers.synthetic_u8_image(450, 600, 400 + i, "noise") for i in range(4)]
    ref = _gpu(base, (224, 224), impl="mma")
    for nb in (1, 11, 170, 297):
        idx = torch.from_numpy(rng.integers(0, 4, nb)).cuda()
        x = torch.from_numpy(np.stack(base)).cuda()[idx].contiguous()
        assert torch.equal(ops.preprocess_u8hwc(x, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="mma"), ref[idx]), nb
    for v in (0, 1, 127, 200, 255):                           # flat images: exactly zero stays zero, the rest within one ulp
        const = np.full((450, 600, 3), v, np.uint8)           # (the fp16 weights sum to 1 +- 3e-4 and are not renormalised)
        out = _gpu([const], (224, 224), impl="mma")[0, :, 1:225, :3].float().cpu().numpy()
        exact = torch.tensor(np.float32(v) / 255.0).to(torch.bfloat16).float().item()
        assert np.all(np.abs(out - exact) <= exact * 2.0 ** -7), v
    lib = _lib.load()
    try:                                                      # the 4-warp instance gives the same bits as the 8-warp default
        assert lib.sia_debug_set_mma_warps(4) == 0
        assert torch.equal(_gpu(base, (224, 224), impl="mma"), ref)
    finally:
        assert lib.sia_debug_set_mma_warps(8) == 0
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    x = torch.from_numpy(np.stack(base[:2])).cuda()
    got = ops.preprocess_u8hwc(x, (224, 224), ops.LAYOUT_NHWC4_BF16, mean=mean, std=std, impl="mma").float().cpu().numpy()
    want = np.stack([R.transform_u8(im, (224, 224)) for im in base[:2]]).transpose(0, 2, 3, 1)
    want = (want - np.array(mean, np.float32)) / np.array(std, np.float32)
    assert np.abs(got[:, :, 1:225, :3] - want).max() <= 2.0 ** -6 + 3e-3       # bf16 rounding at |v| <= 2.7 + fp16 V
    assert np.all(got[:, :, 0] == 0) and np.all(got[:, :, 225:] == 0) and np.all(got[..., 3] == 0)   # pads stay zero
    with pytest.raises(_lib.SiaError):                        # 131-pixel rows are not a multiple of 8
        _gpu([helpers.synthetic_u8_image(97, 131, 8, "noise")], (64, 64), impl="mma")


def test_mma_pass_feeds_the_engine_bit_identically_across_slots():
    """The engine's graph-captured launches of this kernel: the four input slots give the same activations for the same
    bytes (no state leaks between launches, ring reuse across CTA ranges)."""
    from skin_image_analysis_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    u8 = torch.randint(0, 256, (64, 450, 600, 3), dtype=torch.uint8, device="cuda", generator=g)
    a = ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="mma")
    out = torch.full_like(a, 7.0)
    b = ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="mma", out=out)
    assert torch.equal(a, b)
    old = ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="tensor_core2").float()
    assert bool(((a.float() - old).abs() <= old.abs().clamp_min(2.0 ** -126) * 2.0 ** -7).all())
