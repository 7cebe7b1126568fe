"""K1-K3 parity: sia_preprocess_u8hwc vs the oracle transform (and the reference-generated fixtures)."""
import os

import numpy as np
import pytest
import torch

from oracle import resize as R
from tests import helpers

pytestmark = pytest.mark.gpu

F32_TOL = 1e-6          # fp32 output vs the oracle's fp64-accumulate result


def _gpu(u8_list, size, layout, **kw):
    from skin_image_analysis_b200 import ops
    x = torch.from_numpy(np.stack(u8_list)).cuda()
    return ops.preprocess_u8hwc(x, size, layout, **kw)


@pytest.mark.parametrize("kind", ["noise", "smooth", "extremes"])
def test_f32_nchw_matches_oracle_224(kind):
    from skin_image_analysis_b200 import ops
    imgs = [helpers.synthetic_u8_image(450, 600, 100 + i, kind) for i in range(3)]
    got = _gpu(imgs, (224, 224), ops.LAYOUT_NCHW_F32).cpu().numpy()
    for i, im in enumerate(imgs):
        want = R.transform_u8(im, (224, 224))
        assert got[i].shape == want.shape
        assert np.abs(got[i] - want).max() <= F32_TOL


def test_reference_fixture_cases(golden_dir):
    """The same cases the reference's Rescale+ToTensor produced in tests/golden/transform.npz."""
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200.resize_weights import rescale_size
    g = np.load(os.path.join(golden_dir, "transform.npz"))
    cases = [("noise_450x600_224", 450, 600, 31, "noise", (224, 224)),
             ("noise_450x600_512", 450, 600, 34, "noise", (512, 512)),
             ("noise_97x131_int64", 97, 131, 35, "noise", 64),
             ("noise_131x97_int48", 131, 97, 36, "noise", 48)]
    for name, h, w, seed, kind, size in cases:
        u8 = helpers.synthetic_u8_image(h, w, seed, kind)
        oh, ow = rescale_size(h, w, size)
        got = _gpu([u8], (oh, ow), ops.LAYOUT_NCHW_F32)[0].cpu().numpy()
        assert tuple(g[name + "_shape"]) == got.shape
        ref = g[name + "_full"] if name + "_full" in g else g[name + "_sub"]
        sub = got if name + "_full" in g else got[:, ::7, ::5]
        assert np.abs(sub - ref).max() <= F32_TOL, name


@pytest.mark.parametrize("fixed_point", [False, True])
def test_bf16_layouts_within_one_ulp(fixed_point):
    """bf16 outputs: fp32 arithmetic (fixed_point=False) and the 15-bit integer-dot-product horizontal
    pass (the default for bf16) both stay within one bf16 ulp of the bf16-rounded oracle."""
    from skin_image_analysis_b200 import ops
    imgs = [helpers.synthetic_u8_image(450, 600, 200 + i, k) for i, k in enumerate(["smooth", "noise"])]
    want = np.stack([R.transform_u8(im, (224, 224)) for im in imgs])
    want_bf = torch.from_numpy(want).to(torch.bfloat16).float().numpy()
    nchw = _gpu(imgs, (224, 224), ops.LAYOUT_NCHW_BF16, fixed_point=fixed_point).float().cpu().numpy()
    nhwc4 = _gpu(imgs, (224, 224), ops.LAYOUT_NHWC4_BF16, fixed_point=fixed_point,
                 impl="cuda_core").float().cpu().numpy()
    assert nhwc4.shape == (2, 224, 224 + ops.NHWC4_PAD, 4) and np.all(nhwc4[..., 3] == 0)
    assert np.all(nhwc4[:, :, 0] == 0) and np.all(nhwc4[:, :, 225:] == 0)        # zero pad columns
    assert np.array_equal(nhwc4[:, :, 1:225, :3].transpose(0, 3, 1, 2), nchw)
    ulp = np.maximum(np.abs(want_bf), 2.0 ** -126) * 2.0 ** -7
    assert np.all(np.abs(nchw - want_bf) <= ulp)
    # before rounding the two arithmetic variants differ from the oracle by <= 1e-6 / <= 2.5e-4 of full scale
    assert np.abs(nchw - want).max() <= 2.0 ** -8 + (2.5e-4 if fixed_point else 1e-6)
    assert (nchw != want_bf).mean() < (0.25 if fixed_point else 0.01)


def test_fixed_point_pass_keeps_flat_images_exact():
    from skin_image_analysis_b200 import ops
    for v in (0, 1, 127, 200, 255):
        const = np.full((450, 600, 3), v, np.uint8)
        out = _gpu([const], (224, 224), ops.LAYOUT_NCHW_BF16)[0].float().cpu().numpy()
        want = torch.tensor(np.float32(v) / 255.0).to(torch.bfloat16).float().item()
        assert np.all(out == want), v


def test_mean_std_and_rows_per_cta_invariance():
    from skin_image_analysis_b200 import ops
    im = helpers.synthetic_u8_image(450, 600, 7, "noise")
    base = _gpu([im], (224, 224), ops.LAYOUT_NCHW_F32)[0]
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    norm = _gpu([im], (224, 224), ops.LAYOUT_NCHW_F32, mean=mean, std=std)[0]
    m = torch.tensor(mean, device="cuda").view(3, 1, 1)
    s = torch.tensor(std, device="cuda").view(3, 1, 1)
    assert torch.allclose(norm, (base - m) / s, atol=2e-6, rtol=0)
    for rows in (1, 7, 16, 224):
        assert torch.equal(_gpu([im], (224, 224), ops.LAYOUT_NCHW_F32, rows_per_cta=rows)[0], base)


def test_constant_image_and_batch_independence():
    from skin_image_analysis_b200 import ops
    const = np.full((450, 600, 3), 200, np.uint8)
    noise = helpers.synthetic_u8_image(450, 600, 9, "noise")
    out = _gpu([const, noise, const], (224, 224), ops.LAYOUT_NCHW_F32)
    assert np.abs(out[0].cpu().numpy() - np.float32(200 / 255.0)).max() <= 2e-7
    assert torch.equal(out[0], out[2])
    assert torch.equal(out[1], _gpu([noise], (224, 224), ops.LAYOUT_NCHW_F32)[0])


def test_rescale_dropin_tuple_semantics():
    from skin_image_analysis_b200.tone_bias_dataset import Rescale, ToTensor
    u8 = helpers.synthetic_u8_image(97, 131, 35, "noise")
    sample = (np.float32(u8) / 255.0, 1, 42)
    img, label, idx = Rescale(64)(sample)
    assert (label, idx) == (1, 42) and img.shape == (64, 86, 3) and img.dtype == np.float32
    t, _, _ = ToTensor()((img, label, idx))
    assert np.abs(t.numpy() - R.transform_u8(u8, 64)).max() <= F32_TOL


@pytest.mark.parametrize("src_hw,size,batch", [((450, 600), (224, 224), 5), ((480, 640), (224, 224), 2),
                                               ((450, 600), (512, 512), 2), ((300, 400), (224, 224), 3)])
def test_tensor_core_pass_matches_oracle(src_hw, size, batch):
    """csrc/preprocess_tc.cu (vertical pass as a tcgen05 GEMM): within one bf16 ulp of the bf16-rounded
    oracle, <= 2.5e-4 of full scale before rounding, bit-identical to the numpy model of its arithmetic
    up to fp32 summation order, and batch / CTA-assignment independent."""
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200 import resize_weights as rw
    kinds = ["noise", "smooth", "extremes", "noise", "smooth"]
    imgs = [helpers.synthetic_u8_image(src_hw[0], src_hw[1], 300 + i, kinds[i]) for i in range(batch)]
    got = _gpu(imgs, size, ops.LAYOUT_NHWC4_BF16, impl="tensor_core").float().cpu().numpy()
    assert got.shape == (batch, size[0], size[1] + ops.NHWC4_PAD, 4)
    assert np.all(got[..., 3] == 0) and np.all(got[:, :, 0] == 0) and np.all(got[:, :, size[1] + 1:] == 0)
    want = np.stack([R.transform_u8(im, size) for im in imgs]).transpose(0, 2, 3, 1)
    want_bf = torch.from_numpy(want).to(torch.bfloat16).float().numpy()
    px = got[:, :, 1:size[1] + 1, :3]
    ulp = np.maximum(np.abs(want_bf), 2.0 ** -126) * 2.0 ** -7
    assert np.all(np.abs(px - want_bf) <= ulp)
    assert np.abs(px - want).max() <= 2.0 ** -8 + 2.5e-4
    assert (px != want_bf).mean() < 0.05
    t = rw.build_tc_tables(src_hw[0], src_hw[1], size[0], size[1])
    model = rw.tc_emulate(imgs[0], t, size[0], size[1])
    model_bf = torch.from_numpy(model).to(torch.bfloat16).float().numpy()
    assert (px[0] != model_bf).mean() < 1e-3            # fp32 summation order only
    old = _gpu(imgs, size, ops.LAYOUT_NHWC4_BF16, impl="cuda_core").float().cpu().numpy()
    assert np.all(np.abs(got - old) <= np.maximum(np.abs(old), 2.0 ** -126) * 2.0 ** -7)
    single = _gpu(imgs[-1:], size, ops.LAYOUT_NHWC4_BF16, impl="tensor_core").float().cpu().numpy()
    assert np.array_equal(single[0], got[-1])


def test_tensor_core_pass_large_batch_and_mean_std():
    """More images than CTAs per tile (every CTA loops), non-trivial mean / std, flat images exact."""
    from skin_image_analysis_b200 import ops
    rng = np.random.default_rng(5)
    base = [helpers.synthetic_u8_image(450, 600, 400 + i, "noise") for i in range(4)]
    idx = rng.integers(0, 4, 170)
    x = torch.from_numpy(np.stack(base)).cuda()[torch.from_numpy(idx).cuda()]
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    got = ops.preprocess_u8hwc(x, (224, 224), ops.LAYOUT_NHWC4_BF16, mean=mean, std=std, impl="tensor_core")
    ref = ops.preprocess_u8hwc(torch.from_numpy(np.stack(base)).cuda(), (224, 224), ops.LAYOUT_NHWC4_BF16,
                               mean=mean, std=std, impl="tensor_core")
    assert torch.equal(got, ref[torch.from_numpy(idx).cuda()])
    want = np.stack([R.transform_u8(im, (224, 224)) for im in base]).transpose(0, 2, 3, 1)
    want = (want - np.float32(mean)) / np.float32(std)
    px = ref[:, :, 1:225, :3].float().cpu().numpy()
    assert np.abs(px - want).max() <= 2.0 ** -6 + 2e-3          # bf16 rounding at |v| <= 2.7 + the fp16 weights
    for v in (0, 1, 127, 200, 255):
        const = np.full((450, 600, 3), v, np.uint8)
        out = _gpu([const], (224, 224), ops.LAYOUT_NHWC4_BF16, impl="tensor_core")[0, :, 1:225, :3].float().cpu().numpy()
        assert np.all(out == torch.tensor(np.float32(v) / 255.0).to(torch.bfloat16).float().item()), v


@pytest.mark.parametrize("src_hw,batch", [((450, 600), 5), ((480, 640), 3)])
def test_two_product_tensor_core_pass_matches_oracle(src_hw, batch):
    """csrc/preprocess_tc2.cu (impl="tensor_core2": the horizontal pass is a second tcgen05.mma whose A operand is the
    fp16-repacked accumulator of the vertical product, read from tensor memory): within one bf16 ulp of the
    bf16-rounded oracle, <= 2.5e-4 of full scale before rounding, bit-identical to the numpy model of its arithmetic
    up to fp32 summation order, within one ulp of the one-product kernel, zero pads, batch independent."""
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200 import resize_weights as rw
    kinds = ["noise", "smooth", "extremes", "noise", "smooth"]
    imgs = [helpers.synthetic_u8_image(src_hw[0], src_hw[1], 300 + i, kinds[i]) for i in range(batch)]
    got = _gpu(imgs, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="tensor_core2").float().cpu().numpy()
    assert got.shape == (batch, 224, 224 + ops.NHWC4_PAD, 4)
    assert np.all(got[..., 3] == 0) and np.all(got[:, :, 0] == 0) and np.all(got[:, :, 225:] == 0)
    want = np.stack([R.transform_u8(im, (224, 224)) for im in imgs]).transpose(0, 2, 3, 1)
    want_bf = torch.from_numpy(want).to(torch.bfloat16).float().numpy()
    px = got[:, :, 1:225, :3]
    ulp = np.maximum(np.abs(want_bf), 2.0 ** -126) * 2.0 ** -7
    assert np.all(np.abs(px - want_bf) <= ulp)
    assert np.abs(px - want).max() <= 2.0 ** -8 + 2.6e-4
    assert (px != want_bf).mean() < 0.05                      # V is rounded to fp16: ~2 % of the outputs round the other way
    t = rw.build_tc2_tables(src_hw[0], src_hw[1], 224, 224)
    model_bf = torch.from_numpy(rw.tc2_emulate(imgs[0], t, 224, 224)).to(torch.bfloat16).float().numpy()
    assert (px[0] != model_bf).mean() < 1e-3                  # fp32 summation order only
    old = _gpu(imgs, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="tensor_core").float().cpu().numpy()
    assert np.all(np.abs(got - old) <= np.maximum(np.abs(old), 2.0 ** -126) * 2.0 ** -7)
    single = _gpu(imgs[-1:], (224, 224), ops.LAYOUT_NHWC4_BF16, impl="tensor_core2").float().cpu().numpy()
    assert np.array_equal(single[0], got[-1])


def test_two_product_pass_large_batches_and_flat_images():
    """More images than CTAs per tile (pairs in flight, a single trailing image), flat images exact, unsupported
    geometries refused loudly."""
    from skin_image_analysis_b200 import _lib, ops
    rng = np.random.default_rng(5)
    base = [helpers.synthetic_u8_image(450, 600, 400 + i, "noise") for i in range(4)]
    ref = ops.preprocess_u8hwc(torch.from_numpy(np.stack(base)).cuda(), (224, 224), ops.LAYOUT_NHWC4_BF16,
                               impl="tensor_core2")
    for nb in (170, 297):
        idx = torch.from_numpy(rng.integers(0, 4, nb)).cuda()
        x = torch.from_numpy(np.stack(base)).cuda()[idx].contiguous()
        assert torch.equal(ops.preprocess_u8hwc(x, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="tensor_core2"), ref[idx])
    for v in (0, 1, 127, 200, 255):
        const = np.full((450, 600, 3), v, np.uint8)
        out = _gpu([const], (224, 224), ops.LAYOUT_NHWC4_BF16, impl="tensor_core2")[0, :, 1:225, :3].float().cpu().numpy()
        assert np.all(out == torch.tensor(np.float32(v) / 255.0).to(torch.bfloat16).float().item()), v
    with pytest.raises(_lib.SiaError):                        # 512 x 512 outputs do not fit the 21-slot layout
        _gpu([base[0]], (512, 512), ops.LAYOUT_NHWC4_BF16, impl="tensor_core2")


def test_auto_falls_back_when_the_tensor_core_kernels_do_not_fit():
    """impl="auto" on the NHWC4 layout: a geometry no tensor-core kernel takes (131-pixel rows are not a multiple
    of 8 bytes) goes to the CUDA-core kernel, a 512 x 512 output to the warp-MMA kernel (the two-product tcgen05 kernel
    refuses it: too many columns per block); asking for a kernel explicitly still refuses loudly."""
    from skin_image_analysis_b200 import _lib, ops
    odd = helpers.synthetic_u8_image(97, 131, 8, "noise")
    auto = _gpu([odd], (64, 64), ops.LAYOUT_NHWC4_BF16)
    assert torch.equal(auto, _gpu([odd], (64, 64), ops.LAYOUT_NHWC4_BF16, impl="cuda_core"))
    with pytest.raises(_lib.SiaError):
        _gpu([odd], (64, 64), ops.LAYOUT_NHWC4_BF16, impl="tensor_core")
    big = helpers.synthetic_u8_image(450, 600, 9, "smooth")
    auto = _gpu([big], (512, 512), ops.LAYOUT_NHWC4_BF16)
    assert torch.equal(auto, _gpu([big], (512, 512), ops.LAYOUT_NHWC4_BF16, impl="mma"))
    with pytest.raises(_lib.SiaError):
        _gpu([big], (512, 512), ops.LAYOUT_NHWC4_BF16, impl="tensor_core2")
