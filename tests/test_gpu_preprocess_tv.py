"""SURVEY 8(f) row 2 parity: sia_preprocess_tv_u8hwc (the ToneClassifier test transform: v2.Resize on uint8 ->
ToDtype(scale) -> Normalize, notebooks/ToneClassifier/CNNTrialDataset.py:71-76) vs the oracle and the
reference-generated fixture.  Integer resampling + table lookup: BIT-EXACT, no tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import resize_tv as R
from tests import helpers

pytestmark = pytest.mark.gpu

TV_CASES = [("noise_450x600", 450, 600, 41, "noise"), ("smooth_450x600", 450, 600, 42, "smooth"),
            ("extremes_450x600", 450, 600, 43, "extremes"), ("noise_97x131", 97, 131, 44, "noise"),
            ("noise_600x450", 600, 450, 45, "noise"), ("noise_160x200_up", 160, 200, 46, "noise")]


def _gpu(u8, size, layout, **kw):
    from skin_image_analysis_b200 import ops
    return ops.preprocess_tv_u8hwc(torch.from_numpy(u8).cuda(), size, layout, **kw)


@pytest.mark.parametrize("name,h,w,seed,kind", TV_CASES)
def test_reference_fixture_bit_exact(golden_dir, name, h, w, seed, kind):
    from skin_image_analysis_b200 import ops
    g = np.load(os.path.join(golden_dir, "transform_tv.npz"))
    u8 = helpers.synthetic_u8_image(h, w, seed, kind)
    got = _gpu(u8[None], (224, 224), ops.LAYOUT_NCHW_F32)[0].cpu().numpy()
    assert np.array_equal(got[:, ::5, ::3], g[name + "_sub"])
    assert np.array_equal(got, R.transform_u8_chw(u8, (224, 224)))


@pytest.mark.parametrize("shape,out", [((450, 600), (224, 224)), ((450, 600), (512, 512)), ((33, 47), (20, 31)),
                                       ((64, 64), (64, 64)), ((50, 80), (224, 224)), ((301, 203), (149, 224)),
                                       ((5, 4), (16, 16)), ((224, 300), (224, 224)), ((1024, 768), (224, 224))])
def test_batch_and_odd_sizes_bit_exact(shape, out):
    """Batches of images whose rows / image strides are not 16-byte multiples (every head / body / tail path of the
    staging copy), identity axes, up-sampling and tiny sources."""
    from skin_image_analysis_b200 import ops
    rng = np.random.default_rng(11)
    u8 = rng.integers(0, 256, (3,) + shape + (3,), dtype=np.uint8)
    got = _gpu(u8, out, ops.LAYOUT_NCHW_F32).cpu().numpy()
    for n in range(3):
        assert np.array_equal(got[n], R.transform_u8_chw(u8[n], out)), n
    # a view that starts in the middle of the buffer (image 1 onwards): misaligned source base
    x = torch.from_numpy(u8).cuda()
    got1 = ops.preprocess_tv_u8hwc(x[1:], out, ops.LAYOUT_NCHW_F32).cpu().numpy()
    assert np.array_equal(got1, got[1:])


def test_bf16_layouts_are_the_rounded_float_tensor():
    from skin_image_analysis_b200 import ops
    u8 = np.stack([helpers.synthetic_u8_image(450, 600, 300 + i, k) for i, k in enumerate(["smooth", "noise"])])
    want = torch.from_numpy(np.stack([R.transform_u8_chw(im) for im in u8])).to(torch.bfloat16)
    nchw = _gpu(u8, (224, 224), ops.LAYOUT_NCHW_BF16).cpu()
    assert torch.equal(nchw, want)
    x4 = _gpu(u8, (224, 224), ops.LAYOUT_NHWC4_BF16).cpu()
    assert x4.shape == (2, 224, 232, 4)
    assert torch.equal(x4[:, :, 1:225, :3].permute(0, 3, 1, 2), want)
    assert float(x4[:, :, 0].abs().max()) == 0 and float(x4[:, :, 225:].abs().max()) == 0
    assert float(x4[..., 3].abs().max()) == 0


def test_custom_mean_std_and_out_buffer():
    from skin_image_analysis_b200 import ops
    u8 = helpers.synthetic_u8_image(120, 90, 5, "noise")[None]
    out = torch.full((1, 3, 64, 48), 7.0, device="cuda")
    ret = ops.preprocess_tv_u8hwc(torch.from_numpy(u8).cuda(), (64, 48), ops.LAYOUT_NCHW_F32, mean=(0.1, 0.2, 0.3),
                                  std=(0.5, 1.0, 2.0), out=out)
    assert ret is out
    assert np.array_equal(out[0].cpu().numpy(), R.transform_u8_chw(u8[0], (64, 48), (0.1, 0.2, 0.3), (0.5, 1.0, 2.0)))
    with pytest.raises(ValueError):
        ops.preprocess_tv_u8hwc(torch.from_numpy(u8).cuda(), (64, 48), out=torch.empty((1, 3, 64, 47), device="cuda"))
    with pytest.raises(Exception):
        ops.preprocess_tv_u8hwc(torch.from_numpy(u8), (64, 48))                 # CPU tensor: no fallback


def test_model_consumes_the_tv_layout():
    """The NHWC4 output feeds the same conv path (the ToneClassifier CNN is out of scope; this checks the layout)."""
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200.engine import plan_from_state_dict
    from skin_image_analysis_b200.synthetic import random_state_dict
    u8 = np.stack([helpers.synthetic_u8_image(450, 600, 400 + i, "smooth") for i in range(4)])
    plan = plan_from_state_dict(random_state_dict("SkinCancerListModel", 224, seed=3), torch.device("cuda"))
    x4 = _gpu(u8, (224, 224), ops.LAYOUT_NHWC4_BF16)
    logp, pred = plan.forward_nhwc4(x4)
    torch.cuda.synchronize()
    assert logp.shape == (4, 2) and torch.isfinite(logp).all()


@pytest.mark.parametrize("shape", [(450, 600), (97, 131), (33, 47)])
def test_planar_chw_input_like_read_image(shape):
    """The drop-in surface: uint8 CHW tensors (torchvision.io.read_image, CNNTrialDataset.py:93-95)."""
    from skin_image_analysis_b200.cnn_trial_dataset import TestTransforms
    rng = np.random.default_rng(21)
    u8 = rng.integers(0, 256, (3,) + shape + (3,), dtype=np.uint8)
    chw = torch.from_numpy(u8).permute(0, 3, 1, 2).contiguous().cuda()
    tf = TestTransforms()
    got = tf(chw).cpu().numpy()
    for n in range(3):
        assert np.array_equal(got[n], R.transform_u8_chw(u8[n])), n
    one = tf(chw[2]).cpu().numpy()                                   # a single [3,H,W] image, misaligned base
    assert one.shape == (3, 224, 224) and np.array_equal(one, got[2])
    assert np.array_equal(tf.hwc(torch.from_numpy(u8).cuda()).cpu().numpy(), got)
    with pytest.raises(TypeError):
        tf(chw.float())


def test_idp2a_instance_equals_bytewise_kernel():
    """The IDP.2A instance (interleaved source, rows % 4 == 0, 7 x 7 taps) and the byte-wise kernel are the same
    integer arithmetic: bit-identical outputs in every layout, on a batch with extreme bytes."""
    from skin_image_analysis_b200 import _lib, ops
    rng = np.random.default_rng(31)
    u8 = rng.integers(0, 256, (5, 450, 600, 3), dtype=np.uint8)
    u8[3] = helpers.synthetic_u8_image(450, 600, 43, "extremes")
    u8[4] = 255
    x = torch.from_numpy(u8).cuda()
    lib = _lib.load()
    for layout in (ops.LAYOUT_NCHW_F32, ops.LAYOUT_NCHW_BF16, ops.LAYOUT_NHWC4_BF16):
        fast = ops.preprocess_tv_u8hwc(x, (224, 224), layout)
        try:
            lib.sia_debug_tv_force_generic(1)
            slow = ops.preprocess_tv_u8hwc(x, (224, 224), layout)
        finally:
            lib.sia_debug_tv_force_generic(0)
        assert torch.equal(fast, slow), layout
    got = ops.preprocess_tv_u8hwc(x, (224, 224), ops.LAYOUT_NCHW_F32).cpu().numpy()
    for n in (0, 3, 4):
        assert np.array_equal(got[n], R.transform_u8_chw(u8[n])), n
    # another size that takes the IDP.2A path (7 x 7 taps, 4-byte rows), tile boundaries elsewhere
    v8 = rng.integers(0, 256, (2, 500, 560, 3), dtype=np.uint8)
    g2 = ops.preprocess_tv_u8hwc(torch.from_numpy(v8).cuda(), (224, 224), ops.LAYOUT_NCHW_F32).cpu().numpy()
    for n in range(2):
        assert np.array_equal(g2[n], R.transform_u8_chw(v8[n])), n
