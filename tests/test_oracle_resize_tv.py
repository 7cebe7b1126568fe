"""Oracle of the ToneClassifier transform variant (oracle/resize_tv.py) -- CPU only.

Pinned three ways: the committed fixture made by the reference's own ``ISIC(..., "Test").transforms`` Compose
(notebooks/ToneClassifier/CNNTrialDataset.py:71-76; tests/golden/make_golden.py::gen_transform_tv), the torch /
torchvision that ship in this image run on the same inputs, and the host tables the CUDA kernel consumes."""
import os

import numpy as np
import pytest
import torch

from oracle import resize_tv as R
from skin_image_analysis_b200 import resize_weights as rw
from tests import helpers

TV_CASES = [("noise_450x600", 450, 600, 41, "noise"), ("smooth_450x600", 450, 600, 42, "smooth"),
            ("extremes_450x600", 450, 600, 43, "extremes"), ("noise_97x131", 97, 131, 44, "noise"),
            ("noise_600x450", 600, 450, 45, "noise"), ("noise_160x200_up", 160, 200, 46, "noise")]


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "transform_tv.npz"))


@pytest.mark.parametrize("name,h,w,seed,kind", TV_CASES)
def test_matches_reference_fixture_bit_exactly(golden, name, h, w, seed, kind):
    out = R.transform_u8_chw(helpers.synthetic_u8_image(h, w, seed, kind), (224, 224))
    assert out.shape == (3, 224, 224) and out.dtype == np.float32
    assert np.array_equal(out[:, ::5, ::3], golden[name + "_sub"])            # integer resize + same float32 ops
    assert abs(out.astype(np.float64).sum() - float(golden[name + "_sum"][0])) < 1e-6


@pytest.mark.parametrize("shape,out", [((450, 600), (224, 224)), ((450, 600), (512, 512)), ((33, 47), (20, 31)),
                                       ((64, 64), (64, 64)), ((50, 80), (224, 224)), ((300, 200), (149, 224)),
                                       ((5, 4), (16, 16)), ((224, 300), (224, 224))])
def test_uint8_resize_equals_torch(shape, out):
    """ATen's uint8 antialias bilinear resampler (what v2.Resize calls) on the same bytes: bit-exact."""
    img = np.random.default_rng(7).integers(0, 256, shape + (3,), dtype=np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1)[None].contiguous()
    ref = torch.nn.functional.interpolate(t, size=out, mode="bilinear", antialias=True)[0].permute(1, 2, 0).numpy()
    assert np.array_equal(R.resize_u8(img, out), ref)


def test_whole_transform_equals_torchvision():
    tv = pytest.importorskip("torchvision.transforms.v2")
    compose = tv.Compose([tv.Resize((224, 224)), tv.ToDtype(torch.float32, scale=True),
                          tv.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    img = helpers.synthetic_u8_image(450, 600, 9, "smooth")
    ref = compose(torch.from_numpy(img).permute(2, 0, 1).contiguous()).numpy()
    assert np.array_equal(R.transform_u8_chw(img), ref)


@pytest.mark.parametrize("n_in,n_out", [(600, 224), (450, 224), (450, 512), (224, 224), (4, 16), (131, 224), (1000, 7)])
def test_kernel_tables_equal_oracle_tables(n_in, n_out):
    """resize_weights.tv_axis (what the kernel runs on) against the oracle's taps: same windows after undoing the
    edge shift, same int16 weights, same precision; windows never leave the source."""
    ax = rw.tv_axis(n_in, n_out)
    xmin, xsize, wi, prec = R.axis_tables(n_in, n_out)
    assert np.all(ax.xmin >= 0) and np.all(ax.xmin + ax.taps <= n_in)
    if n_in == n_out:
        assert ax.taps == 1 and np.all(ax.w == 1 << ax.precision)
        return
    assert ax.precision == prec
    for i in range(n_out):
        dense_a = np.zeros(n_in, np.int64)
        dense_a[ax.xmin[i]:ax.xmin[i] + ax.taps] = ax.w[i]
        dense_b = np.zeros(n_in, np.int64)
        dense_b[xmin[i]:xmin[i] + xsize[i]] = wi[i, :xsize[i]]
        assert np.array_equal(dense_a, dense_b)


def test_lut_is_the_float32_normalisation():
    lut = rw.tv_normalise_lut(R.IMAGENET_MEAN, R.IMAGENET_STD)
    b = torch.arange(256, dtype=torch.uint8)
    for c in range(3):
        x = b.to(torch.float32).mul_(1.0 / 255.0)
        ref = (x - torch.tensor(R.IMAGENET_MEAN[c], dtype=torch.float32)) / torch.tensor(R.IMAGENET_STD[c], dtype=torch.float32)
        assert np.array_equal(lut[c], ref.numpy())
    assert np.array_equal(lut, R.normalise_lut())


def test_tile_plan_fits_shared_memory():
    y = rw.tv_axis(450, 224)
    tile, rows = rw.tv_tile_plan(y, 1800, 224, 7)
    assert tile >= 1 and rows >= y.taps
    for i0 in range(0, 224, tile):
        i1 = min(i0 + tile, 224)
        assert y.xmin[i1 - 1] + y.taps - y.xmin[i0] <= rows
    with pytest.raises(ValueError):
        rw.tv_tile_plan(rw.tv_axis(4000, 8), 3 * 40000, 8, 7)
