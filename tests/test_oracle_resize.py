"""Oracle (oracle/resize.py) vs scipy.ndimage and the reference-generated fixtures -- CPU only."""
import os

import numpy as np
import pytest

from oracle import resize as R
from tests import helpers

CASES = [("noise_450x600_224", 450, 600, 31, "noise", (224, 224)),
         ("smooth_450x600_224", 450, 600, 32, "smooth", (224, 224)),
         ("extremes_450x600_224", 450, 600, 33, "extremes", (224, 224)),
         ("noise_450x600_512", 450, 600, 34, "noise", (512, 512)),
         ("noise_97x131_int64", 97, 131, 35, "noise", 64),
         ("noise_131x97_int48", 131, 97, 36, "noise", 48)]


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "transform.npz"))


@pytest.mark.parametrize("name,h,w,seed,kind,size", CASES)
def test_transform_matches_reference_fixture(golden, name, h, w, seed, kind, size):
    """Rescale + ToTensor run by the reference itself (tests/golden/make_golden.py)."""
    u8 = helpers.synthetic_u8_image(h, w, seed, kind)
    out = R.transform_u8(u8, size)
    assert tuple(golden[name + "_shape"]) == out.shape
    ref = golden[name + "_full"] if name + "_full" in golden else golden[name + "_sub"]
    got = out if name + "_full" in golden else out[:, ::7, ::5]
    # numpy restatement vs scipy: identical up to one float32 ulp of rounding in the fp64 sums
    np.testing.assert_allclose(got, ref, rtol=0, atol=6e-8)
    assert abs(out.astype(np.float64).sum() - float(golden[name + "_sum"][0])) < 1e-3


@pytest.mark.parametrize("shape,out", [((450, 600), (224, 224)), ((450, 600), (512, 512)), ((33, 47), (20, 31)),
                                       ((64, 64), (64, 64)), ((50, 80), (224, 224)), ((300, 200), (149, 224))])
def test_numpy_restatement_equals_scipy(shape, out):
    rng = np.random.default_rng(5)
    img = R.u8_to_unit_float(rng.integers(0, 256, shape + (3,), dtype=np.uint8))
    a, b = R.resize(img, out), R.resize_scipy(img, out)
    assert a.dtype == b.dtype == np.float32 and a.shape == b.shape
    np.testing.assert_allclose(a, b, rtol=0, atol=6e-8)


def test_rescale_int_keeps_aspect_like_reference():
    assert R.rescale_size(450, 600, 224) == (224, 298)        # short side matched, int() truncation
    assert R.rescale_size(600, 450, 224) == (298, 224)
    assert R.rescale_size(450, 600, (224, 224)) == (224, 224)
    assert R.rescale_size(97, 131, 64) == (64, int(64 * 131 / 97))


def test_clip_is_a_noop_and_range_is_preserved():
    for seed, kind in [(1, "noise"), (2, "extremes"), (3, "smooth")]:
        img = R.u8_to_unit_float(helpers.synthetic_u8_image(450, 600, seed, kind))
        out = R.resize(img, (224, 224))
        assert out.min() >= img.min() and out.max() <= img.max()


def test_constant_image_is_a_fixed_point():
    img = np.full((450, 600, 3), np.float32(200 / 255.0))
    out = R.resize(img, (224, 224))
    assert np.all(out == img[0, 0, 0])


def test_weight_matrix_composition_matches():
    img = R.u8_to_unit_float(helpers.synthetic_u8_image(450, 600, 8, "noise")).astype(np.float64)
    wy, wx = R.axis_weight_matrix(450, 224, True), R.axis_weight_matrix(600, 224, True)
    comp = np.einsum("jw,iwc->ijc", wx, np.tensordot(wy, img, axes=(1, 0)), optimize=True)
    np.testing.assert_allclose(comp, R.resize(img.astype(np.float32), (224, 224)), rtol=0, atol=2e-7)
    np.testing.assert_allclose(wy.sum(1), 1.0, atol=1e-12)
    assert (np.abs(wy) > 0).sum(1).max() == 6 and (np.abs(wx) > 0).sum(1).max() == 8
