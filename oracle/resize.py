"""Oracle: the tone_bias transform, restated in numpy (TEST INFRASTRUCTURE ONLY).

Follows, stage by stage:

  * ``HibaDataset.__getitem__``  tone_bias_dataset.py:335   ``np.float32(u8) / 255.0``
  * ``Rescale.__call__``         tone_bias_dataset.py:411-427  size logic + ``skimage.transform.resize``
  * ``ToTensor.__call__``        tone_bias_dataset.py:464-473  HWC -> CHW

``skimage.transform.resize`` (scikit-image==0.24.0, requirements.txt:128; not vendored,
not installed here) is restated from its published algorithm with the defaults the
reference call site uses (order=1, mode='reflect', anti_aliasing=None, clip=True):

    factors = in_shape / out_shape
    if any axis shrinks:  sigma = max(0, (factors-1)/2)
                          img = ndi.gaussian_filter(img, sigma, mode='mirror', cval=0)   # truncate=4
    out = ndi.zoom(img, 1/factors, order=1, mode='mirror', cval=0, grid_mode=True)
    out = clip(out, img_in.min(), img_in.max())

and the two scipy.ndimage primitives (scipy==1.14.0, requirements.txt:130) are restated
below in plain numpy:

  * ``gaussian_filter``: per axis with sigma > 1e-15, radius = int(4*sigma + 0.5), weights
    exp(-x^2/(2 sigma^2)) normalised in float64, 1-D correlation with 'mirror' (reflect about
    the centre of the edge sample) boundary, accumulated in float64 and stored in the array
    dtype (float32) after EACH axis.
  * ``zoom(order=1, grid_mode=True, mode='mirror')``: output sample i reads the input at
    ``(i + 0.5) * in/out - 0.5``; the coordinate is folded by the mirror rule, then the two
    neighbours are blended linearly; separable product weights, float64 accumulate, float32 store.

``resize_scipy`` is the same function over the scipy.ndimage that ships in this image; the tests
require the two to agree (that is what pins this file).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "u8_to_unit_float", "rescale_size", "gaussian_kernel1d", "aa_sigma", "resize",
    "resize_scipy", "rescale", "to_tensor_chw", "transform_u8", "axis_weight_matrix",
]


def u8_to_unit_float(image_u8: np.ndarray) -> np.ndarray:
    """tone_bias_dataset.py:335 -- float32 image in [0,1]."""
    return np.float32(image_u8) / 255.0


def rescale_size(h: int, w: int, output_size) -> tuple[int, int]:
    """tone_bias_dataset.py:414-423 -- int => short side matched, aspect kept, int() truncation."""
    if isinstance(output_size, int):
        if h > w:
            new_h, new_w = output_size * h / w, output_size
        else:
            new_h, new_w = output_size, output_size * w / h
    else:
        new_h, new_w = output_size
    return int(new_h), int(new_w)


def gaussian_kernel1d(sigma: float, radius: int) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(order=0): float64, normalised to sum 1."""
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    phi = np.exp(-0.5 / (sigma * sigma) * x * x)
    return phi / phi.sum()


def aa_sigma(n_in: int, n_out: int) -> float:
    """skimage resize: anti_aliasing_sigma = max(0, (in/out - 1) / 2) per axis."""
    return max(0.0, (n_in / n_out - 1.0) / 2.0)


def _mirror_index(idx: np.ndarray, n: int) -> np.ndarray:
    """ndimage 'mirror': reflect about the centre of the edge samples (d c b | a b c d | c b a)."""
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    idx = np.abs(idx) % period
    return np.where(idx >= n, period - idx, idx)


def _correlate1d_mirror(a: np.ndarray, weights: np.ndarray, axis: int) -> np.ndarray:
    """1-D correlation along ``axis``; float64 accumulate, result cast to a.dtype."""
    radius = (len(weights) - 1) // 2
    n = a.shape[axis]
    a64 = np.moveaxis(a, axis, 0).astype(np.float64)
    acc = np.zeros_like(a64)
    base = np.arange(n)
    for j, wgt in enumerate(weights):
        acc += wgt * a64[_mirror_index(base + (j - radius), n)]
    return np.moveaxis(acc, 0, axis).astype(a.dtype)


def gaussian_filter_mirror(a: np.ndarray, sigmas) -> np.ndarray:
    out = a
    for axis, sigma in enumerate(sigmas):
        if sigma > 1e-15:
            radius = int(4.0 * float(sigma) + 0.5)
            out = _correlate1d_mirror(out, gaussian_kernel1d(float(sigma), radius), axis)
    return out


def _linear_taps(n_in: int, n_out: int):
    """Per output sample: the two (mirrored) source indices and their float64 weights."""
    i = np.arange(n_out, dtype=np.float64)
    cc = (i + 0.5) * (n_in / n_out) - 0.5 if n_in != n_out else i.copy()
    # mirror-fold the continuous coordinate into [0, n_in-1]
    if n_in > 1:
        period = 2.0 * (n_in - 1)
        cc = np.abs(cc) % period
        cc = np.where(cc > n_in - 1, period - cc, cc)
    else:
        cc = np.zeros_like(cc)
    i0 = np.floor(cc).astype(np.int64)
    t = cc - i0
    i1 = _mirror_index(i0 + 1, n_in)
    return i0, i1, 1.0 - t, t


def zoom_linear_grid(a: np.ndarray, out_hw: tuple[int, int]) -> np.ndarray:
    """ndi.zoom(order=1, grid_mode=True, mode='mirror') over the two leading axes."""
    h_in, w_in = a.shape[:2]
    h_out, w_out = out_hw
    y0, y1, wy0, wy1 = _linear_taps(h_in, h_out)
    x0, x1, wx0, wx1 = _linear_taps(w_in, w_out)
    a64 = a.astype(np.float64)
    sh = (h_out, w_out) + (1,) * (a.ndim - 2)
    wy0 = wy0.reshape(-1, 1, *([1] * (a.ndim - 2)))
    wy1 = wy1.reshape(-1, 1, *([1] * (a.ndim - 2)))
    wx0 = wx0.reshape(1, -1, *([1] * (a.ndim - 2)))
    wx1 = wx1.reshape(1, -1, *([1] * (a.ndim - 2)))
    del sh
    r0, r1 = a64[y0], a64[y1]
    out = (wy0 * wx0) * r0[:, x0] + (wy0 * wx1) * r0[:, x1] \
        + (wy1 * wx0) * r1[:, x0] + (wy1 * wx1) * r1[:, x1]
    return out.astype(a.dtype)


def resize(image: np.ndarray, output_shape: tuple[int, int]) -> np.ndarray:
    """``skimage.transform.resize(image, (h, w))`` with the reference's defaults
    (call site tone_bias_dataset.py:425).  ``image`` is HWC (or HW) float32."""
    image = np.asarray(image)
    if image.dtype != np.float32:           # the reference always feeds float32 (:335)
        image = image.astype(np.float64)
    h_in, w_in = image.shape[:2]
    h_out, w_out = int(output_shape[0]), int(output_shape[1])
    if (h_in, w_in) == (h_out, w_out):
        return image.copy()
    filtered = image
    if h_out < h_in or w_out < w_in:
        sig = [aa_sigma(h_in, h_out), aa_sigma(w_in, w_out)] + [0.0] * (image.ndim - 2)
        filtered = gaussian_filter_mirror(image, sig)
    out = zoom_linear_grid(filtered, (h_out, w_out))
    return np.clip(out, image.min(), image.max(), out=out)


def resize_scipy(image: np.ndarray, output_shape: tuple[int, int]) -> np.ndarray:
    """Same algorithm through the scipy.ndimage of this image -- pins ``resize``."""
    import scipy.ndimage as ndi
    image = np.asarray(image)
    out_shape = tuple(int(v) for v in output_shape) + image.shape[2:]
    if out_shape == image.shape:
        return image.copy()
    factors = np.divide(image.shape, out_shape)
    filtered = image
    if any(o < i for o, i in zip(out_shape, image.shape)):
        filtered = ndi.gaussian_filter(image, np.maximum(0, (factors - 1) / 2), cval=0, mode="mirror")
    out = ndi.zoom(filtered, [1 / f for f in factors], order=1, mode="mirror", cval=0, grid_mode=True)
    return np.clip(out, image.min(), image.max(), out=out)


def rescale(sample, output_size, resize_fn=resize):
    """``Rescale(output_size)(sample)`` -- tone_bias_dataset.py:411-427."""
    image, label, index = sample
    new_h, new_w = rescale_size(image.shape[0], image.shape[1], output_size)
    return resize_fn(image, (new_h, new_w)), label, index


def to_tensor_chw(sample):
    """``ToTensor()(sample)`` minus the torch wrapper -- tone_bias_dataset.py:470 (CHW view)."""
    image, label, index = sample
    return image.transpose((2, 0, 1)), label, index


def transform_u8(image_u8: np.ndarray, output_size=(224, 224), resize_fn=resize) -> np.ndarray:
    """u8 HWC decode buffer -> float32 CHW in [0,1]: the whole a1..a3 chain for one image."""
    img = u8_to_unit_float(image_u8)
    img, _, _ = rescale((img, 0, 0), output_size, resize_fn)
    return np.ascontiguousarray(img.transpose((2, 0, 1)))


def axis_weight_matrix(n_in: int, n_out: int, antialias: bool) -> np.ndarray:
    """Dense float64 [n_out, n_in] matrix of the composed (zoom o gaussian) operator along one
    axis, boundary folding included.  Used by tests to check the product's banded weight tables
    (skin_image_analysis_b200/resize_weights.py) -- not used by ``resize`` itself."""
    g = np.eye(n_in, dtype=np.float64)
    sigma = aa_sigma(n_in, n_out) if antialias else 0.0
    if sigma > 1e-15:
        radius = int(4.0 * sigma + 0.5)
        k = gaussian_kernel1d(sigma, radius)
        g = np.zeros((n_in, n_in), dtype=np.float64)
        base = np.arange(n_in)
        for j, wgt in enumerate(k):
            np.add.at(g, (base, _mirror_index(base + (j - radius), n_in)), wgt)
    i0, i1, w0, w1 = _linear_taps(n_in, n_out)
    z = np.zeros((n_out, n_in), dtype=np.float64)
    rows = np.arange(n_out)
    np.add.at(z, (rows, i0), w0)
    np.add.at(z, (rows, i1), w1)
    return z @ g
