"""Seeded synthetic inputs shared by the tests, the golden-fixture generator and bench.py.

Everything here is a *recipe*: the same call rebuilds the same data anywhere, so the
fixtures under tests/golden/ only need to hold the reference's OUTPUTS.
"""
from __future__ import annotations

import math

import numpy as np

CLASS_NAMES = ["benign", "malignant"]
FITZPATRICK = ["I", "II", "III", "IV", "V", "VI"]


def synthetic_instances(n: int, seed: int, with_oddities: bool = True) -> dict:
    """``predict_with_instance``-shaped dict[int -> instance dict] (tone_bias_dataset.py:389-392
    keys + 'prediction').  ``with_oddities`` sprinkles NaN / unknown sex values, which the
    reference's ``filter`` drops from both sex groups."""
    rng = np.random.default_rng(seed)
    label = rng.integers(0, 2, n)
    # a model that is right ~75 % of the time, less often on positives
    flip = rng.random(n) < np.where(label == 1, 0.45, 0.12)
    pred = np.where(flip, 1 - label, label)
    ftype = rng.choice(6, n, p=[0.35, 0.45, 0.10, 0.05, 0.03, 0.02])
    sex = rng.integers(0, 2, n)
    control = rng.integers(0, 2, n)
    odd = rng.random(n) < (0.02 if with_oddities else 0.0)
    keys = rng.permutation(4 * n)[:n]           # sparse, unordered dataframe indexes
    out = {}
    for i in range(n):
        st = FITZPATRICK[ftype[i]]
        if odd[i]:
            sx = float("nan") if i % 2 == 0 else "unknown"
        else:
            sx = "male" if sex[i] == 0 else "female"
        out[int(keys[i])] = {
            "file_path": f"/data/ISIC_{keys[i]:07d}.jpg", "image_name": f"ISIC_{keys[i]:07d}",
            "patient_id": f"IP_{int(keys[i]) % 997:04d}", "diagnosis": "nevus" if label[i] == 0 else "melanoma",
            "benign_malignant": CLASS_NAMES[label[i]], "age": float(20 + 5 * (int(keys[i]) % 13)),
            "sex": sx, "location": "torso", "skin_type": st,
            "skin_tone": "light" if st in ("I", "II") else "dark",
            "control": "rich" if control[i] == 0 else "poor",
            "prediction": CLASS_NAMES[pred[i]],
        }
    return out


def synthetic_u8_image(h: int, w: int, seed: int, kind: str = "noise") -> np.ndarray:
    """HWC uint8 'decode buffer'.  'noise' maximises anti-alias sensitivity; 'smooth' is the
    image-like variant of SURVEY section 8d (up-sampled low-frequency field + +-8 LSB noise)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "smooth":
        coarse = rng.random((15, 20, 3))
        yy = np.linspace(0, 14, h)
        xx = np.linspace(0, 19, w)
        y0 = np.floor(yy).astype(int).clip(0, 13)
        x0 = np.floor(xx).astype(int).clip(0, 18)
        ty = (yy - y0)[:, None, None]
        tx = (xx - x0)[None, :, None]
        f = (coarse[y0][:, x0] * (1 - ty) * (1 - tx) + coarse[y0 + 1][:, x0] * ty * (1 - tx)
             + coarse[y0][:, x0 + 1] * (1 - ty) * tx + coarse[y0 + 1][:, x0 + 1] * ty * tx)
        img = f * 255 + rng.integers(-8, 9, (h, w, 3))
        return np.ascontiguousarray(img.clip(0, 255).astype(np.uint8))
    if kind == "extremes":
        img = np.zeros((h, w, 3), np.uint8)
        img[::2, ::3] = 255
        img[0, :] = 255
        img[:, -1] = 255
        return img
    raise ValueError(kind)


def synthetic_batch_f32(batch: int, size: int, seed: int):
    """[B,3,S,S] float32 in [0,1): the model-stage input of BASELINE configs[0]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, size, size, generator=g, dtype=torch.float32)


def counter_metadata(index: np.ndarray, seed: int):
    """Counter-based per-logical-index metadata -- the recipe lives in the package so bench.py and
    the tests share it (skin_image_analysis_b200/synthetic.py)."""
    from skin_image_analysis_b200.synthetic import counter_metadata as _cm
    return _cm(index, seed)


def close(a, b, rel=1e-12):
    return math.isclose(a, b, rel_tol=rel, abs_tol=1e-15)
