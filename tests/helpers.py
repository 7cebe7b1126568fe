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


def synthetic_metadata_df(n: int, seed: int):
    """ISIC-metadata-shaped dataframe with the columns ``HibaDataset.lookup_path`` reads
    (tone_bias_dataset.py:366-392); one NaN ``sex`` so the "falls in no group" rule is exercised."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    ftype = rng.choice(6, n, p=[0.3, 0.3, 0.15, 0.1, 0.1, 0.05])
    label = rng.integers(0, 2, n)
    sex = np.where(rng.integers(0, 2, n) == 0, "male", "female").astype(object)
    if n > 3:
        sex[3] = float("nan")
    types = [FITZPATRICK[t] for t in ftype]
    return pd.DataFrame({
        "isic_id": [f"ISIC_{9000000 + 7 * i:07d}" for i in range(n)],
        "patient_id": [f"IP_{1000 + i % 5}" for i in range(n)],
        "diagnosis": ["melanoma" if v else "nevus" for v in label],
        "benign_malignant": [CLASS_NAMES[v] for v in label],
        "age_approx": [float(30 + 5 * (i % 9)) for i in range(n)],
        "sex": sex,
        "anatom_site_general": ["torso" if i % 2 else "lower extremity" for i in range(n)],
        "fitzpatrick_skin_type": types,
        "skin_tone": ["light" if t in ("I", "II") else "dark" for t in types],
        "control": ["rich" if v == 0 else "poor" for v in rng.integers(0, 2, n)],
    })


def write_image_files(root_dir: str, df, h: int, w: int, seed: int, kind: str = "smooth"):
    """One LOSSLESS image file per row under the ``<isic_id>.jpg`` name ``HibaDataset.get_file_path`` builds
    (tone_bias_dataset.py:354-360).  The bytes are PNG (readers sniff the content, not the suffix), so the decode
    buffer is exactly the seeded array returned here -- JPEG would make it decoder-dependent."""
    from PIL import Image
    images = []
    for i, name in enumerate(df["isic_id"]):
        u8 = synthetic_u8_image(h, w, seed + i, kind)
        Image.fromarray(u8, "RGB").save(__import__("os").path.join(root_dir, name + ".jpg"), format="PNG")
        images.append(u8)
    return images


class FixedLogitModel:
    """Stand-in "model" for the evaluate_model / evaluate_model_by_class stdout fixtures: log-probabilities are a
    fixed seeded function of the per-image mean, so the reference functions and the drop-ins see the same outputs."""

    def __init__(self, seed: int):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.w = torch.randn(2, generator=g)

    def eval(self):
        return self

    def __call__(self, images):
        import torch
        m = images.float().mean(dim=(1, 2, 3)) - 0.5
        z = torch.stack([m * self.w[0], m * self.w[1] + 0.01], 1)
        return torch.log_softmax(z * 40.0, 1)


def fixed_eval_loader(n_batches: int, batch: int, seed: int):
    """List-of-batches stand-in for a DataLoader: (images [b,3,8,8], labels [b], indexes [b]); last batch ragged."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out, k = [], 0
    for b in range(n_batches):
        nb = batch if b < n_batches - 1 else max(1, batch - 3)
        out.append((torch.rand(nb, 3, 8, 8, generator=g), torch.randint(0, 2, (nb,), generator=g),
                    torch.arange(k, k + nb)))
        k += nb
    return out
