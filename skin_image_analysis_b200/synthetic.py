"""Seeded synthetic workloads for the benchmark and the tests (there is no dataset offline).

* ``counter_metadata``  label / Fitzpatrick type / sex / control for a LOGICAL image index from a
  counter-based hash of (seed, index): identical on every rank and for any partition of the index space,
  so sharded and single-GPU runs can be compared bit for bit.
* ``device_u8_batches``  distinct ISIC-shaped uint8 batches resident in HBM.
"""
from __future__ import annotations

import numpy as np
import torch

FITZPATRICK_CDF = np.cumsum([0.35, 0.45, 0.10, 0.05, 0.03, 0.02])   # ~80 % light, like the notebook's test set


def counter_metadata(index: np.ndarray, seed: int):
    """-> (label, fitzpatrick_type, sex, control) uint8 arrays (splitmix64 of seed and index)."""
    with np.errstate(over="ignore"):
        x = index.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(seed + 1)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    label = (x & np.uint64(1)).astype(np.uint8)
    u = ((x >> np.uint64(8)) & np.uint64(0xFFFF)).astype(np.float64) / 65536.0
    ftype = np.searchsorted(FITZPATRICK_CDF, u, side="right").clip(0, 5).astype(np.uint8)
    sex = ((x >> np.uint64(32)) & np.uint64(1)).astype(np.uint8)
    control = ((x >> np.uint64(40)) & np.uint64(1)).astype(np.uint8)
    return label, ftype, sex, control


def device_u8_batches(n_batches: int, batch: int, h: int, w: int, seed: int, device) -> list[torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    return [torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device=device, generator=g)
            for _ in range(n_batches)]


ARCHS = {   # reference state_dict prefixes (tone_bias_model.py:56-152 and :155-299) and layer widths
    "SkinCancerListModel": dict(convs=[("layers.0", 32, 3, 7), ("layers.3", 64, 32, 3), ("layers.6", 128, 64, 3)],
                                linears=[("layers.10", 512), ("layers.13", 256), ("layers.16", 2)]),
    "SkinCancerModel": dict(convs=[("conv1", 32, 3, 7), ("conv2", 64, 32, 3), ("conv3", 128, 64, 3),
                                   ("conv4", 256, 128, 3)],
                            linears=[("fc4", 512), ("fc5", 256), ("fc6", 2)]),
    # tone_bias_optuna.create_best_model (:96-120): nn.Sequential indices of the Conv2d / Linear children
    "optuna_best": dict(convs=[("0", 192, 3, 7), ("3", 172, 192, 3), ("6", 22, 172, 3), ("9", 86, 22, 3)],
                        linears=[("13", 227), ("16", 80), ("19", 86), ("22", 2)]),
}


def random_state_dict(kind: str, image_size: int = 224, seed: int = 0) -> dict:
    """Random-init weights of a reference architecture with the reference's init (xavier_normal_ weights,
    tone_bias_model.py:136-137; torch-default uniform biases) for any input size: the reference hard-codes 224
    (:69-70), for other sizes only the first Linear's in_features changes."""
    import math
    g = torch.Generator().manual_seed(seed)
    out, side, ch = {}, image_size, 3
    for prefix, cout, cin, k in ARCHS[kind]["convs"]:
        fan_in, fan_out = cin * k * k, cout * k * k
        out[prefix + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * math.sqrt(2.0 / (fan_in + fan_out))
        out[prefix + ".bias"] = (torch.rand((cout,), generator=g) * 2 - 1) / math.sqrt(fan_in)
        side //= 2
        ch = cout
    feat = ch * side * side
    for prefix, width in ARCHS[kind]["linears"]:
        out[prefix + ".weight"] = torch.randn((width, feat), generator=g) * math.sqrt(2.0 / (feat + width))
        out[prefix + ".bias"] = (torch.rand((width,), generator=g) * 2 - 1) / math.sqrt(feat)
        feat = width
    return out
