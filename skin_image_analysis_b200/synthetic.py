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
