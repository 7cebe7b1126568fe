#!/usr/bin/env python
"""Sharded evaluation over NCCL (run with torchrun on >= 2 GPUs): rank r evaluates the contiguous shard
shard_range(N, r, W) of N logical images (pixels from a small ring of synthetic batches, metadata unique per
logical index), the 576-byte count tensor is all-reduced once, and rank 0 checks the result bit for bit
against evaluating all N images itself.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from skin_image_analysis_b200 import distributed as D
from skin_image_analysis_b200.engine import EvalEngine
from bench import build_state
from skin_image_analysis_b200.synthetic import counter_metadata, device_u8_batches


def evaluate(eng, ring, lo, hi, batch):
    """Counts of logical images [lo, hi): image i uses pixels ring[(i // batch) % len(ring)][i % batch]."""
    eng.reset_counts()
    dev = eng.device
    for b0 in range((lo // batch) * batch, hi, batch):
        idx = np.arange(b0, b0 + batch, dtype=np.int64)
        label, ftype, sex, control = counter_metadata(idx, seed=9)
        groups = np.stack([ftype, sex, control])
        groups[:, (idx < lo) | (idx >= hi)] = 255            # images outside the shard: in no group
        eng.step(ring[(b0 // batch) % len(ring)], torch.from_numpy(label).to(dev), torch.from_numpy(groups).to(dev))
        eng.synchronize()
    return eng.counts.clone()


def main():
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    batch, n = 32, 1000
    state = build_state(dev)                     # random init, head centred so that both classes are predicted
    eng = EvalEngine(state, batch, (450, 600), 224, device=dev, n_slots=1)
    ring = device_u8_batches(3, batch, 450, 600, seed=5, device=dev)      # same seed: same pixels on every rank
    lo, hi = D.shard_range(n, rank, world)
    counts = evaluate(eng, ring, lo, hi, batch)
    with torch.cuda.stream(eng.stream):
        D.allreduce_counts(counts)
    eng.synchronize()
    torch.cuda.synchronize()
    if rank == 0:
        full = evaluate(eng, ring, 0, n, batch)
        ok = bool(torch.equal(full, counts))
        print(json.dumps({"world": world, "n": n, "bit_exact": ok, "counted": int(counts[0].sum()),
                          "malignant_predicted": int(counts[0, :, :, 1].sum())}), flush=True)
        if not ok:
            sys.exit(1)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
