"""World-size-2 checks of the sharding / all-reduce host logic on the gloo backend (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_counts(pred, label, groups, n_groups=6):
    c = np.zeros((groups.shape[0], n_groups, 2, 2), np.int64)
    for a in range(groups.shape[0]):
        ok = groups[a] < n_groups
        np.add.at(c, (a, groups[a][ok], label[ok], pred[ok]), 1)
    return c


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from skin_image_analysis_b200 import distributed as D
    r, w, _ = D.init_from_env(backend="gloo")
    lo, hi = D.shard_range(n, r, w)
    idx = np.arange(lo, hi)
    label, ftype, sex, control = helpers.counter_metadata(idx, seed=3)
    pred = (helpers.counter_metadata(idx, seed=4)[0] ^ label) & 1
    counts = torch.from_numpy(_np_counts(pred, label, np.stack([ftype, sex, control])))
    D.allreduce_counts(counts)
    t = D.max_over_ranks(float(rank + 1), device="cpu")
    q.put((rank, counts.numpy().tolist(), t, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 2])
def test_allreduced_counts_equal_single_process(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(n)
    label, ftype, sex, control = helpers.counter_metadata(idx, seed=3)
    pred = (helpers.counter_metadata(idx, seed=4)[0] ^ label) & 1
    full = _np_counts(pred, label, np.stack([ftype, sex, control]))
    for rank, counts, t, _rng in res:
        assert counts == full.tolist()
        assert t == 2.0
    ranges = sorted(r[3] for r in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == n and ranges[0][1] == ranges[1][0]


def test_shard_range_partitions():
    from skin_image_analysis_b200.distributed import shard_range
    for n in (0, 1, 7, 1_000_000):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_shard_batches_cover_every_image_once():
    """configs[2] sharding (bench.py --workload shard1m): whole batches, contiguous, balanced, ragged tail only in the
    last global batch."""
    from skin_image_analysis_b200.distributed import shard_batches
    for n_images, batch in [(1_000_000, 256), (1000, 32), (5, 8), (512, 256)]:
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi, nb = shard_batches(n_images, batch, r, world)
                assert nb == -(-n_images // batch) and 0 <= lo <= hi <= nb
                seen += list(range(lo, hi))
            assert seen == list(range(nb))
            sizes = [shard_batches(n_images, batch, r, world)[1] - shard_batches(n_images, batch, r, world)[0]
                     for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
