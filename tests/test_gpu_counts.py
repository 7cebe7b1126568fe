"""K7 parity: sia_confusion_counts vs the oracle's dict-based analysis -- bit-exact."""
import contextlib
import io
import json
import os

import numpy as np
import pytest
import torch

from oracle import analysis as oa
from tests import helpers

pytestmark = pytest.mark.gpu


def _np_counts(pred, label, groups, n_groups):
    a = groups.shape[0]
    c = np.zeros((a, n_groups, 2, 2), np.int64)
    for ai in range(a):
        ok = groups[ai] < n_groups
        np.add.at(c, (ai, groups[ai][ok], label[ok], pred[ok]), 1)
    return c


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 65537, 1_000_003])
def test_counts_match_numpy(n):
    from skin_image_analysis_b200 import ops
    rng = np.random.default_rng(n)
    pred = rng.integers(0, 2, n).astype(np.uint8)
    label = rng.integers(0, 2, n).astype(np.uint8)
    groups = np.stack([rng.integers(0, 6, n), rng.choice([0, 1, 255], n, p=[.45, .5, .05]),
                       rng.integers(0, 2, n)]).astype(np.uint8)
    got = ops.confusion_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(label).cuda(),
                               torch.from_numpy(groups).cuda(), 6).cpu().numpy()
    assert np.array_equal(got, _np_counts(pred, label, groups, 6))
    assert got[0].sum() == n


def test_counts_accumulate_and_empty():
    from skin_image_analysis_b200 import ops
    z = torch.zeros(0, dtype=torch.uint8, device="cuda")
    c = ops.confusion_counts(z, z, torch.zeros((3, 0), dtype=torch.uint8, device="cuda"), 6)
    assert int(c.sum()) == 0
    p = torch.ones(100, dtype=torch.uint8, device="cuda")
    g = torch.zeros((1, 100), dtype=torch.uint8, device="cuda")
    ops.confusion_counts(p, p, g, 2, counts=c[:1, :2].contiguous())
    c2 = torch.zeros((1, 2, 2, 2), dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.confusion_counts(p, p, g, 2, counts=c2)
    assert c2[0, 0, 1, 1].item() == 300 and int(c2.sum()) == 300


def test_all_same_bin_contention():
    from skin_image_analysis_b200 import ops
    n = 3_000_000
    p = torch.zeros(n, dtype=torch.uint8, device="cuda")
    g = torch.full((2, n), 5, dtype=torch.uint8, device="cuda")
    c = ops.confusion_counts(p, p, g, 6)
    assert c[0, 5, 0, 0].item() == n and c[1, 5, 0, 0].item() == n and int(c.sum()) == 2 * n


@pytest.mark.parametrize("case", ["n500", "n64", "n1087"])
def test_analyse_predictions_dropin_matches_reference_fixture(golden_dir, case):
    """The product's analyse_predictions (CUDA counts + host arithmetic) vs the reference's own output."""
    from skin_image_analysis_b200 import tone_bias_test as tt
    with open(os.path.join(golden_dir, "analysis_synth.json")) as f:
        g = json.load(f)[case]
    inst = helpers.synthetic_instances(g["n"], g["seed"], g["with_oddities"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = tt.analyse_predictions(inst)
    assert json.loads(json.dumps(res)) == g["result"]
    assert buf.getvalue() == g["stdout"]
    assert [len(c) for c in tt.confusion_matrix(inst)] == g["cells"]
    assert tt.values_counts(inst, "skin_tone", "dark") == g["dark_count"]
    dark, light = tt.filter(inst, "skin_tone", "dark"), tt.filter(inst, "skin_tone", "light")
    assert tt.disparate_impact_analysis(dark, light) == oa.disparate_impact_analysis(dark, light)


def test_dropin_error_behaviour():
    from skin_image_analysis_b200 import tone_bias_test as tt
    inst = helpers.synthetic_instances(40, 5, False)
    light_only = {k: v for k, v in inst.items() if v["skin_tone"] == "light"}
    with pytest.raises(ZeroDivisionError), contextlib.redirect_stdout(io.StringIO()):
        tt.analyse_predictions(light_only)
    k = next(iter(inst))
    inst[k]["prediction"] = "indeterminate"
    with pytest.raises(ValueError):
        tt.confusion_matrix(inst)


def test_sharded_counts_equal_single_pass():
    """Any partition of the index space gives the same summed counts (what the all-reduce relies on)."""
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200.distributed import shard_range
    n = 200_000
    idx = np.arange(n)
    label, ftype, sex, control = helpers.counter_metadata(idx, seed=3)
    pred = (helpers.counter_metadata(idx, seed=4)[0] ^ label) & 1
    groups = np.stack([ftype, sex, control])
    full = ops.confusion_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(label).cuda(),
                                torch.from_numpy(groups).cuda(), 6)
    for world in (2, 4, 8, 7):
        acc = torch.zeros_like(full)
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            ops.confusion_counts(torch.from_numpy(pred[lo:hi]).cuda(), torch.from_numpy(label[lo:hi]).cuda(),
                                 torch.from_numpy(np.ascontiguousarray(groups[:, lo:hi])).cuda(), 6, counts=acc)
        assert torch.equal(acc, full)
