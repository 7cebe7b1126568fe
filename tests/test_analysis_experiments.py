"""tone_bias_analysis drop-in (host aggregation) vs outputs of the reference's own
``tone_bias_analysis`` on the same synthetic result files (tests/golden/analysis_experiments.json, produced by
tests/golden/make_golden.py) -- CPU only."""
import json
import math
import os

import pytest

from skin_image_analysis_b200 import tone_bias_analysis as ta


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert set(map(str, a)) == set(map(str, b)), path
        for k in a:
            _same(a[k], b[str(k)] if str(k) in b else b[k], path + "/" + str(k))
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif isinstance(a, float) or isinstance(b, float):
        assert (math.isnan(a) and math.isnan(b)) or math.isclose(a, b, rel_tol=1e-15, abs_tol=0.0), (path, a, b)
    else:
        assert a == b, (path, a, b)


@pytest.fixture()
def experiments(golden_dir, tmp_path):
    with open(os.path.join(golden_dir, "analysis_experiments.json")) as f:
        g = json.load(f)
    for name, files in g["folders"].items():
        (tmp_path / name).mkdir()
        for fname, lines in files.items():
            (tmp_path / name / fname).write_text("\n".join(lines) + "\n")
    return g, tmp_path


def test_read_experiment_flattens_and_renumbers(experiments):
    g, root = experiments
    for name, want in g["read_experiment"].items():
        got = ta.read_experiment(str(root / name))
        assert list(got.keys()) == list(range(1, len(want) + 1))          # global epochs across files
        _same(got, want)


def test_read_experiments_means_and_transpose(experiments):
    g, root = experiments
    lines = []
    for prefix, want in g["read_experiments"].items():
        got = ta.read_experiments(str(root), prefix, 2, out=lines.append)
        _same(got, want)
        _same(ta.transpose_dict(got), g["transpose"][prefix])
        assert ta.get_measure(got, "accuracy") == g["transpose"][prefix]["accuracy"]
    assert any(s.startswith("FILE ") for s in lines) and any("AGGREGATE accuracy" in s for s in lines)


def test_compute_ci_matches_reference():
    here = os.path.join(os.path.dirname(__file__), "golden", "analysis_experiments.json")
    with open(here) as f:
        cases = json.load(f)["compute_ci"]
    assert {len(c["data"]) for c in cases} >= {30, 31}                    # both sides of the t / normal switch
    for c in cases:
        lo, hi = ta.compute_ci(c["data"], c["level"])
        assert math.isclose(lo, c["low"], rel_tol=1e-13) and math.isclose(hi, c["high"], rel_tol=1e-13)


def test_epoch_sanity_check_and_error_types(tmp_path):
    rec = {"epoch": 3, "tone_di_results": {"tp_min": 1, "fp_min": 1, "min_count": 4, "tp_maj": 1, "fp_maj": 0,
                                           "maj_count": 2, "di": 1.0, "f1": 0.5},
           "gender_di_results": {"di": 1.0}, "control_di_results": {"di": 1.0}}
    d = tmp_path / "exp"
    d.mkdir()
    (d / "a.json").write_text(json.dumps(rec) + "\n")
    with pytest.raises(ValueError):                                      # first record claims epoch 3
        ta.read_experiment(str(d))
    rec["epoch"] = 1
    rec["tone_di_results"]["min_count"] = 0
    (d / "a.json").write_text(json.dumps(rec) + "\n")
    with pytest.raises(ZeroDivisionError):                               # unguarded, as in the reference
        ta.read_experiment(str(d))


def test_writer_round_trip(tmp_path):
    """append_results_line writes what read_experiment reads (tone_bias_train.py:410-424 format)."""
    import numpy as np
    from skin_image_analysis_b200 import tone_bias_test as tt
    from tests import helpers

    def host_counts(instances):                       # numpy stand-in for the count kernel (no GPU in this suite)
        _keys, pred, label, groups = tt.encode_instances(instances)
        c = np.zeros((groups.shape[0], tt.N_GROUPS, 2, 2), np.int64)
        for a in range(groups.shape[0]):
            ok = groups[a] < tt.N_GROUPS
            np.add.at(c[a], (groups[a][ok], label[ok], pred[ok]), 1)
        return c

    d = tmp_path / "balanced_x"
    d.mkdir()
    path = str(d / "2024-01-01_00-00-00.json")
    recs = []
    for epoch in (1, 2, 3):
        res = tt.results_from_counts(host_counts(helpers.synthetic_instances(200, epoch, False)), out=lambda *a: None)
        recs.append(ta.append_results_line(path, res, 0.5 / epoch, 0.6 + 0.1 * epoch, epoch))
    got = ta.read_experiment(str(d))
    assert list(got) == [1, 2, 3]
    for e, r in zip((1, 2, 3), recs):
        assert got[e]["accuracy"] == r["accuracy"] and got[e]["tone_di"] == r["tone_di_results"]["di"]
        assert got[e]["avg_batch_loss"] == 0.5 / e
