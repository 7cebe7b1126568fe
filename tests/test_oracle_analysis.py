"""Oracle (oracle/analysis.py) vs the reference's golden vectors -- CPU only."""
import contextlib
import io
import json
import math
import os

import pytest

from oracle import analysis as oa
from tests import helpers


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert set(a) == set(b), path
        for k in a:
            _same(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, float) or isinstance(b, float):
        assert math.isclose(a, b, rel_tol=1e-15, abs_tol=0.0), (path, a, b)
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("attr", ["tone", "sex"])
def test_notebook_known_answers(golden_dir, attr):
    """The reference's saved notebook run (raw 3591-3617 / 3643-3669) through the oracle arithmetic."""
    nb = _load(golden_dir, "notebook_di.json")[attr]
    c = nb["cells"]
    r = oa.di_from_cells(c["tp_min"], c["tn_min"], c["fp_min"], c["fn_min"],
                         c["tp_maj"], c["tn_maj"], c["fp_maj"], c["fn_maj"])
    p = nb["printed"]
    for key in ("min_precision", "min_recall", "min_f1", "maj_precision", "maj_recall", "maj_f1", "f1",
                "selection_rate_min", "selection_rate_maj", "di", "min_prevalence", "maj_prevalence"):
        assert f"{r[key]:.3f}" == f"{p[key]:.3f}", key
    for key in ("min_selected", "min_count", "maj_selected", "maj_count"):
        assert r[key] == p[key]
    assert f"{r['selection_rate_maj'] / r['selection_rate_min']:.3f}" == f"{p['di_inverse']:.3f}"
    assert f"{(c['tp_min'] + c['tn_min']) / r['min_count']:.3f}" == f"{p['min_group_accuracy']:.3f}"
    assert f"{(c['tp_maj'] + c['tn_maj']) / r['maj_count']:.3f}" == f"{p['maj_group_accuracy']:.3f}"


def test_notebook_sizes_consistent(golden_dir):
    nb = _load(golden_dir, "notebook_di.json")
    t, s = nb["tone"]["cells"], nb["sex"]["cells"]
    assert t["tp_min"] + t["tn_min"] + t["fp_min"] + t["fn_min"] == nb["sizes"]["dark"]
    assert t["tp_maj"] + t["tn_maj"] + t["fp_maj"] + t["fn_maj"] == nb["sizes"]["light"]
    assert s["tp_min"] + s["tn_min"] + s["fp_min"] + s["fn_min"] == nb["sizes"]["female"]
    assert s["tp_maj"] + s["tn_maj"] + s["fp_maj"] + s["fn_maj"] == nb["sizes"]["male"]
    # one instance has a sex that is neither 'male' nor 'female' -> in no sex group (filter semantics)
    assert nb["sizes"]["male"] + nb["sizes"]["female"] == nb["sizes"]["total"] - 1
    correct = t["tp_min"] + t["tn_min"] + t["tp_maj"] + t["tn_maj"]
    assert correct == nb["overall"]["correct"]
    assert t["tp_min"] + t["fn_min"] == nb["prevalence"]["dark_pos"]
    assert t["tp_maj"] + t["fn_maj"] == nb["prevalence"]["light_pos"]


@pytest.mark.parametrize("case", ["n500", "n64", "n1087"])
def test_reference_generated_fixture(golden_dir, case):
    g = _load(golden_dir, "analysis_synth.json")[case]
    inst = helpers.synthetic_instances(g["n"], g["seed"], g["with_oddities"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = oa.analyse_predictions(inst)
    _same(json.loads(json.dumps(res)), g["result"])
    assert buf.getvalue() == g["stdout"]
    assert [len(c) for c in oa.confusion_matrix(inst)] == g["cells"]
    assert oa.values_counts(inst, "skin_tone", "dark") == g["dark_count"]


def test_confusion_matrix_rejects_unknown_label():
    inst = helpers.synthetic_instances(10, 3, False)
    k = next(iter(inst))
    inst[k]["prediction"] = "indeterminate"
    with pytest.raises(ValueError):
        oa.confusion_matrix(inst)


def test_empty_group_divides_by_zero_like_reference():
    inst = helpers.synthetic_instances(40, 5, False)
    light_only = {k: v for k, v in inst.items() if v["skin_tone"] == "light"}
    with pytest.raises(ZeroDivisionError):
        oa.analyse_predictions(light_only, out=lambda *a: None)


def test_counts_table_matches_filters():
    inst = helpers.synthetic_instances(300, 9, True)
    tab = oa.counts_table(inst, {"skin_tone": ["light", "dark"], "sex": ["male", "female"],
                                 "control": ["rich", "poor"], "skin_type": helpers.FITZPATRICK})
    for feat, values in [("skin_tone", ["light", "dark"]), ("sex", ["male", "female"])]:
        for g, v in enumerate(values):
            sub = oa.filter(inst, feat, v)
            tp, tn, fp, fn = (len(c) for c in oa.confusion_matrix(sub))
            assert tab[feat][g] == [[tn, fp], [fn, tp]]
    assert sum(sum(map(sum, t)) for t in tab["skin_type"]) == len(inst)
    assert sum(sum(map(sum, t)) for t in tab["sex"]) < len(inst)      # the NaN / unknown rows
