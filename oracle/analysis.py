"""Oracle: the per-group confusion reduction and fairness metrics, restated in pure
Python over instance dicts (TEST INFRASTRUCTURE ONLY).

Follows tone_bias_test.py:
  * ``confusion_matrix``            :240-272  ('malignant' is the positive class; ValueError when
                                              an instance falls in none of the four cells)
  * ``values_counts`` / ``filter``  :274-289  (== on the feature value; NaN or any other value
                                              lands in no group)
  * ``disparate_impact_analysis``   :292-445  (27-key result; 0.0 guards at :309, :339, :348, :361;
                                              unguarded divisions by group size at :327-333)
  * ``analyse_predictions``         :448-561  (overall accuracy; dark/light, female/male, poor/rich;
                                              the stdout lines are part of the behaviour)
and the count -> selection-rate flattening of tone_bias_analysis.py:357-367.

Known-answer vectors: the reference's saved notebook run
(notebooks/jgi_hiba_2022_torch.ipynb raw 3591-3617 and 3643-3669), kept in
tests/golden/notebook_di.json and checked by tests/test_oracle_analysis.py.
"""
from __future__ import annotations

POSITIVE = "malignant"
NEGATIVE = "benign"

DI_KEYS = (
    "accuracy", "precision", "recall", "f1", "selection_rate_min", "selection_rate_maj", "di",
    "min_prevalence", "maj_prevalence", "min_selected", "min_count", "maj_selected", "maj_count",
    "min_precision", "min_recall", "min_f1", "maj_precision", "maj_recall", "maj_f1",
    "tp_min", "tn_min", "fp_min", "fn_min", "tp_maj", "tn_maj", "fp_maj", "fn_maj",
)


def confusion_matrix(instances: dict):
    cells = {"tp": {}, "tn": {}, "fp": {}, "fn": {}}
    for index in sorted(instances):
        inst = instances[index]
        pred, label = inst["prediction"], inst["benign_malignant"]
        if pred == POSITIVE and label == POSITIVE:
            cells["tp"][index] = inst
        elif pred == NEGATIVE and label == NEGATIVE:
            cells["tn"][index] = inst
        elif pred == POSITIVE and label == NEGATIVE:
            cells["fp"][index] = inst
        elif pred == NEGATIVE and label == POSITIVE:
            cells["fn"][index] = inst
    placed = sum(len(c) for c in cells.values())
    if placed != len(instances):
        raise ValueError(
            f"tp={len(cells['tp'])} + tn={len(cells['tn'])} + fp={len(cells['fp'])} + "
            f"fn={len(cells['fn'])} != {len(instances)}")
    return cells["tp"], cells["tn"], cells["fp"], cells["fn"]


def values_counts(instances: dict, feature: str, value) -> int:
    return sum(1 for inst in instances.values() if inst[feature] == value)


def filter(instances: dict, feature: str, value) -> dict:  # noqa: A001 - reference name
    return {k: inst for k, inst in instances.items() if inst[feature] == value}


def _prf(tp: int, fp: int, fn: int):
    """precision / recall / f1 with the reference's tp == 0 guard."""
    if tp <= 0:
        return 0.0, 0.0, 0.0
    p = tp / (tp + fp)
    r = tp / (tp + fn)
    return p, r, 2 * ((p * r) / (p + r))


def di_from_cells(tp_min, tn_min, fp_min, fn_min, tp_maj, tn_maj, fp_maj, fn_maj) -> dict:
    """The arithmetic of ``disparate_impact_analysis`` on the eight cell sizes."""
    tp, tn, fp, fn = tp_min + tp_maj, tn_min + tn_maj, fp_min + fp_maj, fn_min + fn_maj
    accuracy = (tp + tn) / (tp + tn + fp + fn)
    precision, recall, f1 = _prf(tp, fp, fn)
    min_count = tp_min + tn_min + fp_min + fn_min
    maj_count = tp_maj + tn_maj + fp_maj + fn_maj
    min_selected, maj_selected = tp_min + fp_min, tp_maj + fp_maj
    sr_min = min_selected / min_count
    sr_maj = maj_selected / maj_count
    min_prev = (tp_min + fn_min) / min_count
    maj_prev = (tp_maj + fn_maj) / maj_count
    min_p, min_r, min_f1 = _prf(tp_min, fp_min, fn_min)
    maj_p, maj_r, maj_f1 = _prf(tp_maj, fp_maj, fn_maj)
    di = sr_min / sr_maj if sr_maj > 0.0 else 0.0
    values = (
        accuracy, precision, recall, f1, sr_min, sr_maj, di, min_prev, maj_prev,
        min_selected, min_count, maj_selected, maj_count,
        min_p, min_r, min_f1, maj_p, maj_r, maj_f1,
        tp_min, tn_min, fp_min, fn_min, tp_maj, tn_maj, fp_maj, fn_maj,
    )
    return dict(zip(DI_KEYS, values))


def disparate_impact_analysis(min_instances: dict, maj_instances: dict) -> dict:
    cm_min = [len(c) for c in confusion_matrix(min_instances)]
    cm_maj = [len(c) for c in confusion_matrix(maj_instances)]
    return di_from_cells(*cm_min, *cm_maj)


def analyse_predictions(instances: dict, out=print) -> dict:
    correct = sum(1 for i in instances.values() if i["prediction"] == i["benign_malignant"])
    total = len(instances)
    out(f"Total={total} correct={correct} my accuracy={correct / total:.3f}")

    dark, light = filter(instances, "skin_tone", "dark"), filter(instances, "skin_tone", "light")
    out(f"dark {len(dark)}")
    out(f"light {len(light)}")
    male, female = filter(instances, "sex", "male"), filter(instances, "sex", "female")
    out(f"male {len(male)}")
    out(f"female {len(female)}")
    out(f"total {len(instances)}")
    rich, poor = filter(instances, "control", "rich"), filter(instances, "control", "poor")
    out(f"rich {len(rich)}")
    out(f"poor {len(poor)}")

    tp, _tn, _fp, _fn = confusion_matrix(instances)
    m, f, g = values_counts(tp, "sex", "male"), values_counts(tp, "sex", "female"), len(tp)
    out(f"TP: male_count={m} female_count={f}")
    if g > 0:
        out(f"TP: P(   male | mole=malignant ) = {m / g}")
        out(f"TP: P( female | mole=malignant ) = {f / g}")
    out(f"TP: male + female = {m + f}  total = {g}")

    m, f, g = len(male), len(female), len(instances)
    out()
    out(f"TEST_SET: male_count={m} female_count={f}")
    out(f"TEST_SET: P(   male ) = {m / g:.3f}")
    out(f"TEST_SET: P( female ) = {f / g:.3f}")
    out(f"TEST_SET: male + female = {m + f}  total = {g}")

    lc, dc = len(light), len(dark)
    out()
    out(f"TEST_SET: light_count={lc} dark_count={dc}")
    if g > 0:
        out(f"TEST_SET: P( light ) = {lc / g:.3f}")
        out(f"TEST_SET: P(  dark ) = {dc / g:.3f}")
    out(f"TEST_SET: light + dark = {lc + dc}  total = {g}")

    dpos = values_counts(dark, "benign_malignant", POSITIVE)
    lpos = values_counts(light, "benign_malignant", POSITIVE)
    dprev, lprev = dpos / len(dark), lpos / len(light)      # unguarded, as in the reference
    out(f"Dark Prevalence: {dpos} / {len(dark)} = {dprev:.2f}")
    out(f"Light Prevalence: {lpos} / {len(light)} = {lprev:.2f}")

    out("DISPARATE IMPACT: SKIN TONE")
    tone = disparate_impact_analysis(dark, light)
    out("DISPARATE IMPACT: GENDER")
    gender = disparate_impact_analysis(female, male)
    out("DISPARATE IMPACT: CONTROL")
    control = disparate_impact_analysis(poor, rich)

    return {
        "correct": correct, "total": total, "accuracy": correct / total,
        "dark": len(dark), "light": len(light), "male": len(male), "female": len(female),
        "tone_di_results": tone, "gender_di_results": gender, "control_di_results": control,
    }


def counts_table(instances: dict, attributes: dict[str, list]) -> dict[str, list]:
    """[group][label][pred] integer tables per attribute -- the layout the CUDA count kernel
    produces.  ``attributes`` maps feature name -> ordered list of group values; label / pred
    index 1 = 'malignant'.  Instances whose feature value is in no group are dropped (filter
    semantics, tone_bias_test.py:283-289)."""
    out = {}
    for feat, values in attributes.items():
        tab = [[[0, 0], [0, 0]] for _ in values]
        for inst in instances.values():
            for g, v in enumerate(values):
                if inst[feat] == v:
                    lab = 1 if inst["benign_malignant"] == POSITIVE else 0
                    prd = 1 if inst["prediction"] == POSITIVE else 0
                    tab[g][lab][prd] += 1
        out[feat] = tab
    return out


def selection_rates_from_result(result: dict) -> tuple[float, float]:
    """tone_bias_analysis.py:357-367 -- recompute tone selection rates from the counts."""
    t = result["tone_di_results"]
    sr_min = (t["tp_min"] + t["fp_min"]) / t["min_count"]
    sr_maj = (t["tp_maj"] + t["fp_maj"]) / t["maj_count"]
    return sr_min, sr_maj
