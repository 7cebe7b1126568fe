"""Drop-in for ``tone_bias_test`` (reference src/tone_bias_test.py) -- the evaluation reduction.

Same function names, arguments, return schemas, printed lines and exception types as the
reference; the counting itself -- the reference's per-instance Python loops (:207-234, :240-289)
-- is one launch of the ``sia_confusion_counts`` kernel (csrc/counts.cu) over byte-encoded
(prediction, label, group) arrays.  Everything after the counts is a few dozen Python divisions
(:292-445, :448-561) and stays on the host in float64 so the metrics are bit-identical.

New, faster entry points with no reference analogue:
  * ``encode_instances``        instance dicts -> uint8 arrays
  * ``counts_from_arrays``      CUDA arrays -> counts[A][G][2][2] int64
  * ``results_from_counts``     counts -> the reference's ``analyse_predictions`` result dict
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import SiaError

POSITIVE, NEGATIVE = "malignant", "benign"
FITZPATRICK = ("I", "II", "III", "IV", "V", "VI")

# attribute order of the counts tensor used by analyse_predictions
ATTRIBUTES = (
    ("skin_tone", ("light", "dark")),
    ("sex", ("male", "female")),
    ("control", ("rich", "poor")),
    ("skin_type", FITZPATRICK),
    (None, (None,)),                      # every instance, group 0: overall totals
)
N_GROUPS = 6
NO_GROUP = 255

DI_KEYS = (
    "accuracy", "precision", "recall", "f1", "selection_rate_min", "selection_rate_maj", "di",
    "min_prevalence", "maj_prevalence", "min_selected", "min_count", "maj_selected", "maj_count",
    "min_precision", "min_recall", "min_f1", "maj_precision", "maj_recall", "maj_f1",
    "tp_min", "tn_min", "fp_min", "fn_min", "tp_maj", "tn_maj", "fp_maj", "fn_maj",
)


# -------------------------------------------------------------------------------------------------
# encoding + the CUDA reduction
# -------------------------------------------------------------------------------------------------
def _class_code(value) -> int:
    return 1 if value == POSITIVE else 0 if value == NEGATIVE else 2


def encode_instances(instances: dict, attributes=ATTRIBUTES):
    """-> (keys, pred u8[N], label u8[N], groups u8[A,N]); class code 2 = neither class name."""
    keys = sorted(instances.keys())
    n = len(keys)
    pred = np.empty(n, np.uint8)
    label = np.empty(n, np.uint8)
    groups = np.full((len(attributes), n), NO_GROUP, np.uint8)
    for i, k in enumerate(keys):
        inst = instances[k]
        pred[i] = _class_code(inst["prediction"])
        label[i] = _class_code(inst["benign_malignant"])
        for a, (feature, values) in enumerate(attributes):
            if feature is None:
                groups[a, i] = 0
                continue
            v = inst[feature]
            for g, gv in enumerate(values):
                if v == gv:                     # == like the reference's filter(): NaN matches nothing
                    groups[a, i] = g
                    break
    return keys, pred, label, groups


def counts_from_arrays(pred: torch.Tensor, label: torch.Tensor, groups: torch.Tensor, n_groups: int = N_GROUPS,
                       counts: torch.Tensor | None = None) -> torch.Tensor:
    """CUDA uint8 arrays -> int64 counts[A][G][label][pred] on the device (accumulates into ``counts``)."""
    return ops.confusion_counts(pred, label, groups, n_groups, counts)


def _device():
    if not torch.cuda.is_available():
        raise SiaError("no CUDA device: the confusion-count reduction has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _count_instances(instances: dict, attributes=ATTRIBUTES, n_groups: int = N_GROUPS) -> np.ndarray:
    """instances -> host int64 counts[A][G][2][2] via the kernel; ValueError like the reference when an
    instance is in none of the four confusion cells (tone_bias_test.py:268-271)."""
    _keys, pred, label, groups = encode_instances(instances, attributes)
    dev = _device()
    bad = (pred > 1) | (label > 1)
    if bad.any():
        groups = groups.copy()
        groups[:, bad] = NO_GROUP
        pred, label = np.where(bad, 0, pred).astype(np.uint8), np.where(bad, 0, label).astype(np.uint8)
    if len(_keys) == 0:
        return np.zeros((len(attributes), n_groups, 2, 2), np.int64)
    c = counts_from_arrays(torch.from_numpy(pred).to(dev), torch.from_numpy(label).to(dev),
                           torch.from_numpy(groups).to(dev), n_groups).cpu().numpy()
    if bad.any():
        total = c[-1, 0] if attributes[-1][0] is None else None
        if total is not None:
            tn, fp, fn, tp = int(total[0, 0]), int(total[0, 1]), int(total[1, 0]), int(total[1, 1])
            raise ValueError(f"tp={tp} + tn={tn} + fp={fp} + fn={fn} != {len(instances)}")
        raise ValueError("instances outside the benign/malignant confusion cells")
    return c


def _cells(table) -> tuple[int, int, int, int]:
    """[label][pred] table -> (tp, tn, fp, fn) with 'malignant' positive (tone_bias_test.py:253-267)."""
    return int(table[1][1]), int(table[0][0]), int(table[0][1]), int(table[1][0])


# -------------------------------------------------------------------------------------------------
# host arithmetic on the counts (float64, same operation order as the reference)
# -------------------------------------------------------------------------------------------------
def _prf(tp, fp, fn):
    if tp > 0:
        precision = tp / (tp + fp)
        recall = tp / (tp + fn)
        return precision, recall, 2 * ((precision * recall) / (precision + recall))
    return 0.0, 0.0, 0.0


def di_from_tables(min_table, maj_table) -> dict:
    """``disparate_impact_analysis`` (tone_bias_test.py:292-445) on two [label][pred] count tables."""
    tp_min, tn_min, fp_min, fn_min = _cells(min_table)
    tp_maj, tn_maj, fp_maj, fn_maj = _cells(maj_table)
    tp, tn, fp, fn = tp_min + tp_maj, tn_min + tn_maj, fp_min + fp_maj, fn_min + fn_maj
    accuracy = (tp + tn) / (tp + tn + fp + fn)
    precision, recall, f1 = _prf(tp, fp, fn)
    min_count = tp_min + tn_min + fp_min + fn_min
    maj_count = tp_maj + tn_maj + fp_maj + fn_maj
    min_selected, maj_selected = tp_min + fp_min, tp_maj + fp_maj
    selection_rate_min = min_selected / min_count          # ZeroDivisionError on an empty group,
    selection_rate_maj = maj_selected / maj_count          # exactly like the reference (:327-328)
    min_prevalence = (tp_min + fn_min) / min_count
    maj_prevalence = (tp_maj + fn_maj) / maj_count
    min_precision, min_recall, min_f1 = _prf(tp_min, fp_min, fn_min)
    maj_precision, maj_recall, maj_f1 = _prf(tp_maj, fp_maj, fn_maj)
    di = 0.0
    if selection_rate_maj > 0.0:
        di = selection_rate_min / selection_rate_maj
    vals = (accuracy, precision, recall, f1, selection_rate_min, selection_rate_maj, di, min_prevalence,
            maj_prevalence, min_selected, min_count, maj_selected, maj_count, min_precision, min_recall, min_f1,
            maj_precision, maj_recall, maj_f1, tp_min, tn_min, fp_min, fn_min, tp_maj, tn_maj, fp_maj, fn_maj)
    return dict(zip(DI_KEYS, vals))


def results_from_counts(counts, out=print) -> dict:
    """counts[A][G][2][2] in ``ATTRIBUTES`` order -> the dict (and the stdout lines) of the reference's
    ``analyse_predictions`` (tone_bias_test.py:448-561)."""
    c = np.asarray(counts.cpu() if isinstance(counts, torch.Tensor) else counts)
    tone, sex, control, total_t = c[0], c[1], c[2], c[4][0]
    size = lambda t: int(t.sum())                                         # noqa: E731
    total = size(total_t)
    correct = int(total_t[0][0] + total_t[1][1])
    out(f"Total={total} correct={correct} my accuracy={correct / total:.3f}")
    light, dark, male, female, rich, poor = (size(tone[0]), size(tone[1]), size(sex[0]), size(sex[1]),
                                             size(control[0]), size(control[1]))
    out(f"dark {dark}")
    out(f"light {light}")
    out(f"male {male}")
    out(f"female {female}")
    out(f"total {total}")
    out(f"rich {rich}")
    out(f"poor {poor}")
    tp_all = int(total_t[1][1])
    tp_m, tp_f = int(sex[0][1][1]), int(sex[1][1][1])
    out(f"TP: male_count={tp_m} female_count={tp_f}")
    if tp_all > 0:
        out(f"TP: P(   male | mole=malignant ) = {tp_m / tp_all}")
        out(f"TP: P( female | mole=malignant ) = {tp_f / tp_all}")
    out(f"TP: male + female = {tp_m + tp_f}  total = {tp_all}")
    out()
    out(f"TEST_SET: male_count={male} female_count={female}")
    out(f"TEST_SET: P(   male ) = {male / total:.3f}")
    out(f"TEST_SET: P( female ) = {female / total:.3f}")
    out(f"TEST_SET: male + female = {male + female}  total = {total}")
    out()
    out(f"TEST_SET: light_count={light} dark_count={dark}")
    if total > 0:
        out(f"TEST_SET: P( light ) = {light / total:.3f}")
        out(f"TEST_SET: P(  dark ) = {dark / total:.3f}")
    out(f"TEST_SET: light + dark = {light + dark}  total = {total}")
    dark_pos, light_pos = int(tone[1][1].sum()), int(tone[0][1].sum())
    dark_prev = dark_pos / dark                 # unguarded in the reference (:529-530)
    light_prev = light_pos / light
    out(f"Dark Prevalence: {dark_pos} / {dark} = {dark_prev:.2f}")
    out(f"Light Prevalence: {light_pos} / {light} = {light_prev:.2f}")
    out("DISPARATE IMPACT: SKIN TONE")
    tone_di = di_from_tables(tone[1], tone[0])              # (dark = min, light = maj)   :538
    out("DISPARATE IMPACT: GENDER")
    gender_di = di_from_tables(sex[1], sex[0])              # (female = min, male = maj)  :540
    out("DISPARATE IMPACT: CONTROL")
    control_di = di_from_tables(control[1], control[0])     # (poor = min, rich = maj)    :542
    return {
        "correct": correct, "total": total, "accuracy": correct / total,
        "dark": dark, "light": light, "male": male, "female": female,
        "tone_di_results": tone_di, "gender_di_results": gender_di, "control_di_results": control_di,
    }


def results_from_type_counts(counts, out=print) -> dict:
    """Same result dict from the compact sharded-evaluation tensor counts[3][6][2][2] whose attributes
    are (Fitzpatrick type I..VI, sex {male,female}, control {rich,poor}): light = I+II, dark = III..VI
    (convert_type2tone, tone_bias_dataset.py:84-98); totals come from the type table because every
    evaluated image has exactly one type (tone_bias_dataset.py:191)."""
    c = np.asarray(counts.cpu() if isinstance(counts, torch.Tensor) else counts).astype(np.int64)
    full = np.zeros((5, N_GROUPS, 2, 2), np.int64)
    full[0, 0] = c[0, 0] + c[0, 1]
    full[0, 1] = c[0, 2:6].sum(0)
    full[1, :2] = c[1, :2]
    full[2, :2] = c[2, :2]
    full[3] = c[0]
    full[4, 0] = c[0].sum(0)
    return results_from_counts(full, out=out)


# -------------------------------------------------------------------------------------------------
# the reference's function surface
# -------------------------------------------------------------------------------------------------
def confusion_matrix(instances):
    """-> (tp, tn, fp, fn) instance dicts; sizes cross-checked against the kernel's counts."""
    tp_i, tn_i, fp_i, fn_i = {}, {}, {}, {}
    for index in sorted(instances.keys()):
        inst = instances[index]
        cell = (_class_code(inst["benign_malignant"]), _class_code(inst["prediction"]))
        if cell == (1, 1):
            tp_i[index] = inst
        elif cell == (0, 0):
            tn_i[index] = inst
        elif cell == (0, 1):
            fp_i[index] = inst
        elif cell == (1, 0):
            fn_i[index] = inst
    counts = _count_instances(instances, attributes=((None, (None,)),), n_groups=1)   # raises like the reference
    if _cells(counts[0][0]) != (len(tp_i), len(tn_i), len(fp_i), len(fn_i)):
        raise SiaError("confusion cell sizes disagree with the device reduction")
    return tp_i, tn_i, fp_i, fn_i


def values_counts(instances, feature, value):
    """Number of instances with ``instance[feature] == value`` (reference :274-280), counted on the device."""
    c = _count_instances_allow_any_class(instances, ((feature, (value,)),))
    return int(c[0][0].sum())


def filter(instances, feature, value):  # noqa: A001 - reference name
    """Sub-dict selection (reference :283-289).  Pure dict plumbing: stays on the host."""
    return {k: v for k, v in instances.items() if v[feature] == value}


def _count_instances_allow_any_class(instances, attributes):
    _keys, pred, label, groups = encode_instances(instances, attributes)
    if len(_keys) == 0:
        return np.zeros((len(attributes), 1, 2, 2), np.int64)
    dev = _device()
    pred, label = (pred == 1).astype(np.uint8), (label == 1).astype(np.uint8)
    return counts_from_arrays(torch.from_numpy(pred).to(dev), torch.from_numpy(label).to(dev),
                              torch.from_numpy(groups).to(dev), 1).cpu().numpy()


def disparate_impact_analysis(min_instances, maj_instances):
    """27-key result (reference :292-445).  Both groups go through one kernel launch."""
    merged, tag = {}, {}
    for gi, group in enumerate((min_instances, maj_instances)):
        for k, v in group.items():
            merged[(gi, k)] = v
            tag[(gi, k)] = gi
    # confusion_matrix() is applied per group in the reference -> ValueError per group
    for group in (min_instances, maj_instances):
        bad = [k for k, v in group.items() if _class_code(v["prediction"]) > 1 or _class_code(v["benign_malignant"]) > 1]
        if bad:
            confusion_matrix(group)
    if not merged:
        raise ZeroDivisionError("division by zero")
    dev = _device()
    keys = list(merged.keys())
    pred = np.array([_class_code(merged[k]["prediction"]) for k in keys], np.uint8)
    label = np.array([_class_code(merged[k]["benign_malignant"]) for k in keys], np.uint8)
    groups = np.array([[tag[k] for k in keys]], np.uint8)
    c = counts_from_arrays(torch.from_numpy(pred).to(dev), torch.from_numpy(label).to(dev),
                           torch.from_numpy(groups).to(dev), 2).cpu().numpy()
    return di_from_tables(c[0][0], c[0][1])


def analyse_predictions(instances):
    """Reference :448-561: same prints, same result dict, same ZeroDivisionError on empty groups."""
    if len(instances) == 0:
        raise ZeroDivisionError("division by zero")
    return results_from_counts(_count_instances(instances))


def _predicted_labels(model, images):
    """Predicted class index per image.  Models of this package return the label computed by the tail kernel itself
    (``predict``: first maximal index, the tie rule of ``torch.max(outputs, 1)``, reference :199); any other module
    goes through ``torch.max`` like the reference."""
    predict = getattr(model, "predict", None)
    if callable(predict):
        return predict(images)[1].long()
    return torch.max(model(images).data, 1)[1]


def predict_with_instance(model, device, test_loader, test_dataset, class_names):
    """Batched inference -> dict[index -> instance dict + 'prediction'] (reference :161-237).  ``images.to(device)``
    is also where a deferred DataLoader batch (tone_bias_dataset.DeferredBatch) runs its fused resize kernel."""
    model.eval()
    instances = dict()
    with torch.no_grad():
        for images, labels, indexes in test_loader:
            images = images.to(device)
            predicted = _predicted_labels(model, images).cpu().tolist()
            for i, pred in enumerate(predicted):
                index = int(indexes[i])
                instance = test_dataset.lookup_path(index)
                instance["prediction"] = class_names[pred]
                instances[index] = instance
    return instances


def evaluate_model(device, model, testloader):
    """Plain accuracy (reference :99-126)."""
    model.eval()
    correct = total = 0
    with torch.no_grad():
        for batch_number, (images, labels, indexes) in enumerate(testloader):
            images, labels = images.to(device), labels.to(device)
            print(f"BATCH {batch_number}: indexes {indexes}")
            predicted = _predicted_labels(model, images)
            total += labels.size(0)
            correct += (predicted == labels).sum().item()
    print(f"Accuracy of the network on the {len(testloader)} batches")
    print(f"test images: {correct/total:4f} (correct {correct} / total {total})")


def evaluate_model_by_class(device, model, testloader, class_names):
    """Per-class accuracy (reference :129-159)."""
    correct_pred = {c: 0 for c in class_names}
    total_pred = {c: 0 for c in class_names}
    with torch.no_grad():
        for images, labels, indexes in testloader:
            images, labels = images.to(device), labels.to(device)
            predictions = _predicted_labels(model, images)
            for label, prediction in zip(labels.cpu().tolist(), predictions.cpu().tolist()):
                if label == prediction:
                    correct_pred[class_names[label]] += 1
                total_pred[class_names[label]] += 1
    for classname, correct_count in correct_pred.items():
        n = total_pred[classname]
        print(f"    {correct_count} / {n}")
        accuracy = 100 * float(correct_count) / n if n > 0 else 0.0
        print(f"Accuracy for class: {classname:5s} is {accuracy:.1f} %")
