"""Drop-in for the host-side aggregation of ``tone_bias_analysis`` (reference src/tone_bias_analysis.py).

The evaluation path ends in a JSON-lines results file (one record per epoch, written by the caller of the hot
path, tone_bias_train.py:410-424) that this module reads back, flattens and averages across experiment folders.
Only the arithmetic is here; the matplotlib plotting of the reference's ``main``
(:513-632) is out of scope.  Same function names, arguments, return structures, exception types and float64
arithmetic as the reference:

  * ``compute_ci``        :12-39    mean +- score * std(ddof=0) / sqrt(n); Student t (n-1 dof) for n <= 30, else normal
  * ``get_files``         :42-47
  * ``get_measure`` / ``transpose_dict``   :281-319
  * ``read_experiment``   :324-398  flatten one folder of *.json result files into {global_epoch: record}
  * ``read_experiments``  :400-510  mean of every scalar measure per global epoch across the folders with a prefix

plus ``append_results_line`` -- the writer side of the format (tone_bias_train.py:410-424), so results produced by
this package's ``analyse_predictions`` / ``results_from_counts`` are consumed by the unmodified reference scripts.
"""
from __future__ import annotations

import json
import os

import numpy as np
import scipy.stats as stats

__all__ = ["compute_ci", "get_files", "get_measure", "transpose_dict", "read_experiment", "read_experiments",
           "append_results_line", "flatten_results"]


def compute_ci(data, confidence_level):
    """Two-sided confidence interval of the mean of ``data`` (reference :12-39): population standard deviation
    (``np.std``, ddof = 0), t quantile with n - 1 degrees of freedom up to 30 samples, normal quantile above."""
    n = len(data)
    centre = np.mean(data)
    spread = np.std(data)
    tail = 1 - (1 - confidence_level) / 2
    score = stats.t.ppf(tail, n - 1) if n <= 30 else stats.norm.ppf(tail)
    half_width = score * spread / np.sqrt(n)
    return (centre - half_width, centre + half_width)


def get_files(dir_path, ext):
    return [name for name in os.listdir(dir_path) if name.endswith(ext)]


def get_measure(results_dict, measure_name):
    """{global_epoch: record} -> the values of one measure in epoch (insertion) order."""
    return [results_dict[epoch][measure_name] for epoch in results_dict.keys()]


def transpose_dict(results_dict):
    """{global_epoch: record} -> {measure: [values by epoch]}; epoch 1 defines the measure names."""
    return {name: get_measure(results_dict, name) for name in results_dict[1].keys()}


def flatten_results(results: dict) -> dict:
    """Adds the top-level measures the plots use to one epoch record (reference :357-375): tone selection rates
    recomputed from the counts, the three disparate-impact values and f1.  ZeroDivisionError / KeyError propagate
    exactly as in the reference."""
    tone = results["tone_di_results"]
    results["tone_di_selection_rate_min"] = (tone["tp_min"] + tone["fp_min"]) / tone["min_count"]
    results["tone_di_selection_rate_maj"] = (tone["tp_maj"] + tone["fp_maj"]) / tone["maj_count"]
    results["tone_di"] = tone["di"]
    results["f1"] = tone["f1"]
    results["gender_di"] = results["gender_di_results"]["di"]
    results["control_di"] = results["control_di_results"]["di"]
    return results


def read_experiment(exp_path):
    """All ``*.json`` files of one experiment folder, in name (= date-time) order, as {global_epoch: record}.
    Epochs are renumbered consecutively across files; a record whose own epoch exceeds the running global
    epoch raises ValueError (reference :377-379)."""
    experiment_results = {}
    global_epoch = 1
    for name in sorted(get_files(exp_path, ".json")):
        with open(os.path.join(exp_path, name), "r") as json_file:
            for line in json_file:
                results = flatten_results(json.loads(line))
                epoch = results["epoch"]
                if epoch > global_epoch:
                    raise ValueError(f"Unexpected epoch {epoch}, greater than {global_epoch}")
                results["epoch"] = global_epoch
                experiment_results[global_epoch] = results
                global_epoch += 1
    return experiment_results


def read_experiments(experiments_folder, prefix, epoch_to_detail, out=print):
    """Mean of every non-dict measure per global epoch over the experiment folders whose name starts with
    ``prefix`` (reference :400-510).  Returns {epoch: {measure: mean}}; folders may have different lengths, each
    epoch is averaged over the folders that reached it."""
    experiments = {}
    for name in os.listdir(experiments_folder):
        if name.startswith(prefix):
            path = os.path.join(experiments_folder, name)
            experiments[path] = read_experiment(path)

    sums: dict[int, dict[str, float]] = {}
    counts: dict[tuple[int, str], int] = {}
    values: dict[tuple[int, str], list] = {}
    for path, experiment in experiments.items():
        out(f"FILE {path} epochs {len(experiment)}")
        for epoch, record in experiment.items():
            acc = sums.setdefault(epoch, {})
            for measure, value in record.items():
                if isinstance(value, dict):
                    continue
                if measure not in acc:
                    acc[measure] = 0.0
                    counts[(epoch, measure)] = 0
                acc[measure] += value
                counts[(epoch, measure)] += 1
                values.setdefault((epoch, measure), []).append(value)
            if epoch == epoch_to_detail:
                out(f"EPOCH DETAILS {epoch_to_detail} experiment {path}")
                _print_epoch_results(record, out)

    for epoch, acc in sums.items():
        for measure in acc:
            total, n = acc[measure], counts[(epoch, measure)]
            if epoch == epoch_to_detail:
                out(f"    AGGREGATE {measure} -> {total}   count {n}  avg={total / n:.4f}")
            acc[measure] = total / n
    return sums


def experiment_confidence_intervals(experiments_folder, prefix, confidence_level=0.90):
    """{(epoch, measure): (low, high)} of the per-folder values behind ``read_experiments`` (the reference
    computes these at :490-497 and discards them)."""
    per_key: dict[tuple[int, str], list] = {}
    for name in os.listdir(experiments_folder):
        if name.startswith(prefix):
            for epoch, record in read_experiment(os.path.join(experiments_folder, name)).items():
                for measure, value in record.items():
                    if not isinstance(value, dict):
                        per_key.setdefault((epoch, measure), []).append(value)
    return {k: compute_ci(v, confidence_level) for k, v in per_key.items()}


def _print_epoch_results(record, out=print):
    tone = record["tone_di_results"]
    out(f"    keys {record.keys()}")
    out(f"    accuracy {record['accuracy']}")
    out(f"    tone_di {record['tone_di']}")
    out(f"    tone_di_results {tone.keys()}")
    out("     TONE DI RESULTS")
    for key, value in tone.items():
        out(f"        [{key}] -> {value}")
    pos_min, pos_maj = tone["tp_min"] + tone["fp_min"], tone["tp_maj"] + tone["fp_maj"]
    sr_min, sr_maj = pos_min / tone["min_count"], pos_maj / tone["maj_count"]
    out(f"    selection_rate_min = {pos_min}/{tone['min_count']} = {sr_min:.4f}")
    out(f"    selection_rate_maj = {pos_maj}/{tone['maj_count']} = {sr_maj:.4f}")
    out(f"    tone_di = {sr_min:.4f}/{sr_maj:.4f} = {sr_min / sr_maj:.4f}")


def append_results_line(path_name, test_results: dict, avg_batch_loss, train_accuracy, epoch) -> dict:
    """Appends one epoch record in the reference's wire format (tone_bias_train.py:410-424): the
    ``analyse_predictions`` dict plus ``avg_batch_loss``, ``train_accuracy`` and ``epoch``, one JSON object per
    line."""
    record = dict(test_results)
    record["avg_batch_loss"] = avg_batch_loss
    record["train_accuracy"] = train_accuracy
    record["epoch"] = epoch
    with open(path_name, "a") as results_file:
        results_file.write(json.dumps(record))
        results_file.write("\n")
    return record
