#!/usr/bin/env python
"""Generate tests/golden/*.json|*.npz by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is pinned
  * analysis_synth.json : reference ``analyse_predictions`` / ``disparate_impact_analysis`` /
                          ``confusion_matrix`` outputs (result dict + captured stdout) on the seeded
                          instance dicts of tests/helpers.py (tone_bias_test.py:240-561).
  * model_<kind>.npz    : reference ``SkinCancerListModel`` / ``SkinCancerModel`` (tone_bias_model.py)
                          log-probabilities + a slice of every block output for the seeded weights of
                          oracle.model.synthetic_state_dict and the seeded input batch.
  * transform.npz       : reference ``Rescale`` + ``ToTensor`` (tone_bias_dataset.py:397-473) on seeded
                          u8 images, with ``skimage.transform.resize`` bound to the scipy restatement
                          (scikit-image itself is not installable here -- see oracle/__init__.py).
  * analysis_experiments.json : synthetic per-epoch result files + the reference's ``tone_bias_analysis``
                          outputs for them (read_experiment(s), transpose_dict, compute_ci; :12-39, :281-510).
  * transform_tv.npz    : the ToneClassifier test transform -- the ``transforms`` Compose built by the reference's
                          own ``ISIC(image_path, "Test")`` dataset (notebooks/ToneClassifier/CNNTrialDataset.py:71-76:
                          v2.Resize((224,224)) -> v2.ToDtype(float32, scale=True) -> v2.Normalize(ImageNet)) applied
                          to seeded u8 CHW tensors, run on the torch / torchvision of this image.
  * notebook_di.json    : hand-transcribed from the reference's saved notebook outputs
                          (notebooks/jgi_hiba_2022_torch.ipynb raw 3591-3617, 3643-3669, 3318-3322,
                          3433-3434, 3178); written by this script so the provenance is in one place.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import model as omodel          # noqa: E402
from oracle import ref_import               # noqa: E402
from tests import helpers                   # noqa: E402


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    return o


def gen_analysis(ref):
    cases = {}
    for name, n, seed, odd in [("n500", 500, 11, True), ("n64", 64, 12, False), ("n1087", 1087, 13, True)]:
        inst = helpers.synthetic_instances(n, seed, odd)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = ref.test.analyse_predictions(inst)
        tp, tn, fp, fn = ref.test.confusion_matrix(inst)
        cases[name] = {
            "n": n, "seed": seed, "with_oddities": odd, "result": _jsonable(res), "stdout": buf.getvalue(),
            "cells": [len(tp), len(tn), len(fp), len(fn)],
            "dark_count": ref.test.values_counts(inst, "skin_tone", "dark"),
        }
    with open(os.path.join(HERE, "analysis_synth.json"), "w") as f:
        json.dump(cases, f, indent=1, sort_keys=True)


def gen_model(ref):
    torch.manual_seed(0)
    for kind, cls in [(omodel.LIST_MODEL, ref.model.SkinCancerListModel),
                      (omodel.FOUR_CONV_MODEL, ref.hiba.SkinCancerModel)]:
        state = omodel.synthetic_state_dict(kind, seed=7)
        m = cls(helpers.CLASS_NAMES).eval()
        missing = m.load_state_dict(state, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        x = helpers.synthetic_batch_f32(4, 224, seed=21)
        with torch.no_grad():
            logp = m(x)
            # block outputs through the reference's own submodules
            feats = {}
            if kind == omodel.LIST_MODEL:
                h = x
                for i, layer in enumerate(m.layers):
                    h = layer(h)
                    if i in (2, 5, 8, 11, 14):
                        feats[f"after_{i}"] = h
            else:
                h = m.pool1(m.act1(m.conv1(x))); feats["conv1"] = h
                h = m.pool2(m.act2(m.conv2(h))); feats["conv2"] = h
                h = m.pool3(m.act3(m.conv3(h))); feats["conv3"] = h
                h = m.pool4(m.act4(m.conv4(h))); feats["conv4"] = h
        arrays = {"logp": logp.numpy(), "pred": torch.max(logp, 1)[1].numpy()}
        for k, v in feats.items():
            v = v.numpy()
            arrays["feat_" + k + "_sum"] = np.array([v.astype(np.float64).sum(), np.abs(v).astype(np.float64).sum()])
            arrays["feat_" + k + "_slice"] = v.reshape(v.shape[0], -1)[:, ::max(1, v[0].size // 64)][:, :64].copy()
        np.savez_compressed(os.path.join(HERE, f"model_{kind}.npz"), **arrays)


def gen_transform(ref):
    arrays = {}
    tf_tuple = [ref.dataset.Rescale((224, 224)), ref.dataset.ToTensor()]
    cases = [("noise_450x600_224", 450, 600, 31, "noise", (224, 224)),
             ("smooth_450x600_224", 450, 600, 32, "smooth", (224, 224)),
             ("extremes_450x600_224", 450, 600, 33, "extremes", (224, 224)),
             ("noise_450x600_512", 450, 600, 34, "noise", (512, 512)),
             ("noise_97x131_int64", 97, 131, 35, "noise", 64),
             ("noise_131x97_int48", 131, 97, 36, "noise", 48)]
    for name, h, w, seed, kind, size in cases:
        u8 = helpers.synthetic_u8_image(h, w, seed, kind)
        image = np.float32(u8) / 255.0             # tone_bias_dataset.py:335
        sample = (image, 1, 42)
        sample = ref.dataset.Rescale(size)(sample)
        t, label, idx = ref.dataset.ToTensor()(sample)
        assert (label, idx) == (1, 42)
        t = t.contiguous().numpy()
        arrays[name + "_shape"] = np.array(t.shape)
        arrays[name + "_sum"] = np.array([t.astype(np.float64).sum()])
        if t.size <= 40000:
            arrays[name + "_full"] = t
        else:
            arrays[name + "_sub"] = t[:, ::7, ::5].copy()
    del tf_tuple
    np.savez_compressed(os.path.join(HERE, "transform.npz"), **arrays)


TV_CASES = [("noise_450x600", 450, 600, 41, "noise"), ("smooth_450x600", 450, 600, 42, "smooth"),
            ("extremes_450x600", 450, 600, 43, "extremes"), ("noise_97x131", 97, 131, 44, "noise"),
            ("noise_600x450", 600, 450, 45, "noise"), ("noise_160x200_up", 160, 200, 46, "noise")]


def gen_transform_tv():
    """The reference's own Compose, obtained by constructing its Dataset on a stub directory."""
    import tempfile
    sys.path.insert(0, "/root/reference/notebooks/ToneClassifier")
    import CNNTrialDataset as ref_ds                                   # the reference module, unmodified
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "testmeta.csv"), "w") as f:
            f.write("isic_id,fitzpatrick_skin_type\nISIC_0000000,II\n")
        open(os.path.join(d, "ISIC_0000000.JPG"), "wb").close()
        ds = ref_ds.ISIC(d, "Test")
    arrays = {}
    for name, h, w, seed, kind in TV_CASES:
        u8 = helpers.synthetic_u8_image(h, w, seed, kind)
        t = ds.transforms(torch.from_numpy(u8).permute(2, 0, 1).contiguous()).contiguous().numpy()
        assert t.shape == (3, 224, 224) and t.dtype == np.float32
        arrays[name + "_sum"] = np.array([t.astype(np.float64).sum()])
        arrays[name + "_sub"] = t[:, ::5, ::3].copy()
    np.savez_compressed(os.path.join(HERE, "transform_tv.npz"), **arrays)


def gen_notebook():
    nb = {
        "source": "notebooks/jgi_hiba_2022_torch.ipynb (saved cell outputs)",
        "tone": {"raw_lines": "3591-3617",
                 "cells": {"tp_min": 6, "tn_min": 150, "fp_min": 11, "fn_min": 17,
                           "tp_maj": 113, "tn_maj": 606, "fp_maj": 48, "fn_maj": 136},
                 "printed": {"min_prevalence": 0.125, "maj_prevalence": 0.276, "min_precision": 0.353,
                             "min_recall": 0.261, "min_f1": 0.300, "maj_precision": 0.702, "maj_recall": 0.454,
                             "maj_f1": 0.551, "f1": 0.529, "min_group_accuracy": 0.848,
                             "maj_group_accuracy": 0.796, "min_selected": 17, "min_count": 184,
                             "maj_selected": 161, "maj_count": 903, "selection_rate_min": 0.092,
                             "selection_rate_maj": 0.178, "di": 0.518, "di_inverse": 1.930}},
        "sex": {"raw_lines": "3643-3669",
                "cells": {"tp_min": 48, "tn_min": 414, "fp_min": 33, "fn_min": 75,
                          "tp_maj": 70, "tn_maj": 342, "fp_maj": 26, "fn_maj": 78},
                "printed": {"min_prevalence": 0.216, "maj_prevalence": 0.287, "min_precision": 0.593,
                            "min_recall": 0.390, "min_f1": 0.471, "maj_precision": 0.729, "maj_recall": 0.473,
                            "maj_f1": 0.574, "f1": 0.527, "min_group_accuracy": 0.811,
                            "maj_group_accuracy": 0.798, "min_selected": 81, "min_count": 570,
                            "maj_selected": 96, "maj_count": 516, "selection_rate_min": 0.142,
                            "selection_rate_maj": 0.186, "di": 0.764, "di_inverse": 1.309}},
        "sizes": {"raw_lines": "3318-3322", "dark": 184, "light": 903, "male": 516, "female": 570, "total": 1087},
        "prevalence": {"raw_lines": "3433-3434", "dark_pos": 23, "light_pos": 249},
        "overall": {"raw_lines": "3178", "correct": 875, "total": 1087, "accuracy": 0.805},
    }
    with open(os.path.join(HERE, "notebook_di.json"), "w") as f:
        json.dump(nb, f, indent=1, sort_keys=True)


def experiment_lines():
    """The JSON-lines result files of three synthetic experiment folders (what tone_bias_train.py:410-424 appends
    after every epoch), built from the reference's own analyse_predictions on seeded instances."""
    ref = ref_import.load()
    folders = {}
    for f, (name, n_files, epochs) in enumerate([("balanced_2024-09-21_00-38-39", 2, [3, 2]),
                                                 ("balanced_2024-09-22_10-18-46", 1, [4]),
                                                 ("imbalanced_2024-09-23_09-00-00", 1, [2])]):
        files = {}
        for k in range(n_files):
            lines = []
            for e in range(1, epochs[k] + 1):
                inst = helpers.synthetic_instances(300 + 20 * e, 1000 + 100 * f + 10 * k + e, with_oddities=(e % 2 == 0))
                with contextlib.redirect_stdout(io.StringIO()):
                    res = ref.test.analyse_predictions(inst)
                res["avg_batch_loss"] = 0.7 / (e + k + f + 1)
                res["train_accuracy"] = 0.5 + 0.04 * (e + k) + 0.01 * f
                res["epoch"] = e
                lines.append(json.dumps(_jsonable(res)))
            files[f"2024-09-2{1 + f}_0{k}-00-00.json"] = lines
        folders[name] = files
    return folders


def gen_experiments(ref):
    """analysis_experiments.json: inputs (the result files) + what the reference's tone_bias_analysis returns
    for them (read_experiment, read_experiments, transpose_dict, get_measure, compute_ci; :12-39, :281-510)."""
    import tempfile
    ana = ref.analysis
    folders = experiment_lines()
    out = {"folders": folders, "read_experiment": {}, "read_experiments": {}, "transpose": {}, "compute_ci": []}
    with tempfile.TemporaryDirectory() as tmp:
        for name, files in folders.items():
            os.makedirs(os.path.join(tmp, name))
            for fname, lines in files.items():
                with open(os.path.join(tmp, name, fname), "w") as f:
                    f.write("\n".join(lines) + "\n")
        with contextlib.redirect_stdout(io.StringIO()):
            for name in folders:
                out["read_experiment"][name] = _jsonable(ana.read_experiment(os.path.join(tmp, name)))
            for prefix in ("balanced", "imbalanced"):
                avg = ana.read_experiments(tmp, prefix, 2)
                out["read_experiments"][prefix] = _jsonable(avg)
                out["transpose"][prefix] = _jsonable(ana.transpose_dict(avg))
    rng = np.random.default_rng(3)
    for n, level in [(2, 0.90), (5, 0.95), (30, 0.90), (31, 0.90), (200, 0.99)]:
        data = rng.normal(0.6, 0.1, n).tolist()
        lo, hi = ana.compute_ci(data, level)
        out["compute_ci"].append({"data": data, "level": level, "low": float(lo), "high": float(hi)})
    with open(os.path.join(HERE, "analysis_experiments.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def gen_optuna_best(ref):
    """model_optuna_best.npz: the reference's own ``tone_bias_optuna.create_best_model()`` (:116-120; conv 192 / 172 /
    22 / 86, linear 227 / 80 / 86) built under torch.manual_seed(123) -- the drop-in builder creates the same layers
    in the same order, so the same seed reproduces the weights -- and its log-probabilities for a seeded batch."""
    import importlib
    ro = importlib.import_module("tone_bias_optuna")
    torch.manual_seed(123)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ro.create_best_model().eval()
    x = helpers.synthetic_batch_f32(3, 224, seed=33)
    with torch.no_grad():
        logp = model(x)
    state = model.state_dict()
    checks = {k.replace(".", "_") + "_sum": float(v.double().sum()) for k, v in state.items()}
    np.savez(os.path.join(HERE, "model_optuna_best.npz"), logp=logp.numpy(), pred=logp.argmax(1).numpy(),
             keys=np.array(list(state.keys())), **{k: np.float64(v) for k, v in checks.items()})
    # the oracle restatement must agree with the reference class on the same weights
    got = omodel.forward_sequential(state, x)
    assert float((got - logp).abs().max()) < 1e-6


def _tiny_storage_(param, shape_of_one):
    """Replaces a big parameter by a stride-0 expansion of a small random block: torch.save keeps the small storage,
    so a genuine whole-module pickle of the reference class stays under a megabyte.  The model is still a perfectly
    valid set of weights -- its reference outputs are what the fixture pins."""
    if param.dim() == 2:          # a Linear: one positive scalar, sized so that the pre-activations stay O(1)
        small = torch.full(shape_of_one, 4.0 / param.shape[1])
    else:
        small = torch.randn(shape_of_one) * (2.0 / float(np.prod(param.shape[1:])) ** 0.5)
    param.data = small.expand(param.shape)


def gen_session_models(ref):
    """session_model_<kind>.pth: WHOLE-MODULE pickles written by the reference's own ``save_model``
    (tone_bias_model.py:305-315) from the reference's own classes (module paths ``tone_bias_model.SkinCancerListModel``
    and ``jgi_hiba_2022_model.SkinCancerModel``), plus the reference's log-probabilities for a seeded batch
    (session_models.npz).  ``load_model`` of the drop-in must open these files."""
    import tempfile
    out = {}
    x = helpers.synthetic_batch_f32(3, 224, seed=55) * torch.tensor([0.3, 0.65, 1.0]).view(3, 1, 1, 1)
    for kind, mod, cls in [(omodel.LIST_MODEL, ref.model, "SkinCancerListModel"),
                           (omodel.FOUR_CONV_MODEL, ref.hiba, "SkinCancerModel")]:
        torch.manual_seed(77)
        m = getattr(mod, cls)(helpers.CLASS_NAMES).eval()
        with torch.no_grad():                      # He-scaled weights: the activations (and the logits) vary per image
            for p in m.parameters():
                if p.dim() > 1:
                    p.copy_(torch.randn(p.shape) * (2.0 / p[0].numel()) ** 0.5)
                else:
                    p.copy_((torch.rand(p.shape) - 0.5) * 0.2)
        big = [p for _n, p in m.named_parameters() if p.numel() > 200_000]
        for p in big:
            _tiny_storage_(p, (1,) + tuple(p.shape[1:]) if p.dim() == 4 else (1, 1))
        # size the first Linear's scalar on the actual features so that its pre-activations stay O(1)
        first_fc = next(mm for mm in m.modules() if isinstance(mm, torch.nn.Linear))
        seen = {}
        def _see(_m, inp, _o):
            seen["s"] = float(inp[0].sum(1).max())
        hook = first_fc.register_forward_hook(_see)
        with torch.no_grad():
            m(x)
        hook.remove()
        first_fc.weight.data = torch.full((1, 1), 2.0 / seen["s"]).expand(first_fc.weight.shape)
        with torch.no_grad():
            # give the head a visible margin between images
            logp = m(x)
        path = os.path.join(HERE, f"session_model_{kind}.pth")
        mod.save_model(m, path)
        assert os.path.getsize(path) < 1_000_000, os.path.getsize(path)
        again = mod.load_model(path, helpers.CLASS_NAMES)             # the reference's own loader reads it back
        with torch.no_grad():
            assert torch.equal(again(x), logp)
        out[kind + "_logp"] = logp.numpy()
        out[kind + "_class"] = np.array(type(m).__module__ + "." + type(m).__qualname__)
    np.savez(os.path.join(HERE, "session_models.npz"), **out)


DEEP_TRIALS = {
    # define_isic_model search-space corners (tone_bias_optuna.py:123-173): 5 and 7 pooling blocks -> 7x7 and 1x1 maps
    "deep5": {"n_conv_layers": 4, "n_units_l0": 16, "n_units_conv_l0": 24, "n_units_conv_l1": 16, "n_units_conv_l2": 40,
              "n_units_conv_l3": 16, "n_linear_layers": 2, "n_units_linear_l0": 32, "dropout_l0": 0.3,
              "n_units_linear_l1": 16, "dropout_l1": 0.3},
    "deep7": {"n_conv_layers": 6, "n_units_l0": 16, "n_units_conv_l0": 16, "n_units_conv_l1": 32, "n_units_conv_l2": 16,
              "n_units_conv_l3": 48, "n_units_conv_l4": 16, "n_units_conv_l5": 64, "n_linear_layers": 3,
              "n_units_linear_l0": 64, "dropout_l0": 0.2, "n_units_linear_l1": 32, "dropout_l1": 0.2,
              "n_units_linear_l2": 16, "dropout_l2": 0.2},
}


def gen_deep_sequential(ref):
    """session_model_<deepN>.pth: the reference's ``define_isic_model`` (an ``nn.Sequential``) with 5 / 7 pooling blocks,
    He-initialised so the activations survive, pickled whole; + its CPU log-probabilities.  Pins MaxPool2d's floor on
    odd sizes (14 -> 7 -> 3 -> 1) and ``load_model`` on a Sequential pickle."""
    import importlib
    ro = importlib.import_module("tone_bias_optuna")
    out = {}
    x = helpers.synthetic_batch_f32(5, 224, seed=66)
    for name, hp in DEEP_TRIALS.items():
        torch.manual_seed(88)
        with contextlib.redirect_stdout(io.StringIO()):
            m = ro.define_isic_model(2, ro.TrialDummy(dict(hp))).eval()
        gen = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for k, v in m.state_dict().items():
                if k.endswith(".weight"):
                    v.copy_(torch.randn(v.shape, generator=gen) * (2.0 / v[0].numel()) ** 0.5)
                else:
                    v.copy_((torch.rand(v.shape, generator=gen) - 0.5) * 0.2)
            logp = m(x)
        path = os.path.join(HERE, f"session_model_{name}.pth")
        torch.save(m, path)                                            # what tone_bias_model.save_model does (:315)
        assert os.path.getsize(path) < 1_000_000, os.path.getsize(path)
        out[name + "_logp"] = logp.numpy()
    np.savez(os.path.join(HERE, "deep_sequential.npz"), **out)


def gen_dataset_and_eval(ref):
    """dataset_eval.json: (a) the reference's own ``HibaDataset`` (tone_bias_dataset.py:258-393) on a seeded dataframe
    and lossless image files -- ``len``, ``lookup_path`` dicts, labels, the image ``__getitem__`` returns without a
    transform (``skimage.io.imread`` is bound to a PIL reader: scikit-image is not installable here);
    (b) stdout of the reference's ``evaluate_model`` / ``evaluate_model_by_class`` (tone_bias_test.py:99-159) on the
    seeded stand-in model / loader of tests/helpers.py; (c) the reference's ``predict_with_instance`` (:161-237) on the
    same stand-ins with the reference dataset supplying ``lookup_path``."""
    import tempfile
    from PIL import Image
    sk_io = sys.modules["skimage.io"]
    sk_io.imread = lambda path: np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
    ref.dataset.skimage.io.imread = sk_io.imread
    out = {}
    df = helpers.synthetic_metadata_df(6, seed=21)
    with tempfile.TemporaryDirectory() as d:
        imgs = helpers.write_image_files(d, df, 20, 28, seed=400, kind="noise")
        ds = ref.dataset.HibaDataset(df, helpers.CLASS_NAMES, root_dir=d, transform=None)
        items = [ds[i] for i in range(len(ds))]
        for (im, _l, _i), u8 in zip(items, imgs):
            assert im.dtype == np.float32 and np.array_equal(im, np.float32(u8) / 255.0)
        look = []
        for i in range(len(ds)):
            inst = ds.lookup_path(i)
            inst["file_path"] = os.path.relpath(inst["file_path"], d)
            look.append(_jsonable(inst))
        out["dataset"] = {"len": len(ds), "labels": [int(it[1]) for it in items], "indexes": [int(it[2]) for it in items],
                          "lookup": look, "class_names": ds.get_class_names(), "class_1": ds.get_class(1),
                          "image_sums": [float(np.float64(it[0]).sum()) for it in items]}
        model, loader = helpers.FixedLogitModel(3), helpers.fixed_eval_loader(4, 6, seed=9)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.test.evaluate_model("cpu", model, loader)
        out["evaluate_model_stdout"] = buf.getvalue()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.test.evaluate_model_by_class("cpu", model, loader, helpers.CLASS_NAMES)
        out["evaluate_model_by_class_stdout"] = buf.getvalue()
        small = [(im[:, :, :4, :4], lab % 2, idx % len(ds)) for im, lab, idx in loader][:1]
        inst = ref.test.predict_with_instance(model, "cpu", small, ds, helpers.CLASS_NAMES)
        for v in inst.values():
            v["file_path"] = os.path.relpath(v["file_path"], d)
        out["predict_with_instance"] = _jsonable(inst)
    with open(os.path.join(HERE, "dataset_eval.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def main():
    ref = ref_import.load()
    if "--new" in sys.argv:                       # only the fixtures added in round 2 (the others stay untouched)
        gen_session_models(ref)
        gen_deep_sequential(ref)
        gen_dataset_and_eval(ref)
        return
    gen_session_models(ref)
    gen_deep_sequential(ref)
    gen_dataset_and_eval(ref)
    gen_notebook()
    gen_optuna_best(ref)
    gen_experiments(ref)
    gen_analysis(ref)
    gen_transform(ref)
    gen_transform_tv()
    gen_model(ref)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
