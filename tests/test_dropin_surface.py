"""CPU checks of the drop-in surface against fixtures produced by the reference's own classes / functions
(tests/golden/make_golden.py: gen_dataset_and_eval): ``HibaDataset`` (tone_bias_dataset.py:258-393),
``evaluate_model`` / ``evaluate_model_by_class`` stdout (tone_bias_test.py:99-159), ``predict_with_instance``
(:161-237) on a stand-in model, and the deferred transform through a real multi-process ``DataLoader``."""
import contextlib
import io
import json
import math
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from tests import helpers


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "dataset_eval.json")) as f:
        return json.load(f)


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b):
        return True
    return a == b


def test_hiba_dataset_matches_the_reference_class(tmp_path, golden):
    from skin_image_analysis_b200.tone_bias_dataset import HibaDataset
    g = golden["dataset"]
    df = helpers.synthetic_metadata_df(6, seed=21)
    imgs = helpers.write_image_files(str(tmp_path), df, 20, 28, seed=400, kind="noise")
    ds = HibaDataset(df, helpers.CLASS_NAMES, root_dir=str(tmp_path), transform=None)
    assert len(ds) == g["len"] and ds.get_class_names() == g["class_names"] and ds.get_class(1) == g["class_1"]
    for i in range(len(ds)):
        image, label, idx = ds[torch.tensor(i)] if i == 2 else ds[i]          # tensor indexes are accepted (:305-306)
        assert image.dtype == np.float32 and np.array_equal(np.asarray(image), np.float32(imgs[i]) / 255.0)
        assert float(np.float64(image).sum()) == g["image_sums"][i]
        assert (label, idx) == (g["labels"][i], g["indexes"][i])
        inst = ds.lookup_path(i)
        inst["file_path"] = os.path.relpath(inst["file_path"], str(tmp_path))
        want = g["lookup"][i]
        assert set(inst) == set(want)
        for k in want:
            assert _same(inst[k] if not isinstance(inst[k], np.generic) else inst[k].item(), want[k]), (i, k)
        assert np.array_equal(ds.read_u8(i), imgs[i])


def test_evaluate_model_stdout_equals_the_reference(golden):
    from skin_image_analysis_b200 import tone_bias_test as tt
    model, loader = helpers.FixedLogitModel(3), helpers.fixed_eval_loader(4, 6, seed=9)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tt.evaluate_model("cpu", model, loader)
    assert buf.getvalue() == golden["evaluate_model_stdout"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tt.evaluate_model_by_class("cpu", model, loader, helpers.CLASS_NAMES)
    assert buf.getvalue() == golden["evaluate_model_by_class_stdout"]


def test_predict_with_instance_equals_the_reference_on_a_stand_in_model(tmp_path, golden):
    from skin_image_analysis_b200 import tone_bias_test as tt
    from skin_image_analysis_b200.tone_bias_dataset import HibaDataset
    df = helpers.synthetic_metadata_df(6, seed=21)
    helpers.write_image_files(str(tmp_path), df, 20, 28, seed=400, kind="noise")
    ds = HibaDataset(df, helpers.CLASS_NAMES, root_dir=str(tmp_path), transform=None)
    model, loader = helpers.FixedLogitModel(3), helpers.fixed_eval_loader(4, 6, seed=9)
    small = [(im[:, :, :4, :4], lab % 2, idx % len(ds)) for im, lab, idx in loader][:1]
    inst = tt.predict_with_instance(model, "cpu", small, ds, helpers.CLASS_NAMES)
    want = golden["predict_with_instance"]
    assert sorted(inst) == sorted(int(k) for k in want)
    for k, v in inst.items():
        v = dict(v, file_path=os.path.relpath(v["file_path"], str(tmp_path)))
        for name, val in want[str(k)].items():
            got = v[name].item() if isinstance(v[name], np.generic) else v[name]
            assert _same(got, val), (k, name)


@pytest.mark.parametrize("workers,context", [(0, None), (2, "fork"), (2, "spawn")])
def test_dataloader_workers_defer_the_transform(tmp_path, workers, context):
    """The reference's call site -- DataLoader(HibaDataset(transform=Compose([Rescale((224,224)), ToTensor()])),
    batch_size, shuffle=True, num_workers=10) (tone_bias_test.py:617-637) -- with the drop-in transforms: worker
    processes cannot touch CUDA, so they hand back the uint8 decode buffers as a DeferredBatch (no CUDA call, hence
    runnable on a box without a GPU); materialising it on a CPU device is refused loudly."""
    import torchvision
    from skin_image_analysis_b200._lib import SiaError
    from skin_image_analysis_b200.tone_bias_dataset import DeferredBatch, HibaDataset, Rescale, ToTensor
    df = helpers.synthetic_metadata_df(7, seed=5)
    imgs = helpers.write_image_files(str(tmp_path), df, 45, 60, seed=700, kind="smooth")
    tf = torchvision.transforms.Compose([Rescale((224, 224), defer=True if workers == 0 else None), ToTensor()])
    ds = HibaDataset(df, helpers.CLASS_NAMES, root_dir=str(tmp_path), transform=tf)
    loader = DataLoader(ds, batch_size=3, shuffle=True, num_workers=workers, multiprocessing_context=context)
    seen = []
    for images, labels, indexes in loader:
        assert isinstance(images, DeferredBatch) and images.shape == (len(indexes), 3, 224, 224)
        assert labels.dtype == torch.int64 and len(images) == len(labels)
        for t, idx, lab in zip(images.images, indexes.tolist(), labels.tolist()):
            assert t.dtype == torch.uint8 and np.array_equal(t.numpy(), imgs[idx])
            assert lab == helpers.CLASS_NAMES.index(df.iloc[idx]["benign_malignant"])
        seen += indexes.tolist()
        with pytest.raises(SiaError):
            images.to("cpu")
    assert sorted(seen) == list(range(7))


def test_rescale_in_the_main_process_without_cuda_fails_loudly(tmp_path):
    from skin_image_analysis_b200._lib import SiaError
    from skin_image_analysis_b200.tone_bias_dataset import RandomCrop, Rescale
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    image = np.float32(helpers.synthetic_u8_image(20, 30, 1)) / 255.0
    with pytest.raises(SiaError):
        Rescale((8, 8))((image, 0, 0))
    deferred = Rescale((8, 8), defer=True)((image, 0, 0))
    with pytest.raises(SiaError):
        RandomCrop(4)(deferred)
