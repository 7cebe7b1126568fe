"""GPU checks of the drop-in surface (rows a1 / a6 / a7 / a12 / f3 of SURVEY section 8): ``load_model`` on whole-module
pickles WRITTEN BY THE REFERENCE'S OWN CLASSES, ``predict_with_instance`` through a real ``DataLoader`` over a
``HibaDataset`` of image files (main-process and multi-process), ``evaluate_model(_by_class)``, floor pooling on odd
sizes, and the eval-mode BatchNorm fold."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from oracle import analysis as oa
from oracle import model as om
from oracle import resize as R
from tests import helpers

pytestmark = pytest.mark.gpu

LOGP_TOL = 1e-2


@pytest.mark.parametrize("kind", [om.LIST_MODEL, om.FOUR_CONV_MODEL])
def test_load_model_opens_the_reference_whole_module_pickle(golden_dir, kind):
    """session_model_<kind>.pth was written by the reference's ``save_model`` from the reference's class
    (tone_bias_model.py:305-315; module path tone_bias_model / jgi_hiba_2022_model): the drop-in ``load_model`` resolves
    the class path to this package, keeps the parameters, and the CUDA forward reproduces the reference's CPU
    log-probabilities for the seeded batch."""
    from skin_image_analysis_b200 import tone_bias_model as tm
    g = np.load(os.path.join(golden_dir, "session_models.npz"))
    model = tm.load_model(os.path.join(golden_dir, f"session_model_{kind}.pth"), helpers.CLASS_NAMES)
    assert isinstance(model, getattr(tm, kind)) and model.get_class_names() == helpers.CLASS_NAMES
    assert list(model.state_dict().keys()) == list(om.param_shapes(kind).keys())
    model = model.to("cuda").eval()
    x = helpers.synthetic_batch_f32(3, 224, seed=55) * torch.tensor([0.3, 0.65, 1.0]).view(3, 1, 1, 1)
    logp, pred = model.predict(x.cuda())
    want = g[kind + "_logp"]
    assert np.abs(logp.cpu().numpy() - want).max() <= LOGP_TOL
    assert np.array_equal(pred.cpu().numpy(), want.argmax(1))
    assert torch.equal(model(x.cuda()), logp)


def test_save_model_round_trip_and_bare_state_dict(tmp_path):
    from skin_image_analysis_b200 import tone_bias_model as tm
    state = om.synthetic_state_dict(om.LIST_MODEL, seed=12)
    m = tm.SkinCancerListModel(helpers.CLASS_NAMES)
    m.load_state_dict(state)
    x = helpers.synthetic_batch_f32(2, 224, seed=8).cuda()
    want = m.cuda().eval()(x)
    path = str(tmp_path / "session_model.pth")
    tm.save_model(m.cpu(), path)                                   # whole-module pickle, like the reference (:315)
    again = tm.load_model(path, helpers.CLASS_NAMES).cuda().eval()
    assert isinstance(again, tm.SkinCancerListModel) and torch.equal(again(x), want)
    torch.save(state, path)                                         # a bare state_dict is accepted too
    again = tm.load_model(path, helpers.CLASS_NAMES).cuda().eval()
    assert torch.equal(again(x), want)


@pytest.mark.parametrize("name", ["deep5", "deep7"])
def test_load_model_on_a_reference_sequential_with_odd_pooling(golden_dir, name):
    """The reference's ``define_isic_model`` with 5 / 7 pooling blocks (tone_bias_optuna.py:123-173; MaxPool2d floors
    14 -> 7 -> 3 -> 1), pickled whole by the reference, against the reference's CPU log-probabilities."""
    from skin_image_analysis_b200 import tone_bias_model as tm
    from skin_image_analysis_b200.tone_bias_optuna import _B200Sequential
    g = np.load(os.path.join(golden_dir, "deep_sequential.npz"))
    model = tm.load_model(os.path.join(golden_dir, f"session_model_{name}.pth"), helpers.CLASS_NAMES)
    assert isinstance(model, _B200Sequential)
    model = model.cuda().eval()
    x = helpers.synthetic_batch_f32(5, 224, seed=66)
    logp, pred = model.predict(x.cuda())
    want = g[name + "_logp"]
    tol = 4 * LOGP_TOL                           # He-scaled weights through 8-10 bf16 layers (see the optuna test)
    assert np.abs(logp.cpu().numpy() - want).max() <= tol
    safe = np.abs(want[:, 1] - want[:, 0]) > 2 * tol
    assert np.array_equal(pred.cpu().numpy()[safe], want.argmax(1)[safe])
    plan = model._plan()
    assert plan.valid[-1] == (7 if name == "deep5" else 1) and any(plan.needs_pad) == (name == "deep7")
    # a digit-keyed state_dict of the same Sequential is rebuilt into the same model
    state = {k: v.cpu() for k, v in model.state_dict().items()}
    rebuilt = tm._adopt(state, helpers.CLASS_NAMES).cuda().eval()
    assert torch.equal(rebuilt(x.cuda()), logp)


def test_pad_nhwc_kernel():
    from skin_image_analysis_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(3, 7, 7, 64, device="cuda", generator=g).to(torch.bfloat16)
    out = ops.pad_nhwc(x, (7, 7), (8, 8))
    assert out.shape == (3, 8, 8, 64) and torch.equal(out[:, :7, :7], x)
    assert bool((out[:, 7] == 0).all()) and bool((out[:, :, 7] == 0).all())
    out = ops.pad_nhwc(x[:, :4, :4].contiguous(), (3, 3), (4, 4))
    assert torch.equal(out[:, :3, :3], x[:, :3, :3]) and bool((out[:, 3] == 0).all()) and bool((out[:, :, 3] == 0).all())


def test_batchnorm_after_conv_is_folded():
    """The BatchNorm2d the reference keeps commented out between conv and ReLU (tone_bias_model.py:88), enabled
    (``batch_norm=True``) with non-trivial running statistics: the folded CUDA forward against torch's own eval-mode
    BatchNorm on the CPU."""
    from skin_image_analysis_b200 import tone_bias_model as tm
    torch.manual_seed(4)
    m = tm.SkinCancerListModel(helpers.CLASS_NAMES, batch_norm=True)
    gen = torch.Generator().manual_seed(6)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=gen) * 0.05)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=gen) * 0.5 + 0.75)
                mod.weight.copy_(torch.rand(mod.num_features, generator=gen) + 0.5)
                mod.bias.copy_(torch.randn(mod.num_features, generator=gen) * 0.05)
    x = helpers.synthetic_batch_f32(4, 224, seed=77)
    m.eval()
    with torch.no_grad():
        want = torch.nn.Module.__call__(_CpuReference(m), x)
    got = m.cuda()(x.cuda()).cpu()
    assert (got - want).abs().max().item() <= LOGP_TOL


class _CpuReference(torch.nn.Module):
    """Runs the drop-in module's OWN children with torch's eager CPU kernels (the layers are stock torch.nn modules)."""

    def __init__(self, model):
        super().__init__()
        self.layers = model.layers

    def forward(self, x):
        return self.layers(x)


def _build_dataset(tmp_path, n, seed, defer=None):
    import torchvision
    from skin_image_analysis_b200.tone_bias_dataset import HibaDataset, Rescale, ToTensor
    df = helpers.synthetic_metadata_df(n, seed=seed)
    imgs = helpers.write_image_files(str(tmp_path), df, 450, 600, seed=900 + seed, kind="smooth")
    tf = torchvision.transforms.Compose([Rescale((224, 224), defer=defer), ToTensor()])
    return df, imgs, HibaDataset(df, helpers.CLASS_NAMES, root_dir=str(tmp_path), transform=tf)


def test_predict_with_instance_through_real_dataloaders(tmp_path):
    """The reference's evaluation main (tone_bias_test.py:617-652) with only the imports changed: HibaDataset over
    image files + Compose([Rescale((224,224)), ToTensor()]) + DataLoader(shuffle=True) + model.to(device) +
    predict_with_instance + analyse_predictions, with num_workers = 0 (per-sample kernel launches in the main process),
    2 forked workers and 2 spawned workers (deferred, batched launches).  All three give the same instances; labels
    equal the CPU oracle's outside the margin band; the printed analysis equals the oracle's for the same instances."""
    from skin_image_analysis_b200 import tone_bias_model as tm
    from skin_image_analysis_b200 import tone_bias_test as tt
    n = 10
    df, imgs, ds = _build_dataset(tmp_path, n, seed=31)
    state = om.synthetic_state_dict(om.LIST_MODEL, seed=13)
    x_ref = torch.from_numpy(np.stack([R.transform_u8(im, (224, 224)) for im in imgs]))
    ref = om.forward(om.LIST_MODEL, state, x_ref)
    state["layers.16.bias"][1] -= float((ref[:, 1] - ref[:, 0]).median())
    ref = om.forward(om.LIST_MODEL, state, x_ref)
    want_pred = ref.argmax(1).numpy()
    safe = ((ref[:, 1] - ref[:, 0]).abs() > 2 * LOGP_TOL).numpy()
    assert safe.sum() >= n // 2 and 0 < want_pred.sum() < n

    model = tm.SkinCancerListModel(helpers.CLASS_NAMES)
    model.load_state_dict(state)
    device = torch.device("cuda:0")
    model = model.to(device)
    results = {}
    for workers, ctx in [(0, None), (2, "fork"), (2, "spawn")]:
        torch.manual_seed(workers)
        loader = DataLoader(ds, batch_size=4, shuffle=True, num_workers=workers, multiprocessing_context=ctx)
        instances = tt.predict_with_instance(model, device, loader, ds, helpers.CLASS_NAMES)
        assert sorted(instances) == list(range(n))
        for i in range(n):
            inst = instances[i]
            assert inst["image_name"] == df.iloc[i]["isic_id"] and inst["benign_malignant"] == df.iloc[i]["benign_malignant"]
            if safe[i]:
                assert inst["prediction"] == helpers.CLASS_NAMES[want_pred[i]], (workers, ctx, i)
        results[(workers, ctx)] = {i: instances[i]["prediction"] for i in range(n)}
    assert results[(0, None)] == results[(2, "fork")] == results[(2, "spawn")]

    buf_got, buf_want = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(buf_got):
        got = tt.analyse_predictions(instances)
    want = oa.analyse_predictions(instances, out=lambda *a: print(*a, file=buf_want))
    assert got == want and buf_got.getvalue() == buf_want.getvalue()


def test_deferred_batch_equals_per_sample_rescale(tmp_path):
    """A DeferredBatch materialised on the GPU == the per-sample Rescale + ToTensor + default collate, bit for bit, and
    both == the oracle transform within the fp32 tolerance of the preprocess tests; mixed source shapes are grouped."""
    from skin_image_analysis_b200.tone_bias_dataset import DeferredBatch, Rescale, ToTensor
    from torch.utils.data import default_collate
    shapes = [(450, 600), (300, 400), (450, 600), (300, 400), (450, 600)]
    u8s = [helpers.synthetic_u8_image(h, w, 50 + i, "smooth") for i, (h, w) in enumerate(shapes)]
    samples = [(np.float32(u) / 255.0, i % 2, i) for i, u in enumerate(u8s)]
    eager = default_collate([ToTensor()(Rescale((224, 224))(s)) for s in samples])
    lazy = default_collate([ToTensor()(Rescale((224, 224), defer=True)(s)) for s in samples])
    assert isinstance(lazy[0], DeferredBatch) and torch.equal(lazy[1], eager[1]) and torch.equal(lazy[2], eager[2])
    out = lazy[0].to(torch.device("cuda:0"))
    assert out.shape == (5, 3, 224, 224) and out.dtype == torch.float32 and out.is_cuda
    assert torch.equal(out.cpu(), eager[0])
    want = np.stack([R.transform_u8(u, (224, 224)) for u in u8s])
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-6
    with pytest.raises(RuntimeError):
        default_collate([Rescale(64, defer=True)(samples[0]), Rescale(64, defer=True)(
            (np.float32(helpers.synthetic_u8_image(100, 300, 1)) / 255.0, 0, 9))])[0].to("cuda")


def test_evaluate_model_on_the_gpu(tmp_path):
    """``evaluate_model`` / ``evaluate_model_by_class`` (tone_bias_test.py:99-159) with the real model and loader: the
    printed lines are the reference's format, filled with the counts of the kernel's own predictions."""
    from skin_image_analysis_b200 import tone_bias_model as tm
    from skin_image_analysis_b200 import tone_bias_test as tt
    n = 7
    df, imgs, ds = _build_dataset(tmp_path, n, seed=41, defer=True)
    state = om.synthetic_state_dict(om.LIST_MODEL, seed=14)
    model = tm.SkinCancerListModel(helpers.CLASS_NAMES)
    model.load_state_dict(state)
    device = torch.device("cuda:0")
    model = model.to(device).eval()
    loader = DataLoader(ds, batch_size=3, shuffle=False, num_workers=0)
    x = torch.from_numpy(np.stack([R.transform_u8(im, (224, 224)) for im in imgs])).to(device)
    pred = model.predict(x)[1].cpu().numpy()
    label = np.array([helpers.CLASS_NAMES.index(v) for v in df["benign_malignant"]])
    correct = int((pred == label).sum())
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tt.evaluate_model(device, model, loader)
    lines = buf.getvalue().splitlines()
    assert lines[0] == "BATCH 0: indexes tensor([0, 1, 2])" and lines[2] == "BATCH 2: indexes tensor([6])"
    assert lines[3] == "Accuracy of the network on the 3 batches"
    assert lines[4] == f"test images: {correct / n:4f} (correct {correct} / total {n})"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tt.evaluate_model_by_class(device, model, loader, helpers.CLASS_NAMES)
    want = []
    for c, name in enumerate(helpers.CLASS_NAMES):
        tot, ok = int((label == c).sum()), int(((label == c) & (pred == c)).sum())
        want += [f"    {ok} / {tot}", f"Accuracy for class: {name:5s} is {(100 * float(ok) / tot if tot else 0.0):.1f} %"]
    assert buf.getvalue().splitlines() == want


def test_full_size_bench_config_vs_cpu_oracle():
    """BASELINE configs[1] at FULL size against the oracle itself: 256 ISIC-shaped 600x450 uint8 images through the
    engine exactly as bench.py drives it (two-product tensor-core preprocess, CUDA-graph replay, all 4 input slots)
    vs ``oracle.resize.transform_u8`` + ``oracle.model.forward`` ON THE CPU (IEEE fp32): log-probabilities <= 1e-2,
    labels equal outside the 2e-2 margin band, counts == ``oracle.analysis.counts_table`` of the predictions."""
    from concurrent.futures import ThreadPoolExecutor
    from skin_image_analysis_b200.engine import EvalEngine
    from skin_image_analysis_b200.synthetic import random_state_dict
    batch = 256
    imgs = np.stack([helpers.synthetic_u8_image(450, 600, 2000 + i, "smooth" if i % 4 else "noise")
                     for i in range(batch)])
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        x_ref = torch.from_numpy(np.stack(list(pool.map(lambda im: R.transform_u8(im, (224, 224)), imgs))))
    state = random_state_dict(om.LIST_MODEL, 224, seed=0)
    ref = om.forward(om.LIST_MODEL, state, x_ref)
    last = [k for k in state if k.endswith(".bias")][-1]
    state[last][1] -= float((ref[:, 1] - ref[:, 0]).median())
    ref = om.forward(om.LIST_MODEL, state, x_ref)
    want_pred = ref.argmax(1).numpy()
    safe = ((ref[:, 1] - ref[:, 0]).abs() > 2 * LOGP_TOL).numpy()
    assert 0 < int(want_pred.sum()) < batch

    label, ftype, sex, control = helpers.counter_metadata(np.arange(batch), seed=9)
    eng = EvalEngine(state, batch, (450, 600), 224, n_slots=4)
    assert all(g is not None for g in eng.graphs)
    u8 = torch.from_numpy(imgs).pin_memory()
    lab, grp = torch.from_numpy(label).pin_memory(), torch.from_numpy(np.stack([ftype, sex, control])).pin_memory()
    for slot in range(4):
        eng.reset_counts()
        eng.step(u8, lab, grp, slot=slot)
        counts = eng.read_counts()
        logp, pred = eng.logp.cpu(), eng.pred.cpu().numpy()
        assert (logp - ref).abs().max().item() <= LOGP_TOL, slot
        assert np.array_equal(pred[safe], want_pred[safe]), slot
        inst = {i: {"benign_malignant": helpers.CLASS_NAMES[label[i]], "prediction": helpers.CLASS_NAMES[pred[i]],
                    "skin_type": helpers.FITZPATRICK[ftype[i]], "sex": ["male", "female"][sex[i]],
                    "control": ["rich", "poor"][control[i]]} for i in range(batch)}
        tab = oa.counts_table(inst, {"skin_type": helpers.FITZPATRICK, "sex": ["male", "female"],
                                     "control": ["rich", "poor"]})
        assert counts[0].tolist() == tab["skin_type"]
        assert counts[1, :2].tolist() == tab["sex"] and counts[2, :2].tolist() == tab["control"]
    # xavier-random weights give nearly constant logits, so the margin band holds most images; the
    # images outside it must still be a real sample
    assert int(safe.sum()) >= 32


def test_gpu_jpeg_decode_front_end(tmp_path):
    """nvJPEG decode (torchvision, a library call) -> sia_chw_u8_to_hwc_u8 -> fused transform.  The
    interleave kernel is exact; the decoded pixels differ from libjpeg's (the reference's decoder) by a few grey levels
    on smooth images, so after the 2x down-sampling resize the transform agrees with the oracle on the PIL-decoded bytes
    to 2 % of full scale -- a front end for throughput, not the bit-exact parity path (that is HibaDataset)."""
    import io
    from PIL import Image
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200.tone_bias_dataset import decode_jpeg_batch
    g = torch.Generator(device="cuda").manual_seed(2)
    planar = torch.randint(0, 256, (3, 3, 20, 28), dtype=torch.uint8, device="cuda", generator=g)
    assert torch.equal(ops.chw_to_hwc_u8(planar), planar.permute(0, 2, 3, 1).contiguous())
    files, pil = [], []
    for i in range(3):
        u8 = helpers.synthetic_u8_image(450, 600, 800 + i, "smooth")
        path = tmp_path / f"img{i}.jpg"
        Image.fromarray(u8, "RGB").save(path, format="JPEG", quality=95, subsampling=0)
        files.append(str(path) if i else path.read_bytes())                # a path or an encoded buffer
        pil.append(np.array(Image.open(path).convert("RGB")))
    batch = decode_jpeg_batch(files)
    assert batch.shape == (3, 450, 600, 3) and batch.dtype == torch.uint8 and batch.is_cuda
    diff = (batch.cpu().numpy().astype(np.int32) - np.stack(pil).astype(np.int32))
    assert np.abs(diff).mean() <= 1.0 and np.abs(diff).max() <= 16        # decoder-dependent, a few grey levels
    got = ops.preprocess_u8hwc(batch, (224, 224), ops.LAYOUT_NCHW_F32).cpu().numpy()
    want = np.stack([R.transform_u8(im, (224, 224)) for im in pil])
    assert np.abs(got - want).max() <= 0.02
