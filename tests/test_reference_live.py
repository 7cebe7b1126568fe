"""Live cross-check of the oracle against the UNMODIFIED reference modules (oracle/ref_import.py).

Runs only where /root/reference exists (the build container); on the GPU box every test here skips -- nothing under
``-m gpu`` depends on it.  It repeats, on fresh seeds, what tests/golden/make_golden.py pins with fixtures: the oracle
forward == the reference classes' forward, the oracle analysis == the reference's ``analyse_predictions`` (result dict
and printed lines), the reference's own ``Rescale`` + ``ToTensor`` glue == ``oracle.resize.transform_u8`` (with
``skimage.transform.resize`` bound to the scipy restatement: scikit-image itself is not installable offline)."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import analysis as oa
from oracle import model as om
from oracle import ref_import
from oracle import resize as R
from tests import helpers

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference is not present on this box")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load()


@pytest.mark.parametrize("kind", [om.LIST_MODEL, om.FOUR_CONV_MODEL])
def test_oracle_forward_equals_the_reference_class(ref, kind):
    cls = ref.model.SkinCancerListModel if kind == om.LIST_MODEL else ref.hiba.SkinCancerModel
    state = om.synthetic_state_dict(kind, seed=101)
    m = cls(helpers.CLASS_NAMES).eval()
    m.load_state_dict(state, strict=True)
    x = helpers.synthetic_batch_f32(2, 224, seed=102)
    with torch.no_grad():
        want = m(x)
    got = om.forward(kind, state, x)
    assert float((got - want).abs().max()) <= 1e-6
    assert torch.equal(om.predict(got), torch.max(want, 1)[1])


def test_oracle_analysis_equals_the_reference(ref):
    inst = helpers.synthetic_instances(257, seed=103, with_oddities=True)
    buf_ref, buf_got = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(buf_ref):
        want = ref.test.analyse_predictions(inst)
    got = oa.analyse_predictions(inst, out=lambda *a: print(*a, file=buf_got))
    assert got == want and buf_got.getvalue() == buf_ref.getvalue()
    want_cells = [len(d) for d in ref.test.confusion_matrix(inst)]
    assert [len(d) for d in oa.confusion_matrix(inst)] == want_cells and sum(want_cells) == len(inst)


def test_reference_rescale_glue_equals_the_oracle_transform(ref):
    u8 = helpers.synthetic_u8_image(90, 120, 104, "smooth")
    sample = (np.float32(u8) / 255.0, 1, 7)                       # tone_bias_dataset.py:335
    sample = ref.dataset.Rescale((48, 64))(sample)
    t, label, idx = ref.dataset.ToTensor()(sample)
    assert (label, idx) == (1, 7)
    want = R.transform_u8(u8, (48, 64), R.resize_scipy)
    assert np.abs(t.numpy() - want).max() <= 1e-6
    # and the scipy route agrees with the plain-numpy restatement of the same algorithm
    assert np.abs(R.transform_u8(u8, (48, 64)) - want).max() <= 1e-6
