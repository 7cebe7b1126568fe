"""Oracle (oracle/model.py) vs fixtures produced by the reference's own nn.Module classes -- CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import model as om
from tests import helpers


@pytest.mark.parametrize("kind", [om.LIST_MODEL, om.FOUR_CONV_MODEL])
def test_forward_matches_reference_fixture(golden_dir, kind):
    g = np.load(os.path.join(golden_dir, f"model_{kind}.npz"))
    state = om.synthetic_state_dict(kind, seed=7)
    x = helpers.synthetic_batch_f32(4, 224, seed=21)
    logp, inter = om.forward(kind, state, x, return_intermediates=True)
    np.testing.assert_allclose(logp.numpy(), g["logp"], rtol=0, atol=2e-5)
    assert np.array_equal(om.predict(logp).numpy(), g["pred"])
    conv_names = [c[0] for c in om.arch(kind)["convs"]]
    gold_feats = sorted((k for k in g.files if k.endswith("_slice")),
                        key=lambda s: int("".join(ch for ch in s if ch.isdigit()) or 0))
    for name, gk in zip(conv_names, gold_feats[:len(conv_names)]):
        v = inter[name].numpy()
        sl = v.reshape(v.shape[0], -1)[:, ::max(1, v[0].size // 64)][:, :64]
        np.testing.assert_allclose(sl, g[gk], rtol=1e-4, atol=1e-5)


def test_param_counts_match_survey():
    n = sum(int(np.prod(s)) for s in om.param_shapes(om.LIST_MODEL).values())
    assert n == 51_609_666
    n = sum(int(np.prod(s)) for s in om.param_shapes(om.FOUR_CONV_MODEL).values())
    assert n == 26_214_722                           # row a5
    assert om.flops_per_image(om.LIST_MODEL) == 1_499_914_240 or abs(om.flops_per_image(om.LIST_MODEL) - 1.4999e9) < 1e6
    assert abs(om.flops_per_image(om.FOUR_CONV_MODEL) - 1.9110e9) < 1e6


def test_logsoftmax_rows_normalised():
    state = om.synthetic_state_dict(om.LIST_MODEL, seed=1)
    x = helpers.synthetic_batch_f32(2, 224, seed=2)
    logp = om.forward(om.LIST_MODEL, state, x)
    assert torch.allclose(logp.exp().sum(1), torch.ones(2), atol=1e-6)


def test_optuna_best_model_builder_reproduces_reference_weights_and_oracle(golden_dir):
    """tone_bias_optuna drop-in: same layers in the same order as the reference builder, so the same torch seed
    gives the same parameters (checked by per-tensor sums recorded from the reference), and the oracle's
    sequential forward of those weights reproduces the reference's log-probabilities."""
    import contextlib
    import io
    import os

    import numpy as np
    import torch

    from oracle import model as om
    from skin_image_analysis_b200 import tone_bias_optuna as to
    from tests import helpers
    g = np.load(os.path.join(golden_dir, "model_optuna_best.npz"))
    torch.manual_seed(123)
    with contextlib.redirect_stdout(io.StringIO()) as out:
        model = to.create_best_model()
    assert "DEBUGGING 16856 = 86 * (14*14)" in out.getvalue()
    state = model.state_dict()
    assert list(state.keys()) == list(g["keys"])
    for k, v in state.items():
        assert float(v.double().sum()) == float(g[k.replace(".", "_") + "_sum"]), k
    x = helpers.synthetic_batch_f32(3, 224, seed=33)
    logp = om.forward_sequential(state, x).numpy()
    assert np.abs(logp - g["logp"]).max() < 1e-5
    trial = to.TrialDummy({"a": 3, "b": 0.5})
    assert trial.suggest_int("a", 1, 6) == 3 and trial.suggest_float("b", 0.2, 0.5) == 0.5
    with pytest.raises(ValueError):
        trial.suggest_int("a", 4, 6)                      # below the minimum
    assert trial.suggest_int("a", 1, 2) == 3              # above the maximum: not enforced by the reference either
