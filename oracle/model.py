"""Oracle: the tone_bias CNN forward, restated as plain fp32 torch functional ops
(TEST INFRASTRUCTURE ONLY -- never imported by the product package).

Follows
  * ``SkinCancerListModel``  tone_bias_model.py:56-152  (3 conv blocks, keys ``layers.{0,3,6,10,13,16}``)
  * ``SkinCancerModel``      tone_bias_model.py:155-299 (4 conv blocks, keys ``conv1..4, fc4..6``)
    == jgi_hiba_2022_model.py:155-299 (byte-identical file)

Each conv block is Conv2d(stride 1, padding='same') -> ReLU -> MaxPool2d(2,2) (:83-92, :169-184),
then Flatten over (C,H,W) (:100), two Linear+ReLU (+Dropout, identity in eval, :111-115) and
Linear -> LogSoftmax(dim=1) (:126-129).  Prediction = ``torch.max(outputs, 1)`` i.e. first
maximal index on ties (tone_bias_test.py:199).

Weights are passed as a ``state_dict`` with the reference's key names, so the very same
tensors can be loaded into the reference classes (tests/golden/make_golden.py does that).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LIST_MODEL = "SkinCancerListModel"
FOUR_CONV_MODEL = "SkinCancerModel"

# (state_dict prefix, out_channels, in_channels, kernel) per conv block; then linear prefixes
_ARCH = {
    LIST_MODEL: dict(
        convs=[("layers.0", 32, 3, 7), ("layers.3", 64, 32, 3), ("layers.6", 128, 64, 3)],
        linears=[("layers.10", 512), ("layers.13", 256), ("layers.16", None)],
    ),
    FOUR_CONV_MODEL: dict(
        convs=[("conv1", 32, 3, 7), ("conv2", 64, 32, 3), ("conv3", 128, 64, 3), ("conv4", 256, 128, 3)],
        linears=[("fc4", 512), ("fc5", 256), ("fc6", None)],
    ),
}


def arch(kind: str):
    return _ARCH[kind]


def param_shapes(kind: str, image_size: int = 224, num_classes: int = 2) -> dict[str, tuple[int, ...]]:
    """Name -> shape of every parameter, in the reference's state_dict order."""
    a = _ARCH[kind]
    shapes: dict[str, tuple[int, ...]] = {}
    side = image_size
    ch = 3
    for prefix, cout, cin, k in a["convs"]:
        shapes[prefix + ".weight"] = (cout, cin, k, k)
        shapes[prefix + ".bias"] = (cout,)
        side //= 2
        ch = cout
    feat = ch * side * side
    for prefix, width in a["linears"]:
        width = num_classes if width is None else width
        shapes[prefix + ".weight"] = (width, feat)
        shapes[prefix + ".bias"] = (width,)
        feat = width
    return shapes


def synthetic_state_dict(kind: str, seed: int, image_size: int = 224, num_classes: int = 2,
                         centre_head: bool = True) -> dict[str, torch.Tensor]:
    """Deterministic random-init weights with the reference's init *distributions*
    (xavier_normal_ on weights, tone_bias_model.py:136-137; torch-default uniform biases),
    generated per tensor from ``seed`` so they can be rebuilt anywhere without shipping 200 MB.

    ``centre_head``: an un-centred random-init head predicts one class for every image
    (SURVEY section 6); zeroing the last bias keeps both heads comparable -- the per-run
    centring by the median margin is done by the callers (bench / tests) on both sides alike.
    """
    out: dict[str, torch.Tensor] = {}
    for i, (name, shape) in enumerate(param_shapes(kind, image_size, num_classes).items()):
        g = torch.Generator().manual_seed(seed * 1000 + i)
        if name.endswith(".weight"):
            recept = math.prod(shape[2:]) if len(shape) > 2 else 1
            fan_in, fan_out = shape[1] * recept, shape[0] * recept
            std = math.sqrt(2.0 / (fan_in + fan_out))
            out[name] = torch.randn(shape, generator=g, dtype=torch.float32) * std
        else:
            w_shape = param_shapes(kind, image_size, num_classes)[name[:-5] + ".weight"]
            fan_in = math.prod(w_shape[1:])
            bound = 1.0 / math.sqrt(fan_in)
            out[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    if centre_head:
        last = list(out)[-1]
        out[last] = torch.zeros_like(out[last])
    return out


@torch.no_grad()
def forward(kind: str, state: dict[str, torch.Tensor], x: torch.Tensor,
            return_intermediates: bool = False):
    """fp32 NCHW forward -> [B, num_classes] log-probabilities."""
    a = _ARCH[kind]
    inter = {}
    h = x.to(torch.float32)
    for prefix, _cout, _cin, _k in a["convs"]:
        h = F.conv2d(h, state[prefix + ".weight"].float(), state[prefix + ".bias"].float(),
                     stride=1, padding="same")
        h = F.max_pool2d(F.relu(h), kernel_size=(2, 2))
        inter[prefix] = h
    h = torch.flatten(h, 1)
    n_lin = len(a["linears"])
    for j, (prefix, _w) in enumerate(a["linears"]):
        h = F.linear(h, state[prefix + ".weight"].float(), state[prefix + ".bias"].float())
        if j < n_lin - 1:
            h = F.relu(h)           # Dropout(0.5) is the identity in eval mode
        inter[prefix] = h
    out = F.log_softmax(h, dim=1)
    if return_intermediates:
        return out, inter
    return out


@torch.no_grad()
def forward_sequential(state: dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """fp32 forward of the ``nn.Sequential`` built by ``define_isic_model`` (tone_bias_optuna.py:123-173) from its
    state_dict (keys ``<index>.weight`` / ``<index>.bias`` in layer order): every 4-D weight is a
    Conv2d(stride 1, padding='same') + ReLU + MaxPool2d(2,2) block, then Flatten, every 2-D weight a Linear with
    ReLU (+ Dropout = identity in eval) except the last, then LogSoftmax(dim=1)."""
    names = sorted({k.rsplit(".", 1)[0] for k in state}, key=lambda n: int(n))
    h = x.to(torch.float32)
    linears = [n for n in names if state[n + ".weight"].dim() == 2]
    for n in names:
        w, b = state[n + ".weight"].float(), state[n + ".bias"].float()
        if w.dim() == 4:
            h = F.max_pool2d(F.relu(F.conv2d(h, w, b, stride=1, padding="same")), kernel_size=(2, 2))
        else:
            if h.dim() > 2:
                h = torch.flatten(h, 1)
            h = F.linear(h, w, b)
            if n != linears[-1]:
                h = F.relu(h)
    return F.log_softmax(h, dim=1)


def logits_from_logprobs_margin(logp: torch.Tensor) -> torch.Tensor:
    """l1 - l0 (the decision margin is invariant under log-softmax)."""
    return logp[:, 1] - logp[:, 0]


def predict(logp: torch.Tensor) -> torch.Tensor:
    """tone_bias_test.py:199 -- ``torch.max(outputs.data, 1)`` indices (first max on ties)."""
    return torch.max(logp, 1)[1]


def flops_per_image(kind: str, image_size: int = 224, num_classes: int = 2) -> int:
    """2 * MACs, unpadded."""
    a = _ARCH[kind]
    side, macs = image_size, 0
    for _p, cout, cin, k in a["convs"]:
        macs += side * side * cout * cin * k * k
        side //= 2
    for name, shape in param_shapes(kind, image_size, num_classes).items():
        if name.endswith(".weight") and len(shape) == 2:
            macs += shape[0] * shape[1]
    return 2 * macs
