"""Drop-in for ``tone_bias_model`` (reference src/tone_bias_model.py) -- evaluation path only.

``SkinCancerListModel`` (reference :56-152) and ``SkinCancerModel`` (:155-299) keep the reference's
constructor, sub-module names, ``state_dict()`` keys / shapes (``layers.{0,3,6,10,13,16}.*`` and
``conv1..4 / fc4..6``), ``get_class_names()`` and the ``[B,2]`` log-probability output, so
``load_state_dict`` from a reference model and ``predict_with_instance`` work unchanged.  The
``nn.Conv2d`` / ``nn.Linear`` children only HOLD the parameters: ``forward`` runs the hand-written
sm_100a kernels (csrc/conv1.cu, conv3x3.cu, linear.cu) on bf16 copies of the weights with fp32
accumulation.  There is no CPU or eager fallback: a CPU tensor, training mode, or a missing
``libsia_b200.so`` raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import SiaError

__all__ = ["SkinCancerListModel", "SkinCancerModel", "create_loss_function", "save_model", "create_model",
           "load_model", "CnnPlan"]


LEGACY_WIDTHS = ([32, 64, 128], [32, 64, 128, 256])     # SkinCancerListModel / SkinCancerModel: no channel padding


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class CnnPlan:
    """Packed bf16 weights + the launch sequence for one architecture on one device.

    conv_params: [(weight[O,I,k,k], bias[O])...] -- a 7x7 block on 3 channels, then 3x3 blocks;
    fc_params: [(weight, bias), ...] -- two or more Linear layers, the last one with 2 outputs (reference layout).
    The two fixed architectures of tone_bias_model.py use their exact widths; any other widths (16..256,
    tone_bias_optuna.define_isic_model) run through the same kernels on channel buffers zero-padded to multiples
    of 64 (padded weights and biases are zero, so padded activations are exactly zero).
    """

    def __init__(self, conv_params, fc_params, image_size: int | None = None):
        dev = conv_params[0][0].device
        if dev.type != "cuda":
            raise SiaError("CnnPlan needs CUDA parameters (there is no CPU fallback)")
        self.device = dev
        widths = [int(w.shape[0]) for w, _ in conv_params]
        if image_size is None:
            # the reference hard-codes 224 (tone_bias_model.py:69-70); other sizes follow from the first Linear:
            # in_features = c_last * (image_size / 2^n_conv)^2
            side = int(round((fc_params[0][0].shape[1] / widths[-1]) ** 0.5))
            image_size = side << len(conv_params)
        self.image_size = image_size
        if len(fc_params) < 2 or fc_params[-1][0].shape[0] != 2:
            raise SiaError("the fused tail handles exactly two classes (benign / malignant) after >= 1 hidden Linear")
        legacy = widths in LEGACY_WIDTHS
        pads = widths if legacy else [_round_up(c, 64) for c in widths]
        if max(pads) > 256:
            raise SiaError("conv widths above 256 channels are not supported")
        self.pads = pads
        self.widths = widths
        self.convs = []          # (packed | [packed chunks], bias | [bias chunks], cin_pad, cout_pad)
        side = image_size
        for i, (w, b) in enumerate(conv_params):
            w = w.detach().float().contiguous()
            b = b.detach().float().contiguous()
            cout, cin, k, _ = w.shape
            if i == 0:
                if (cin, k) != (3, 7):
                    raise SiaError(f"first block must be Conv2d(3, n, 7); got ({cin},{cout},{k})")
                chunks, biases = [], []
                for c0 in range(0, cout, 32):      # the kernel computes 32 output channels per launch
                    wc = torch.zeros((32, 3, 7, 7), dtype=torch.float32, device=dev)
                    bc = torch.zeros((32,), dtype=torch.float32, device=dev)
                    n = min(32, cout - c0)
                    wc[:n], bc[:n] = w[c0:c0 + n], b[c0:c0 + n]
                    chunks.append(ops.pack_conv7x7_c3(wc))
                    biases.append(bc)
                self.convs.append((chunks, biases, 3, pads[0]))
            else:
                if k != 3 or cin != widths[i - 1]:
                    raise SiaError("only chained 3x3 kernels after the first block")
                bp = torch.zeros((pads[i],), dtype=torch.float32, device=dev)
                bp[:cout] = b
                self.convs.append((ops.pack_conv3x3(w, pads[i - 1], pads[i]), bp, pads[i - 1], pads[i]))
            side //= 2
        c_last, c_last_pad = widths[-1], pads[-1]
        (w1, b1) = fc_params[0]
        if w1.shape[1] != c_last * side * side:
            raise SiaError("first Linear does not match the flattened conv output")
        # nn.Flatten on NCHW orders features (C,H,W); activations here are NHWC -> permute columns once
        self.n1 = int(w1.shape[0])
        self.n1_pad = _round_up(self.n1, 128)
        self.w1 = ops.pack_linear_chw_to_hwc(w1.detach().float().contiguous(), c_last, side * side, self.n1_pad, c_last_pad)
        self.b1 = b1.detach().float().contiguous()
        self.feat = self.w1.shape[1]
        rest = [(w.detach().float(), b.detach().float().contiguous()) for w, b in fc_params[1:]]
        self.fused_tail = (len(rest) == 2 and self.n1 == self.n1_pad and self.n1 <= 512 and rest[0][0].shape[0] <= 256)
        if self.fused_tail:
            (w2, self.b2), (w3, self.b3) = rest
            self.w2t = w2.t().contiguous()
            self.w3 = w3.contiguous()
            self.n2 = self.w2t.shape[1]
        else:
            if self.n1 > 512 or any(w.shape[0] > 512 for w, _ in rest):
                raise SiaError("Linear layers wider than 512 are not supported by the tail kernel")
            self.chain = [(w.t().contiguous(), b) for w, b in rest]
        self._ws = {}
        torch.cuda.current_stream(dev).synchronize()

    def splits_for(self, batch: int) -> int:
        tiles = ((batch + 127) // 128) * (self.n1_pad // 128)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        return max(1, min(self.feat // 64, sms // tiles))

    def workspace(self, batch: int):
        ws = self._ws.get(batch)
        if ws is None:
            s = self.image_size
            acts = []
            for (_p, _b, _cin, cout_pad) in self.convs:
                s //= 2
                # zero-initialised: padded channels are never written by the first block and must read as zero
                acts.append(torch.zeros((batch, s, s, cout_pad), dtype=torch.bfloat16, device=self.device))
            splits = self.splits_for(batch)
            ws = dict(acts=acts, splits=splits,
                      partial=torch.empty((splits, batch, self.n1_pad), dtype=torch.float32, device=self.device),
                      logp=torch.empty((batch, 2), dtype=torch.float32, device=self.device),
                      pred=torch.empty((batch,), dtype=torch.uint8, device=self.device))
            self._ws[batch] = ws
        return ws

    def conv_block(self, i: int, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        packed, bias, _cin_pad, cout_pad = self.convs[i]
        if i == 0:
            for j, (pk, bs) in enumerate(zip(packed, bias)):
                ops.conv7x7_c3_relu_pool2(x, pk, bs, out=out, c_offset=32 * j)
            return out
        return ops.conv3x3_relu_pool2(x, packed, bias, cout_pad, out=out)

    def tail(self, part, label=None, groups=None, n_groups: int = 0, counts=None, logp=None, pred=None):
        if self.fused_tail:
            return ops.head_tail(part, self.b1, self.w2t, self.b2, self.w3, self.b3, label=label, groups=groups,
                                 n_groups=n_groups, counts=counts, logp=logp, pred=pred)
        return ops.head_tail_chain(part, self.n1, self.b1, self.chain, label=label, groups=groups, n_groups=n_groups,
                                   counts=counts, logp=logp, pred=pred)

    def forward_nhwc4(self, x4: torch.Tensor, label=None, groups=None, n_groups: int = 0, counts=None):
        """x4: padded NHWC4 [B,S,S+8,4] bf16 -> (logp [B,2] f32, pred [B] u8).  Buffers are reused per batch size."""
        batch = x4.shape[0]
        ws = self.workspace(batch)
        h = x4
        for i in range(len(self.convs)):
            h = self.conv_block(i, h, ws["acts"][i])
        part = ops.linear_splitk(h.view(batch, -1), self.w1, ws["splits"], out=ws["partial"])
        return self.tail(part, label=label, groups=groups, n_groups=n_groups, counts=counts, logp=ws["logp"],
                         pred=ws["pred"])


class _B200Eval(nn.Module):
    """Shared forward of both architectures."""

    def _conv_fc(self):
        raise NotImplementedError

    def _plan(self) -> CnnPlan:
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, "_plan_key", None) != key:
            convs, fcs = self._conv_fc()
            self._plan_obj = CnnPlan([(m.weight, m.bias) for m in convs], [(m.weight, m.bias) for m in fcs])
            self._plan_key = key
        return self._plan_obj

    def forward(self, x):
        if self.training:
            raise SiaError("this build implements the evaluation path only: call model.eval() first "
                           "(the reference does, tone_bias_test.py:175)")
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise SiaError("model input must be a CUDA tensor: the sm_100a path has no CPU fallback")
        with torch.no_grad():
            plan = self._plan()
            if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != plan.image_size or x.shape[3] != plan.image_size:
                raise ValueError(f"expected input [B,3,{plan.image_size},{plan.image_size}] (the reference hard-codes "
                                 "224, tone_bias_model.py:69-70; the size follows from the first Linear)")
            x4 = ops.nchw_f32_to_nhwc4(x.float().contiguous())
            logp, _pred = plan.forward_nhwc4(x4)
            return logp.clone()

    def get_class_names(self):
        return self.class_names


class SkinCancerListModel(_B200Eval):
    """3 conv blocks + 2 linear blocks + classifier + LogSoftmax (reference :56-152)."""

    def __init__(self, class_names):
        super().__init__()
        self.class_names = class_names
        layers = []
        width = height = 224
        in_features = 3
        for i, out_features in enumerate([32, 64, 128]):
            conv = nn.Conv2d(in_features, out_features, kernel_size=7 if i == 0 else 3, stride=1, padding="same")
            nn.init.xavier_normal_(conv.weight)
            layers += [conv, nn.ReLU(), nn.MaxPool2d(kernel_size=(2, 2))]
            width, height = width // 2, height // 2
            in_features = out_features
        layers.append(nn.Flatten())
        in_features = in_features * width * height
        for out_features in [512, 256]:
            lin = nn.Linear(in_features, out_features)
            nn.init.xavier_normal_(lin.weight)
            layers += [lin, nn.ReLU(), nn.Dropout(0.5)]
            in_features = out_features
        head = nn.Linear(in_features, len(class_names))
        nn.init.xavier_normal_(head.weight)
        layers += [head, nn.LogSoftmax(dim=1)]
        self.layers = nn.Sequential(*layers)

    def _conv_fc(self):
        L = self.layers
        return [L[0], L[3], L[6]], [L[10], L[13], L[16]]


class SkinCancerModel(_B200Eval):
    """4 conv blocks (reference :155-299; identical in jgi_hiba_2022_model.py)."""

    def __init__(self, class_names):
        super().__init__()
        self.class_names = class_names
        self.conv1 = nn.Conv2d(3, 32, kernel_size=7, stride=1, padding="same")
        self.act1, self.pool1 = nn.ReLU(), nn.MaxPool2d(kernel_size=(2, 2))
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, stride=1, padding="same")
        self.act2, self.pool2 = nn.ReLU(), nn.MaxPool2d(kernel_size=(2, 2))
        self.conv3 = nn.Conv2d(64, 128, kernel_size=3, stride=1, padding="same")
        self.act3, self.pool3 = nn.ReLU(), nn.MaxPool2d(kernel_size=(2, 2))
        self.conv4 = nn.Conv2d(128, 256, kernel_size=3, stride=1, padding="same")
        self.act4, self.pool4 = nn.ReLU(), nn.MaxPool2d(kernel_size=(2, 2))
        self.flat = nn.Flatten()
        self.fc4 = nn.Linear(256 * 14 * 14, 512)
        self.drop4 = nn.Dropout(0.5)
        self.fc5 = nn.Linear(512, 256)
        self.act5, self.drop5 = nn.ReLU(), nn.Dropout(0.5)
        self.fc6 = nn.Linear(256, len(class_names))
        self.logsoftmax = nn.LogSoftmax(dim=1)
        for m in (self.conv1, self.conv2, self.conv3, self.conv4, self.fc4, self.fc5, self.fc6):
            nn.init.xavier_normal_(m.weight)

    def _conv_fc(self):
        return [self.conv1, self.conv2, self.conv3, self.conv4], [self.fc4, self.fc5, self.fc6]


def create_loss_function():
    return nn.NLLLoss()


def save_model(model, model_path):
    """Whole-module pickle, as the reference does (:305-315)."""
    torch.save(model, model_path)


def create_model(class_names):
    return SkinCancerModel(class_names)


def load_model(model_path, class_names):
    """Loads a ``session_model.pth``.  Accepts (a) a pickle written by ``save_model`` above, (b) a
    bare ``state_dict``, (c) the reference's own whole-module pickle -- its classes are resolved to
    the classes of this module (``tone_bias_model.<Class>`` / ``jgi_hiba_2022_model.<Class>``) by
    registering this module under those names for the duration of the load."""
    import sys
    me = sys.modules[__name__]
    saved = {k: sys.modules.get(k) for k in ("tone_bias_model", "jgi_hiba_2022_model")}
    try:
        for k, v in saved.items():
            if v is None:
                sys.modules[k] = me
        obj = torch.load(model_path, weights_only=False, map_location="cpu")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
    if isinstance(obj, dict):
        kind = SkinCancerListModel if any(k.startswith("layers.") for k in obj) else SkinCancerModel
        model = kind(class_names)
        model.load_state_dict(obj)
        return model
    if isinstance(obj, _B200Eval):
        return obj
    if isinstance(obj, nn.Module):      # a genuine reference object (reference module importable)
        kind = SkinCancerListModel if hasattr(obj, "layers") else SkinCancerModel
        model = kind(class_names)
        model.load_state_dict(obj.state_dict())
        return model
    raise TypeError(f"cannot interpret {type(obj)} as a model")
