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

import os

import torch
import torch.nn as nn

from . import ops
from ._lib import SiaError

__all__ = ["SkinCancerListModel", "SkinCancerModel", "create_loss_function", "save_model", "create_model",
           "load_model", "CnnPlan"]


LEGACY_WIDTHS = ([32, 64, 128], [32, 64, 128, 256])     # SkinCancerListModel / SkinCancerModel: no channel padding


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def infer_image_size(n_blocks: int, c_last: int, in_features: int, prefer: int = 224) -> int:
    """Image size from the first Linear: in_features = c_last * side^2 with side = the image size floor-halved
    n_blocks times.  The reference hard-codes 224 (tone_bias_model.py:69-70, tone_bias_optuna.py:130): 224 wins
    whenever it is consistent; otherwise the smallest consistent size (side << n_blocks)."""
    side = int(round((in_features / c_last) ** 0.5))
    if side < 1 or c_last * side * side != in_features:
        raise SiaError(f"first Linear ({in_features} inputs) does not match {c_last} channels of a square image")
    if (prefer >> n_blocks) == side:
        return prefer
    return side << n_blocks


def fold_batchnorm(weight: torch.Tensor, bias: torch.Tensor, bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm2d after a convolution (the layer the reference keeps commented out between conv and ReLU,
    tone_bias_model.py:88) folded into the convolution: y = g * (conv(x) + b - mean) / sqrt(var + eps) + beta
    == conv'(x) + b' with w' = w * s, b' = (b - mean) * s + beta, s = g / sqrt(var + eps).  The kernels' fused
    bias + ReLU + pool epilogue then needs no extra pass."""
    if bn.running_mean is None or bn.running_var is None:
        raise SiaError("BatchNorm2d without running statistics cannot be evaluated in eval mode")
    s = (bn.running_var.detach().float() + bn.eps).rsqrt()
    beta = torch.zeros_like(s)
    if bn.affine:
        s = s * bn.weight.detach().float()
        beta = bn.bias.detach().float()
    w = weight.detach().float() * s.view(-1, 1, 1, 1)
    b = (bias.detach().float() - bn.running_mean.detach().float()) * s + beta
    return w, b


def conv_bn_fc_params(modules):
    """Leaf modules in forward order -> ([(conv weight, bias) with a directly following BatchNorm2d folded in],
    [(linear weight, bias)])."""
    convs, fcs = [], []
    mods = [m for m in modules if isinstance(m, (nn.Conv2d, nn.BatchNorm2d, nn.Linear))]
    for i, m in enumerate(mods):
        if isinstance(m, nn.Conv2d):
            bias = m.bias if m.bias is not None else torch.zeros(m.out_channels, device=m.weight.device)
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            convs.append(fold_batchnorm(m.weight, bias, nxt) if isinstance(nxt, nn.BatchNorm2d) else (m.weight, bias))
        elif isinstance(m, nn.Linear):
            fcs.append((m.weight, m.bias))
        elif i == 0 or not isinstance(mods[i - 1], nn.Conv2d):
            raise SiaError("BatchNorm2d is only supported directly after a convolution")
    return convs, fcs


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
            image_size = infer_image_size(len(conv_params), widths[-1], int(fc_params[0][0].shape[1]))
        self.image_size = image_size
        if image_size % 2 != 0:
            raise SiaError(f"image size {image_size} is odd: the first block pools 2x2 over an even-sized input")
        if len(fc_params) < 2 or fc_params[-1][0].shape[0] != 2:
            raise SiaError("the fused tail handles exactly two classes (benign / malignant) after >= 1 hidden Linear")
        legacy = widths in LEGACY_WIDTHS
        pads = widths if legacy else [_round_up(c, 64) for c in widths]
        if max(pads) > 256:
            raise SiaError("conv widths above 256 channels are not supported")
        self.pads = pads
        self.widths = widths
        # Spatial bookkeeping.  nn.MaxPool2d floors: valid[k] = valid[k-1] // 2 (224 -> 112 -> 56 -> 28 -> 14 -> 7 -> 3 -> 1
        # for the 7 blocks define_isic_model can ask for).  The conv kernels take even-sized inputs and write h/2 x w/2
        # outputs, so where the valid size turns odd (or the producing block already ran on a padded buffer and its last
        # output row / column is not a real pooled value) the activation is copied into an even-sized buffer that is
        # zero outside the valid corner (sia_pad_nhwc_bf16) -- the zero padding the next 'same' convolution expects.
        n_blocks = len(conv_params)
        self.valid = [image_size]
        for _ in range(n_blocks):
            self.valid.append(self.valid[-1] // 2)
        if self.valid[-1] < 1:
            raise SiaError(f"{n_blocks} pooling blocks reduce a {image_size}x{image_size} image to nothing")
        self.in_hw, self.raw_hw, self.needs_pad = [image_size], [], []
        for k in range(n_blocks):
            raw = self.in_hw[k] // 2
            self.raw_hw.append(raw)
            last = k == n_blocks - 1
            pad = (not last) and (raw != self.valid[k + 1] or raw % 2 != 0)
            self.needs_pad.append(pad)
            if not last:
                self.in_hw.append(_round_up(self.valid[k + 1], 2) if pad else raw)
        self.convs = []          # (packed | [packed chunks], bias | [bias chunks], cin_pad, cout_pad)
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
        c_last, c_last_pad = widths[-1], pads[-1]
        side, raw = self.valid[-1], self.raw_hw[-1]
        (w1, b1) = fc_params[0]
        if w1.shape[1] != c_last * side * side:
            raise SiaError("first Linear does not match the flattened conv output")
        w1 = w1.detach().float()
        if raw != side:           # the last block ran on a padded buffer: its extra output row / column gets zero weights
            w1p = torch.zeros((w1.shape[0], c_last, raw, raw), dtype=torch.float32, device=dev)
            w1p[:, :, :side, :side] = w1.view(-1, c_last, side, side)
            w1 = w1p.view(w1.shape[0], -1)
        # nn.Flatten on NCHW orders features (C,H,W); activations here are NHWC -> permute columns once
        self.n1 = int(w1.shape[0])
        self.n1_pad = _round_up(self.n1, 128)
        self.w1 = ops.pack_linear_chw_to_hwc(w1.contiguous(), c_last, raw * raw, self.n1_pad, c_last_pad)
        self.b1 = b1.detach().float().contiguous()
        self.feat = self.w1.shape[1]
        # tile-major copy for the GEMM (every 128 x 64 weight tile one contiguous 16 KB block); the row-major one is freed
        if os.environ.get("SIA_FC1_TILED", "1") != "0":          # A/B switch for timing
            self.w1 = ops.retile_linear_w(self.w1)
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
        # kernel launches of one forward pass: first block (one per 32 output channels) + 3x3 blocks + pad copies + fc1 + tail
        self.launches = len(self.convs[0][0]) + (n_blocks - 1) + sum(self.needs_pad) + 2
        self._ws = {}
        torch.cuda.current_stream(dev).synchronize()

    def splits_for(self, batch: int) -> int:
        tiles = ((batch + 127) // 128) * (self.n1_pad // 128)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        return max(1, min(self.feat // 64, sms // tiles))

    def workspace(self, batch: int):
        ws = self._ws.get(batch)
        if ws is None:
            acts, padded = [], []
            for k, (_p, _b, _cin, cout_pad) in enumerate(self.convs):
                # zero-initialised: padded channels are never written by the first block and must read as zero
                acts.append(torch.zeros((batch, self.raw_hw[k], self.raw_hw[k], cout_pad), dtype=torch.bfloat16,
                                        device=self.device))
                padded.append(torch.zeros((batch, self.in_hw[k + 1], self.in_hw[k + 1], cout_pad), dtype=torch.bfloat16,
                                          device=self.device) if self.needs_pad[k] else None)
            splits = self.splits_for(batch)
            ws = dict(acts=acts, padded=padded, splits=splits,
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
            if self.needs_pad[i]:
                v = self.valid[i + 1]
                h = ops.pad_nhwc(h, (v, v), (self.in_hw[i + 1],) * 2, out=ws["padded"][i])
        part = ops.linear_splitk(h.view(batch, -1), self.w1, ws["splits"], out=ws["partial"])
        return self.tail(part, label=label, groups=groups, n_groups=n_groups, counts=counts, logp=ws["logp"],
                         pred=ws["pred"])


class _B200Eval(nn.Module):
    """Shared forward of every architecture of the path."""

    image_size = 224            # tone_bias_model.py:69-70, tone_bias_optuna.py:130

    def _plan(self) -> CnnPlan:
        tensors = list(self.parameters()) + list(self.buffers())
        key = tuple((p.data_ptr(), p._version) for p in tensors)
        if getattr(self, "_plan_key", None) != key:
            convs, fcs = conv_bn_fc_params(self.modules())
            self._plan_obj = CnnPlan(convs, fcs, image_size=self.image_size)
            self._plan_key = key
        return self._plan_obj

    def predict(self, x):
        """(log-probabilities [B,2] f32, predicted class [B] u8) of one batch, both from the tail kernel: the label is
        the first maximal index, the tie rule of ``torch.max(outputs, 1)`` (tone_bias_test.py:199)."""
        if self.training:
            raise SiaError("this build implements the evaluation path only: call model.eval() first "
                           "(the reference does, tone_bias_test.py:175)")
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise SiaError("model input must be a CUDA tensor: the sm_100a path has no CPU fallback")
        with torch.no_grad(), torch.cuda.device(x.device):
            plan = self._plan()
            if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != plan.image_size or x.shape[3] != plan.image_size:
                raise ValueError(f"expected input [B,3,{plan.image_size},{plan.image_size}] (the reference hard-codes "
                                 "224, tone_bias_model.py:69-70)")
            x4 = ops.nchw_f32_to_nhwc4(x.float().contiguous())
            logp, pred = plan.forward_nhwc4(x4)
            return logp.clone(), pred.clone()

    def forward(self, x):
        return self.predict(x)[0]

    def get_class_names(self):
        return self.class_names


class SkinCancerListModel(_B200Eval):
    """3 conv blocks + 2 linear blocks + classifier + LogSoftmax (reference :56-152)."""


    def __init__(self, class_names, batch_norm: bool = False):
        super().__init__()
        self.class_names = class_names
        layers = []
        width = height = 224
        in_features = 3
        for i, out_features in enumerate([32, 64, 128]):
            conv = nn.Conv2d(in_features, out_features, kernel_size=7 if i == 0 else 3, stride=1, padding="same")
            nn.init.xavier_normal_(conv.weight)
            # batch_norm=True enables the layer the reference keeps commented out (:88); folded into the conv at run time
            layers += [conv] + ([nn.BatchNorm2d(out_features)] if batch_norm else [])
            layers += [nn.ReLU(), nn.MaxPool2d(kernel_size=(2, 2))]
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
    return _adopt(obj, class_names)


def _sequential_from_state_dict(state: dict):
    """A digit-keyed ``state_dict`` (an ``nn.Sequential`` such as tone_bias_optuna.define_isic_model's) -> a B200
    Sequential with the same keys: conv / linear modules are rebuilt from the tensor shapes, and the parameter-free
    layers between them (ReLU, MaxPool2d, Flatten, Dropout, LogSoftmax) only matter as index placeholders."""
    from .tone_bias_optuna import _B200Sequential
    names = sorted({k.rsplit(".", 1)[0] for k in state}, key=int)
    layers = [nn.Identity() for _ in range(int(names[-1]) + 2)]
    for n in names:
        w = state[n + ".weight"]
        if w.dim() == 4:
            layers[int(n)] = nn.Conv2d(w.shape[1], w.shape[0], kernel_size=w.shape[2], stride=1, padding="same")
        elif w.dim() == 2:
            layers[int(n)] = nn.Linear(w.shape[1], w.shape[0])
        elif w.dim() == 1 and n + ".running_mean" in state:
            layers[int(n)] = nn.BatchNorm2d(w.shape[0])
        else:
            raise TypeError(f"cannot rebuild layer {n} from a parameter of shape {tuple(w.shape)}")
    model = _B200Sequential(*layers)
    model.load_state_dict(state)
    return model


def _adopt(obj, class_names):
    """Whatever ``torch.load`` produced -> a module of this package holding the same parameters."""
    if isinstance(obj, dict):
        if all(k.split(".")[0].isdigit() for k in obj):
            return _sequential_from_state_dict(obj)
        kind = SkinCancerListModel if any(k.startswith("layers.") for k in obj) else SkinCancerModel
        model = kind(class_names)
        model.load_state_dict(obj)
        return model
    if isinstance(obj, _B200Eval):
        return obj
    if isinstance(obj, nn.Sequential):      # e.g. a pickled define_isic_model() of the reference
        from .tone_bias_optuna import _B200Sequential
        return _B200Sequential.from_modules(obj.children())
    if isinstance(obj, nn.Module):          # a genuine reference object (reference module importable)
        kind = SkinCancerListModel if hasattr(obj, "layers") else SkinCancerModel
        model = kind(class_names)
        model.load_state_dict(obj.state_dict())
        return model
    raise TypeError(f"cannot interpret {type(obj)} as a model")
