"""Drop-in surface for the ToneClassifier test transform.

Mirrors notebooks/ToneClassifier/CNNTrialDataset.py of the reference:

  * ``fitzpatrick_converter``  (:11-25)  Fitzpatrick type -> 0 (light: I, II) / 1 (dark: III..VI) / "Error"
  * ``ISIC.transforms`` for ``state != "Train"`` (:71-76):
        v2.Compose([v2.Resize((224, 224)), v2.ToDtype(torch.float32, scale=True),
                    v2.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    applied in ``__getitem__`` (:93-95) to the uint8 CHW tensor ``torchvision.io.read_image`` returns.

``TestTransforms`` is that Compose as ONE sm_100a kernel (``sia_preprocess_tv_u8hwc``, csrc/preprocess_tv.cu),
bit-exact with torchvision (integer resampler + a table of the float32 normalisation).  It takes what the
reference's transform takes -- a uint8 ``[3, H, W]`` tensor -- or a whole batch ``[B, 3, H, W]`` of decoded
images already resident on the GPU, and returns float32 ``[3, 224, 224]`` / ``[B, 3, 224, 224]``.  CUDA tensors
only: there is no CPU fallback (use torchvision itself on the CPU).  The random training augmentations
(:53-58, RandomHorizontalFlip / RandomCrop) and JPEG decoding (:93) are outside the evaluation path.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["fitzpatrick_converter", "TestTransforms", "IMAGENET_MEAN", "IMAGENET_STD"]

IMAGENET_MEAN = ops.IMAGENET_MEAN
IMAGENET_STD = ops.IMAGENET_STD


def fitzpatrick_converter(entry):
    """CNNTrialDataset.py:11-25 -- light (0) for types I / II, dark (1) for III..VI, the string "Error" otherwise."""
    if entry in ("I", "II"):
        return 0
    if entry in ("III", "IV", "V", "VI"):
        return 1
    return "Error"


class TestTransforms:
    """``ISIC(image_path, "Test").transforms`` (CNNTrialDataset.py:71-76) for GPU-resident decode buffers."""
    __test__ = False                     # not a pytest class

    def __init__(self, size=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD, layout: int = ops.LAYOUT_NCHW_F32):
        self.size = (int(size[0]), int(size[1]))
        self.mean, self.std, self.layout = tuple(mean), tuple(std), layout

    def __call__(self, image: torch.Tensor) -> torch.Tensor:
        if image.dtype != torch.uint8:
            raise TypeError("expected the uint8 tensor torchvision.io.read_image returns")
        if image.dim() == 3:
            return ops.preprocess_tv_u8hwc(image.unsqueeze(0), self.size, self.layout, self.mean, self.std,
                                           planar=True)[0]
        if image.dim() == 4:
            return ops.preprocess_tv_u8hwc(image, self.size, self.layout, self.mean, self.std, planar=True)
        raise ValueError("expected [3,H,W] or [B,3,H,W]")

    def hwc(self, images: torch.Tensor) -> torch.Tensor:
        """Same transform for interleaved decode buffers ``[B, H, W, 3]`` (the layout of the main evaluation path)."""
        return ops.preprocess_tv_u8hwc(images, self.size, self.layout, self.mean, self.std, planar=False)
