"""Drop-in for the transform part of ``tone_bias_dataset`` (reference src/tone_bias_dataset.py).

Reference surface kept: ``HibaDataset`` (:258-393), ``Rescale`` (:397-427), ``RandomCrop`` (:430-458),
``ToTensor`` (:461-473), ``convert_type2tone`` (:84-98) -- callables on the sample tuple
``(image HWC float32 in [0,1], label, index)``.

The arithmetic of the chain  u8 -> /255 (:335) -> skimage resize (:425) -> CHW (:470)  runs in the fused
``sia_preprocess_u8hwc`` kernel.  ``Rescale`` drives it one image at a time (API compatibility);
``GpuBatchTransform`` is the batched entry the evaluation engine uses: [B,H,W,3] uint8 on the device
in, the whole normalised batch out, one launch.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ops
from ._lib import SiaError
from .resize_weights import rescale_size

__all__ = ["HibaDataset", "Rescale", "RandomCrop", "ToTensor", "GpuBatchTransform", "convert_type2tone"]


def convert_type2tone(row):
    """Fitzpatrick {I, II} -> 'light', anything else -> 'dark' (reference :84-98)."""
    return "light" if row["fitzpatrick_skin_type"] in ("I", "II") else "dark"


def _unit_float_to_u8(image: np.ndarray) -> np.ndarray:
    """Recovers the uint8 decode buffer from ``np.float32(u8)/255.0`` exactly, or raises."""
    if image.dtype == np.uint8:
        return image
    u8 = np.rint(image.astype(np.float64) * 255.0)
    if u8.min() < 0 or u8.max() > 255:
        raise SiaError("Rescale: image values outside [0,1]; only uint8-derived images are supported")
    u8 = u8.astype(np.uint8)
    if not np.array_equal(np.float32(u8) / 255.0, image.astype(np.float32)):
        raise SiaError("Rescale: image is not float32(uint8)/255 -- the CUDA transform starts from the "
                       "uint8 decode buffer (tone_bias_dataset.py:326-335) and has no float-input variant")
    return u8


class GpuBatchTransform:
    """[B,H,W,3] uint8 CUDA tensor -> resized, scaled, normalised batch (one kernel launch).

    size: (h, w).  mean / std default to the reference's "no normalisation" (values stay in [0,1]);
    ImageNet mean/std gives the notebooks/ToneClassifier variant.  layout: 'nchw_f32' (what the reference
    DataLoader yields), 'nchw_bf16', or 'nhwc4_bf16' (what the first conv kernel consumes).
    """

    _LAYOUTS = {"nchw_f32": ops.LAYOUT_NCHW_F32, "nchw_bf16": ops.LAYOUT_NCHW_BF16,
                "nhwc4_bf16": ops.LAYOUT_NHWC4_BF16}

    def __init__(self, size=(224, 224), mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0), scale=1.0 / 255.0,
                 antialias="skimage", layout="nchw_f32", rows_per_cta=32):
        self.size = (int(size[0]), int(size[1]))
        self.mean, self.std, self.scale = tuple(mean), tuple(std), float(scale)
        self.antialias = antialias
        self.layout = self._LAYOUTS[layout]
        self.rows_per_cta = rows_per_cta

    def __call__(self, batch_u8: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        return ops.preprocess_u8hwc(batch_u8, self.size, self.layout, self.mean, self.std, self.scale,
                                    self.antialias, self.rows_per_cta, out)


class Rescale(object):
    """Rescale the image in a sample to a given size (reference :397-427): tuple -> exact size; int ->
    shorter side matched, aspect ratio kept, int() truncation."""

    def __init__(self, output_size):
        assert isinstance(output_size, (int, tuple))
        self.output_size = output_size

    def __call__(self, sample):
        image, label, index = sample
        h, w = image.shape[:2]
        new_h, new_w = rescale_size(h, w, self.output_size)
        if not torch.cuda.is_available():
            raise SiaError("Rescale runs on the GPU (sia_preprocess_u8hwc); no CUDA device is visible")
        u8 = torch.from_numpy(np.ascontiguousarray(_unit_float_to_u8(image))).cuda().unsqueeze(0)
        out = ops.preprocess_u8hwc(u8, (new_h, new_w), ops.LAYOUT_NCHW_F32)
        img = out[0].permute(1, 2, 0).contiguous().cpu().numpy()
        return (img, label, index)


class RandomCrop(object):
    """Random crop (reference :430-458); indexing only."""

    def __init__(self, output_size):
        assert isinstance(output_size, (int, tuple))
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else output_size
        assert len(self.output_size) == 2

    def __call__(self, sample):
        image, label, index = sample
        h, w = image.shape[:2]
        new_h, new_w = self.output_size
        top = np.random.randint(0, h - new_h + 1)
        left = np.random.randint(0, w - new_w + 1)
        return (image[top: top + new_h, left: left + new_w], label, index)


class ToTensor(object):
    """HWC ndarray -> CHW tensor view (reference :461-473)."""

    def __call__(self, sample):
        image, label, index = sample
        return (torch.from_numpy(image.transpose((2, 0, 1))), label, index)


def _imread_u8(path) -> np.ndarray:
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


class HibaDataset(Dataset):
    """Map-style dataset over an ISIC metadata dataframe (reference :258-393): ``__getitem__`` returns
    ``(image, label, idx)`` after the transform; ``lookup_path`` returns the per-instance metadata dict."""

    def __init__(self, p_metadata_df, class_names, root_dir, transform=None):
        self.metadata_df = p_metadata_df
        self.root_dir = root_dir
        self.transform = transform
        self.image_count = len(self.metadata_df)
        self.class_names = class_names

    def __len__(self):
        return self.image_count

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        instance = self.lookup_path(idx)
        image_np = np.float32(_imread_u8(self.get_file_path(instance["image_name"]))) / 255.0
        label = self.class_names.index(instance["benign_malignant"])
        sample = (image_np, label, idx)
        if self.transform:
            sample = self.transform(sample)
        return sample

    def read_u8(self, idx) -> np.ndarray:
        """The raw uint8 HWC decode buffer -- what ``GpuBatchTransform`` starts from."""
        return _imread_u8(self.get_file_path(self.lookup_path(idx)["image_name"]))

    def get_class_names(self):
        return self.class_names

    def get_class(self, index):
        return self.class_names[index]

    def get_file_path(self, image_name):
        return os.path.join(self.root_dir, image_name + ".jpg")

    def lookup_path(self, idx):
        row = self.metadata_df.iloc[idx]
        image_name = row["isic_id"]
        return {"file_path": self.get_file_path(image_name), "image_name": image_name,
                "patient_id": row["patient_id"], "diagnosis": row["diagnosis"],
                "benign_malignant": row["benign_malignant"], "age": row["age_approx"], "sex": row["sex"],
                "location": row["anatom_site_general"], "skin_type": row["fitzpatrick_skin_type"],
                "skin_tone": row["skin_tone"], "control": row["control"]}
