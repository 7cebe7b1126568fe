"""Drop-in for the transform part of ``tone_bias_dataset`` (reference src/tone_bias_dataset.py).

Reference surface kept: ``HibaDataset`` (:258-393), ``Rescale`` (:397-427), ``RandomCrop`` (:430-458),
``ToTensor`` (:461-473), ``convert_type2tone`` (:84-98) -- callables on the sample tuple
``(image HWC float32 in [0,1], label, index)``.

The arithmetic of the chain  u8 -> /255 (:335) -> skimage resize (:425) -> CHW (:470)  runs in the fused
``sia_preprocess_u8hwc`` kernel.  Three ways to reach it, all parity-tested:

  * ``GpuBatchTransform``: [B,H,W,3] uint8 on the device in, the whole normalised batch out, one launch
    (what the evaluation engine uses).
  * ``Rescale`` called in the MAIN process (``dataset[i]``, ``DataLoader(num_workers=0)``): one image, one
    launch, returns the ndarray the reference returns.
  * ``Rescale`` called inside a ``DataLoader`` WORKER process (the reference's own call site,
    ``DataLoader(test_dataset, batch_size=16, shuffle=True, num_workers=10)``, tone_bias_test.py:637): a forked
    worker cannot initialise CUDA, so the transform is DEFERRED -- the worker returns the uint8 decode buffer
    tagged with the target size (``DeferredImage``), the worker's default collate packs them into a
    ``DeferredBatch`` (registered in ``torch.utils.data``'s collate map), and the reference loop's own
    ``images = images.to(device)`` (tone_bias_test.py:186) uploads the bytes and runs ONE batched kernel launch
    in the main process.  The call site stays unchanged and the transform is batched.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ops
from ._lib import SiaError
from .resize_weights import rescale_size

__all__ = ["HibaDataset", "Rescale", "RandomCrop", "ToTensor", "GpuBatchTransform", "convert_type2tone",
           "DeferredImage", "DeferredBatch", "decode_jpeg_batch"]


def convert_type2tone(row):
    """Fitzpatrick {I, II} -> 'light', anything else -> 'dark' (reference :84-98)."""
    return "light" if row["fitzpatrick_skin_type"] in ("I", "II") else "dark"


class _DecodedImage(np.ndarray):
    """``np.float32(u8) / 255.0`` (what the reference's ``__getitem__`` hands to the transform, :335) that still
    remembers the uint8 decode buffer it came from, so ``Rescale`` need not recover it.  Any view or slice is a plain
    array again (``__array_finalize__`` does not propagate the buffer)."""
    u8 = None

    @classmethod
    def from_u8(cls, u8: np.ndarray):
        obj = (np.float32(u8) / 255.0).view(cls)
        obj.u8 = u8
        return obj

    def __array_finalize__(self, obj):
        self.u8 = None


def _unit_float_to_u8(image: np.ndarray) -> np.ndarray:
    """Recovers the uint8 decode buffer from ``np.float32(u8)/255.0`` exactly, or raises."""
    if image.dtype == np.uint8:
        return image
    if isinstance(image, _DecodedImage) and image.u8 is not None and image.u8.shape == image.shape:
        return image.u8
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


class DeferredImage:
    """What ``Rescale`` returns inside a DataLoader worker: the uint8 HWC decode buffer + the size it has to be
    resized to.  Collates into a ``DeferredBatch``."""
    __slots__ = ("u8", "size")

    def __init__(self, u8: np.ndarray, size):
        self.u8 = np.ascontiguousarray(u8)
        self.size = (int(size[0]), int(size[1]))

    @property
    def shape(self):                      # HWC shape of the image the eager transform would have returned
        return (self.size[0], self.size[1], self.u8.shape[2])


class DeferredBatch:
    """A collated batch of ``DeferredImage``s.  ``.to(cuda_device)`` is the transform: the bytes are uploaded and
    resized by one ``sia_preprocess_u8hwc`` launch per distinct source shape; the result is the float32
    ``[B,3,h,w]`` tensor the reference's DataLoader would have yielded."""

    def __init__(self, images, sizes):
        self.images = list(images)        # uint8 HWC CPU tensors (shared memory across the worker boundary)
        self.sizes = [tuple(s) for s in sizes]

    def __len__(self):
        return len(self.images)

    def size(self, dim=None):
        shape = self.shape
        return shape if dim is None else shape[dim]

    @property
    def shape(self):
        h, w = self.sizes[0]
        return torch.Size((len(self.images), int(self.images[0].shape[2]), h, w))

    def pin_memory(self):
        return DeferredBatch([t.pin_memory() for t in self.images], self.sizes)

    def to(self, device=None, *args, **kwargs):
        device = torch.device("cuda" if device is None else device)
        if device.type != "cuda":
            raise SiaError("a deferred transform batch can only be materialised on a CUDA device "
                           "(sia_preprocess_u8hwc has no CPU fallback)")
        if len(set(self.sizes)) != 1:
            raise RuntimeError(f"stack expects each tensor to be equal size, but got {sorted(set(self.sizes))} "
                               "(Rescale(int) on images of different aspect ratios cannot be batched)")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(idx):
            h, w = self.sizes[0]
            out = torch.empty((len(self.images), 3, h, w), dtype=torch.float32, device=device)
            by_shape = {}
            for i, t in enumerate(self.images):
                by_shape.setdefault(tuple(t.shape), []).append(i)
            for shape, members in by_shape.items():
                u8 = torch.stack([self.images[i] for i in members]).to(device, non_blocking=True)
                res = ops.preprocess_u8hwc(u8, (h, w), ops.LAYOUT_NCHW_F32,
                                           out=out if len(by_shape) == 1 else None)
                if len(by_shape) > 1:
                    out[torch.as_tensor(members, device=device)] = res
        return out

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))


def _collate_deferred(batch, *, collate_fn_map=None):
    return DeferredBatch([torch.from_numpy(d.u8) for d in batch], [d.size for d in batch])


def _register_collate():
    """``torch.utils.data.default_collate`` dispatches on the element type through this map; registering at import
    time covers forked workers (they inherit it) and spawned workers (they import this module to unpickle the
    dataset's transform)."""
    from torch.utils.data._utils import collate as _c
    _c.default_collate_fn_map[DeferredImage] = _collate_deferred


_register_collate()


def _in_loader_worker() -> bool:
    from torch.utils.data import get_worker_info
    return get_worker_info() is not None


def decode_jpeg_batch(files, device="cuda") -> torch.Tensor:
    """SURVEY 8(f) row 4 -- JPEG decode in front of the transform kernel, on the GPU: ``files`` = paths or encoded byte
    buffers of same-sized JPEGs -> [B,H,W,3] uint8 decode buffers on ``device`` (what ``GpuBatchTransform`` /
    ``EvalEngine.step`` take).  The decode itself is nvJPEG through ``torchvision.io.decode_jpeg(device=...)`` (a library
    call: a decoder is outside the hand-written path); its planar output is interleaved by ``sia_chw_u8_to_hwc_u8``.
    nvJPEG and libjpeg (the reference reads with ``skimage.io.imread``, tone_bias_dataset.py:326) differ by a few grey
    levels in chroma up-sampling, so this front end is NOT the bit-exact parity path -- ``HibaDataset`` (PIL decode in
    the DataLoader workers) is."""
    import torchvision
    datas = []
    for f in files:
        if isinstance(f, (str, os.PathLike)):
            datas.append(torchvision.io.read_file(str(f)))
        else:
            datas.append(torch.frombuffer(bytearray(f), dtype=torch.uint8))
    device = torch.device(device)
    if device.type != "cuda":
        raise SiaError("decode_jpeg_batch decodes on a CUDA device")
    planes = torchvision.io.decode_jpeg(datas, device=device, mode=torchvision.io.ImageReadMode.RGB)
    if len({tuple(p.shape) for p in planes}) != 1:
        raise ValueError("decode_jpeg_batch needs same-sized images (batch them by size)")
    with torch.cuda.device(device):
        return ops.chw_to_hwc_u8(torch.stack(planes).contiguous())


class Rescale(object):
    """Rescale the image in a sample to a given size (reference :397-427): tuple -> exact size; int ->
    shorter side matched, aspect ratio kept, int() truncation.

    defer: None (default) -- run the kernel right away in the main process, defer inside a DataLoader worker;
    True -- always defer (also batches ``DataLoader(num_workers=0)``); False -- never defer (a forked worker then
    fails in ``torch.cuda`` with its own "Cannot re-initialize CUDA in forked subprocess")."""

    def __init__(self, output_size, defer=None):
        assert isinstance(output_size, (int, tuple))
        self.output_size = output_size
        self.defer = defer

    def __call__(self, sample):
        image, label, index = sample
        h, w = image.shape[:2]
        new_h, new_w = rescale_size(h, w, self.output_size)
        u8 = _unit_float_to_u8(image)
        if self.defer or (self.defer is None and _in_loader_worker()):
            return (DeferredImage(u8, (new_h, new_w)), label, index)
        if not torch.cuda.is_available():
            raise SiaError("Rescale runs on the GPU (sia_preprocess_u8hwc); no CUDA device is visible")
        dev_u8 = torch.from_numpy(np.ascontiguousarray(u8)).cuda().unsqueeze(0)
        out = ops.preprocess_u8hwc(dev_u8, (new_h, new_w), ops.LAYOUT_NCHW_F32)
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
        if isinstance(image, DeferredImage):
            raise SiaError("RandomCrop after a deferred Rescale: the resized pixels do not exist yet in a DataLoader "
                           "worker; crop before Rescale or use Rescale(..., defer=False) with num_workers=0")
        h, w = image.shape[:2]
        new_h, new_w = self.output_size
        top = np.random.randint(0, h - new_h + 1)
        left = np.random.randint(0, w - new_w + 1)
        return (image[top: top + new_h, left: left + new_w], label, index)


class ToTensor(object):
    """HWC ndarray -> CHW tensor view (reference :461-473)."""

    def __call__(self, sample):
        image, label, index = sample
        if isinstance(image, DeferredImage):          # the CHW order is produced by the deferred kernel launch
            return sample
        return (torch.from_numpy(image.transpose((2, 0, 1))), label, index)


def _imread_u8(path) -> np.ndarray:
    from PIL import Image
    with Image.open(path) as im:
        return np.array(im.convert("RGB"), dtype=np.uint8)       # a writable copy


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
        image_np = _DecodedImage.from_u8(_imread_u8(self.get_file_path(instance["image_name"])))
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
