"""The batched evaluation engine: uint8 decode buffers in, per-group confusion counts out.

One ``EvalEngine`` per process / GPU.  Per batch it launches, on one stream and captured in one
CUDA graph per input slot:

    sia_preprocess_u8hwc        [B,H,W,3] u8   -> [B,224,224,4] bf16        (K1-K3)
    sia_conv7x7_c3_relu_pool2                  -> [B,112,112,32] bf16       (K4)
    sia_conv3x3_relu_pool2  x2                 -> [B,28,28,128] bf16        (K4)
    sia_linear_splitk                          -> [S,B,512] fp32 partials   (K5)
    sia_head_tail                              -> logp [B,2], pred [B], counts += ...   (K6 + K7)

Inputs live in ``n_slots`` device-resident slots (u8 images, labels, group ids) so that host->device
copies of batch k+1 overlap the kernels of batch k.  The count tensor stays on the device for the whole
shard; it is read back (and, multi-GPU, all-reduced -- see distributed.py) once at the end.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import SiaError
from .tone_bias_model import CnnPlan

N_ATTR = 3        # Fitzpatrick type, sex, control   
N_GROUPS = 6


def plan_from_state_dict(state: dict, device) -> CnnPlan:
    """Reference ``state_dict`` (either architecture) -> packed weights on ``device``."""
    if any(k.startswith("layers.") for k in state):
        conv_keys, fc_keys = ["layers.0", "layers.3", "layers.6"], ["layers.10", "layers.13", "layers.16"]
    elif all(k.split(".")[0].isdigit() for k in state):          # an nn.Sequential (define_isic_model)
        names = sorted({k.rsplit(".", 1)[0] for k in state}, key=int)
        conv_keys = [n for n in names if state[n + ".weight"].dim() == 4]
        fc_keys = [n for n in names if state[n + ".weight"].dim() == 2]
    else:
        conv_keys = [k for k in ("conv1", "conv2", "conv3", "conv4") if k + ".weight" in state]
        fc_keys = ["fc4", "fc5", "fc6"]
    g = lambda k: state[k].to(device=device, dtype=torch.float32)      # noqa: E731
    return CnnPlan([(g(k + ".weight"), g(k + ".bias")) for k in conv_keys],
                   [(g(k + ".weight"), g(k + ".bias")) for k in fc_keys])


class EvalEngine:
    def __init__(self, state_dict: dict, batch: int, src_hw=(450, 600), out_size: int = 224, device=None,
                 mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0), use_graph: bool = True, rows_per_cta: int = 32,
                 n_slots: int = 2):
        if not torch.cuda.is_available():
            raise SiaError("EvalEngine needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.batch, self.src_hw, self.out_size = batch, tuple(src_hw), out_size
        self.mean, self.std, self.rows_per_cta = mean, std, rows_per_cta
        self.n_slots = n_slots
        with torch.cuda.device(self.device):
            self.plan = plan_from_state_dict(state_dict, self.device)
            if self.plan.image_size != out_size:
                raise SiaError(f"state_dict is for {self.plan.image_size}x{self.plan.image_size} inputs, "
                               f"out_size is {out_size}")
            d = self.device
            self.u8 = [torch.empty((batch, src_hw[0], src_hw[1], 3), dtype=torch.uint8, device=d)
                       for _ in range(n_slots)]
            self.label = [torch.zeros((batch,), dtype=torch.uint8, device=d) for _ in range(n_slots)]
            self.groups = [torch.full((N_ATTR, batch), 255, dtype=torch.uint8, device=d) for _ in range(n_slots)]
            self.x4 = torch.zeros((batch, out_size, out_size + ops.NHWC4_PAD, 4), dtype=torch.bfloat16, device=d)
            self.counts = torch.zeros((N_ATTR, N_GROUPS, 2, 2), dtype=torch.int64, device=d)
            ws = self.plan.workspace(batch)
            self.logp, self.pred = ws["logp"], ws["pred"]
            self.stream = torch.cuda.Stream(device=d)
            self.copy_stream = torch.cuda.Stream(device=d)
            self.slot_ready = [torch.cuda.Event() for _ in range(n_slots)]     # H2D into the slot finished
            self.slot_free = [torch.cuda.Event() for _ in range(n_slots)]      # kernels reading the slot finished
            self.graphs = [None] * n_slots
            # preprocess + the forward pass (first block: one launch per 32 output channels, 3x3 blocks, fc1, tail)
            self.launches_per_batch = 1 + self.plan.launches
            if use_graph:
                self._capture()

    # --------------------------------------------------------------------------------------------
    def _launch_all(self, slot: int):
        ops.preprocess_u8hwc(self.u8[slot], (self.out_size, self.out_size), ops.LAYOUT_NHWC4_BF16, self.mean,
                             self.std, rows_per_cta=self.rows_per_cta, out=self.x4)
        self.plan.forward_nhwc4(self.x4, label=self.label[slot], groups=self.groups[slot], n_groups=N_GROUPS,
                                counts=self.counts)

    def _capture(self):
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self.u8[0].zero_()
                self._launch_all(0)                     # un-captured warm-up: module load, smem attributes
                self.stream.synchronize()
                for slot in range(self.n_slots):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        self._launch_all(slot)
                    self.graphs[slot] = g
                self.counts.zero_()
                self.stream.synchronize()
            torch.cuda.current_stream().wait_stream(self.stream)

    # --------------------------------------------------------------------------------------------
    def reset_counts(self):
        with torch.cuda.stream(self.stream):
            self.counts.zero_()

    def step_resident(self, slot: int = 0):
        """One batch whose inputs already sit in input slot ``slot`` (device memory)."""
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if self.graphs[slot] is not None:
                self.graphs[slot].replay()
            else:
                self._launch_all(slot)

    def load_slot(self, slot: int, u8, label, groups):
        """Asynchronous copy (host pinned or device source) into an input slot on the copy stream.  Device sources
        may still be being written by work queued on the caller's current stream: the copy is ordered after it."""
        with torch.cuda.device(self.device):
            if any(isinstance(t, torch.Tensor) and t.is_cuda for t in (u8, label, groups)):
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.device(self.device), torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.slot_free[slot])
            self.u8[slot].copy_(u8, non_blocking=True)
            self.label[slot].copy_(label, non_blocking=True)
            self.groups[slot].copy_(groups, non_blocking=True)
            for t in (u8, label, groups):                      # keep temporaries alive until the copy has read them
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(self.copy_stream)
            self.slot_ready[slot].record(self.copy_stream)

    def step_slot(self, slot: int):
        """Kernels of the batch in ``slot`` once its copy has landed."""
        self.stream.wait_event(self.slot_ready[slot])
        self.step_resident(slot)
        self.slot_free[slot].record(self.stream)

    def step(self, u8, label, groups, slot: int = 0):
        """Public per-batch call: copy the batch (host pinned or device tensors) in and evaluate it."""
        self.load_slot(slot, u8, label, groups)
        self.step_slot(slot)

    def forward_u8(self, u8: torch.Tensor):
        """Convenience for tests: (logp, pred) of one batch without touching the counts."""
        if u8.shape[0] != self.batch:
            raise ValueError("batch size mismatch")
        self.synchronize()
        saved = self.counts.clone()
        self.step(u8, self.label[0], self.groups[0], 0)
        self.synchronize()
        self.counts.copy_(saved)
        return self.logp.clone(), self.pred.clone()

    def synchronize(self):
        self.copy_stream.synchronize()
        self.stream.synchronize()

    def read_counts(self) -> torch.Tensor:
        self.synchronize()
        return self.counts.cpu()

    def allreduce_counts(self) -> torch.Tensor:
        """The job's single collective (multi-GPU): sums the count tensor over all ranks ON THE ENGINE STREAM, i.e.
        ordered after every batch queued so far -- ``torch.distributed.all_reduce`` only orders against the caller's
        current stream, so calling it on ``eng.counts`` from the default stream could reduce a partial tensor."""
        from . import distributed as D
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            D.allreduce_counts(self.counts)
        return self.counts
