#!/usr/bin/env python
"""One small launch of every product kernel, for compute-sanitizer (memcheck / racecheck) under gpurun:

    compute-sanitizer --tool memcheck python tools/sanitize_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from skin_image_analysis_b200 import ops
from skin_image_analysis_b200.engine import EvalEngine
from skin_image_analysis_b200.synthetic import random_state_dict

g = torch.Generator(device="cuda").manual_seed(0)
u8 = torch.randint(0, 256, (3, 450, 600, 3), dtype=torch.uint8, device="cuda", generator=g)
for layout in (ops.LAYOUT_NCHW_F32, ops.LAYOUT_NCHW_BF16, ops.LAYOUT_NHWC4_BF16):
    ops.preprocess_u8hwc(u8, (224, 224), layout)
ops.preprocess_u8hwc(u8, (224, 224), ops.LAYOUT_NHWC4_BF16, impl="cuda_core")
ops.preprocess_u8hwc(u8[:1], (512, 512), ops.LAYOUT_NHWC4_BF16)
label = torch.randint(0, 2, (3,), dtype=torch.uint8, device="cuda", generator=g)
groups = torch.randint(0, 7, (3, 3), dtype=torch.uint8, device="cuda", generator=g)
for kind in ("SkinCancerListModel", "SkinCancerModel", "optuna_best"):
    eng = EvalEngine(random_state_dict(kind, 224, seed=1), 3, (450, 600), 224, use_graph=False, n_slots=1)
    eng.step(u8, label, groups)
    eng.synchronize()
    print(kind, eng.logp.cpu().numpy().round(3).tolist(), int(eng.read_counts()[0].sum()))
pred = torch.randint(0, 2, (1000,), dtype=torch.uint8, device="cuda", generator=g)
lab = torch.randint(0, 2, (1000,), dtype=torch.uint8, device="cuda", generator=g)
grp = torch.randint(0, 8, (3, 1000), dtype=torch.uint8, device="cuda", generator=g)
print(int(ops.confusion_counts(pred, lab, grp, 6).sum()))
torch.cuda.synchronize()
print("sanitize target done")
