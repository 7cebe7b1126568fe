"""B200-native batched-evaluation path of jpope8/skin-image-analysis.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the work runs in
hand-written sm_100a CUDA behind the C ABI of ``include/sia_b200.h`` (``libsia_b200.so``).
Module names mirror the reference's (``tone_bias_dataset``, ``tone_bias_model``,
``jgi_hiba_2022_model``, ``tone_bias_test``, ``tone_bias_analysis``) so the path drops in.
"""
__version__ = "0.1.0"
