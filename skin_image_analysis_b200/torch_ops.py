"""``torch.ops.sia_b200.*``: the hot-path operators registered with TORCH_LIBRARY (csrc_torch/torch_ops.cpp), a thin C++
shim over the same C ABI (include/sia_b200.h, libsia_b200.so) the ctypes binding uses.

    from skin_image_analysis_b200 import torch_ops
    torch_ops.load()                      # builds sia_b200_torch.so in-tree on first use (g++, no nvcc needed)
    y = torch.ops.sia_b200.conv3x3_relu_pool2(x_nhwc, w_packed, bias, cout)

Every op TORCH_CHECKs device / dtype / contiguity / shape and launches on the current CUDA stream.  The ctypes path
(``ops.py``) stays for hosts without torch's C++ headers; both call the same kernels, bit for bit
(tests/test_torch_ops.py).
"""
from __future__ import annotations

import os
import subprocess
import sys

import torch

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc_torch", "torch_ops.cpp")
LIB_PATH = os.path.join(HERE, "sia_b200_torch.so")
_loaded = False


def is_fresh() -> bool:
    return (os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC)
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(os.path.join(_build.INCLUDE, "sia_b200.h")))


def build(force: bool = False) -> str:
    """g++ -shared torch_ops.cpp -> sia_b200_torch.so next to libsia_b200.so (linked with rpath $ORIGIN)."""
    if not force and is_fresh():
        return LIB_PATH
    _build.build()
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{_build.INCLUDE}"]
    lib_dirs = ce.library_paths("cuda")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = (["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
            "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, SRC, "-o", LIB_PATH]
           + [f"-L{d}" for d in lib_dirs] + [f"-L{HERE}", "-l:libsia_b200.so", "-lc10", "-lc10_cuda", "-ltorch_cpu",
                                            "-ltorch", "-Wl,-rpath,$ORIGIN"] + [f"-Wl,-rpath,{d}" for d in lib_dirs])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB_PATH


def load():
    """Builds (if stale) and registers the ``sia_b200`` operator namespace; returns ``torch.ops.sia_b200``."""
    global _loaded
    if not _loaded:
        torch.ops.load_library(build())
        _loaded = True
    return torch.ops.sia_b200


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
