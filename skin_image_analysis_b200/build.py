"""Builds libsia_b200.so (hand-written sm_100a CUDA behind the C ABI of include/sia_b200.h) and, separately,
libsia_b200_debug.so (the bring-up probes of include/sia_b200_debug.h; never loaded by the product path).

    python -m skin_image_analysis_b200.build [--force] [-v]

nvcc cross-compiles without a GPU; the .so lands in-tree next to this file (git-ignored, but it
travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_NAME = "libsia_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
STAMP = LIB_PATH + ".srchash"
DEBUG_LIB_PATH = os.path.join(HERE, "libsia_b200_debug.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libsia_b200.so")
    return exe


def source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, "sia_b200.h"), os.path.join(INCLUDE, "sia_b200_debug.h")]
    for name in files:
        path = name if os.path.isabs(name) else os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(DEBUG_LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_fresh():
        return LIB_PATH
    jobs = [(LIB_PATH, "libsia_unity.cu"), (DEBUG_LIB_PATH, "libsia_debug_unity.cu")]
    procs = []
    for out, unity in jobs:                                  # the two libraries compile side by side
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-o", out, os.path.join(CSRC, unity)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for cmd, proc in procs:
        out_text, err_text = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out_text + err_text)
        if verbose:
            print(err_text)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
