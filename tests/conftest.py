import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests never silently pass on a box without a GPU: they are skipped there, loudly."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    """On a failing GPU test, say whether a kernel watchdog fired (the word lives in pinned host memory,
    so it is readable even after the trap killed the CUDA context)."""
    outcome = yield
    rep = outcome.get_result()
    if rep.when == "call" and rep.failed and "gpu" in item.keywords:
        try:
            from skin_image_analysis_b200 import _lib
            wd = _lib.load().sia_watchdog_status(0)
            rep.sections.append(("sia watchdog", f"0x{wd:08x} (site {(wd >> 16) & 0x7fff}, block {wd & 0xffff})"))
        except Exception as exc:      # pragma: no cover
            rep.sections.append(("sia watchdog", f"unavailable: {exc}"))
