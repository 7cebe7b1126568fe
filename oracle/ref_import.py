"""Import the UNMODIFIED reference modules from /root/reference/src (TEST INFRASTRUCTURE ONLY).

Only usable in the build container -- /root/reference does not exist on the GPU box, so
nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.  It is used by
``tests/golden/make_golden.py`` (fixture generation) and by the optional
``tests/test_reference_live.py`` checks, which skip when the directory is absent.

The reference imports ``skimage``, ``matplotlib`` and ``optuna`` at module scope
(tone_bias_dataset.py:33,41; tone_bias_test.py:39-47; tone_bias_optuna.py); none is installed
here, so inert stub modules are registered first.  The only stubbed function
that is ever *called* on the evaluation path is ``skimage.transform.resize``
(tone_bias_dataset.py:425), which is bound to ``oracle.resize.resize_scipy``.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


class _Inert:
    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


def _stub(name: str, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__file__ = "<oracle.ref_import stub>"
        sys.modules[name] = mod
    mod.__dict__.update(attrs)
    return mod


def load():
    """Returns a namespace with the reference modules: dataset, model, test, analysis, hiba."""
    if not available():
        raise RuntimeError(f"{REFERENCE_SRC} is not present (reference only exists in the build container)")
    import torch  # noqa: F401  (import the real heavy modules before any stub is registered)
    import torchvision  # noqa: F401
    import pandas  # noqa: F401
    from oracle import resize as _resize

    def _lazy(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()

    def _have(name):
        try:
            __import__(name)
            return True
        except Exception:
            return False

    if not _have("skimage"):
        sk = _stub("skimage")
        sk.io = _stub("skimage.io", imread=_Inert())
        sk.transform = _stub("skimage.transform", resize=_resize.resize_scipy)
    if not _have("matplotlib"):
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot", rc=lambda *a, **k: None, __getattr__=_lazy)
    if not _have("optuna"):
        op = _stub("optuna", __getattr__=_lazy)
        op.trial = _stub("optuna.trial", TrialState=_Inert())
        op.exceptions = _stub("optuna.exceptions", TrialPruned=Exception)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import importlib
    ns = types.SimpleNamespace()
    ns.dataset = importlib.import_module("tone_bias_dataset")
    ns.model = importlib.import_module("tone_bias_model")
    ns.test = importlib.import_module("tone_bias_test")
    ns.analysis = importlib.import_module("tone_bias_analysis")
    ns.hiba = importlib.import_module("jgi_hiba_2022_model")
    return ns
