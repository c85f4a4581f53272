"""Import the UNMODIFIED Python reference (/root/reference) with import shims.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does
not exist on the GPU box, so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py``
calls this.  ``tests/golden/make_golden.py`` uses it to generate the committed fixtures.

Shims: stub ``matplotlib.pyplot``, ``pyaudio``, ``simpleaudio``,
``python_speech_features``, ``sidekit.frontend.features``; ``np.int = int`` (used at
utils/processing.py:79,83, removed from numpy >= 1.24).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SSP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "GMM_UBM.py"))


class _Anything:
    """Attribute sink for GUI/plot symbols the hot path never touches (e.g. plt.cm.Blues)."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = _StubModule(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_shims(sidekit_mfcc=None, sidekit_plp=None):
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001 - the reference needs the removed alias
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("pyaudio", PyAudio=object, paInt16=8)
    _stub("simpleaudio")
    _stub("python_speech_features")
    sk = _stub("sidekit")
    fe = _stub("sidekit.frontend")
    ft = _stub("sidekit.frontend.features")
    sk.frontend = fe
    fe.features = ft
    if sidekit_mfcc is not None or not hasattr(ft, "mfcc"):
        ft.mfcc = sidekit_mfcc
    if sidekit_plp is not None or not hasattr(ft, "plp"):
        ft.plp = sidekit_plp
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load(module: str, **kw):
    """Import ``utils.processing`` or ``GMM_UBM`` from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_shims(**kw)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module(module)
