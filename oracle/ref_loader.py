"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference scripts under Python 3.

The reference (happyjin/Gaussian_process, mounted read-only at /root/reference) is five
Python-2 scripts.  This loader never edits them on disk: it reads the bytes, applies an
in-memory py2->py3 syntax shim and exec()s them into fresh module
objects whose ``__name__`` is not ``__main__`` so the drivers do not run.

It exists only in the build container (``/root/reference`` is absent on the GPU box), so it
may be used by ``oracle/gen_golden.py`` and by ``-m "not gpu"`` tests that skip when the
reference is absent.  Nothing in the product package imports it.
"""
from __future__ import annotations

import os
import re
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("GPX_REFERENCE_ROOT", "/root/reference")
LOAD_ORDER = (
    "GP_regression",
    "tune_hyperparms_regression",
    "CO2_example",
    "GP_binary_classification",
    "GP_multi_classification",
)

_PRINT_RE = re.compile(r"^(\s*)print\s+(?!\()(.*)$")
_BACKTICK_RE = re.compile(r"`([^`]*)`")


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, m + ".py")) for m in LOAD_ORDER)


class _NoOp:
    """Callable / attribute sink used for matplotlib stand-ins."""

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


def _stub_module(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    sink = _NoOp()
    mod.__getattr__ = lambda attr: sink  # type: ignore[attr-defined]
    return mod


def _install_stubs() -> None:
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.colors"):
        if name not in sys.modules:
            sys.modules[name] = _stub_module(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]  # type: ignore[attr-defined]
    sys.modules["matplotlib"].mlab = sys.modules["matplotlib.mlab"]  # type: ignore[attr-defined]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]  # type: ignore[attr-defined]
    try:
        import sklearn.datasets
        import sklearn.model_selection

        if "sklearn.cross_validation" not in sys.modules:
            cv = types.ModuleType("sklearn.cross_validation")
            cv.train_test_split = sklearn.model_selection.train_test_split
            sys.modules["sklearn.cross_validation"] = cv
        if not hasattr(sklearn.datasets, "fetch_mldata"):
            sklearn.datasets.fetch_mldata = _NoOp()
    except ImportError:  # sklearn is only needed by the classifier scripts' imports
        pass


def _py3_source(raw: bytes) -> str:
    text = raw.decode("utf-8").replace("\r\n", "\n").replace("\r", "\n")
    out = []
    for line in text.split("\n"):
        m = _PRINT_RE.match(line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = _BACKTICK_RE.sub(lambda mm: "repr(%s)" % mm.group(1), line)
        out.append(line)
    return "\n".join(out)


_loaded: dict[str, types.ModuleType] = {}


def load_reference() -> dict[str, types.ModuleType]:
    """Return {module name: module} for the five reference scripts (cached)."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise FileNotFoundError("reference scripts not found under %s" % REFERENCE_ROOT)
    _install_stubs()
    real_spo = np.set_printoptions

    def _spo(*a, **k):
        thr = k.get("threshold", 0)
        if isinstance(thr, float) and thr != thr:  # threshold=np.nan (py2-era numpy)
            k.pop("threshold")
        return None  # never change the test process' print options

    saved = {name: sys.modules.get(name) for name in LOAD_ORDER}
    np.set_printoptions = _spo  # type: ignore[assignment]
    try:
        for name in LOAD_ORDER:
            path = os.path.join(REFERENCE_ROOT, name + ".py")
            with open(path, "rb") as fh:
                src = _py3_source(fh.read())
            mod = types.ModuleType(name)
            mod.__file__ = path
            sys.modules[name] = mod  # cross-imports between the scripts resolve to these
            exec(compile(src, path, "exec"), mod.__dict__)
            _loaded[name] = mod
    finally:
        np.set_printoptions = real_spo  # type: ignore[assignment]
        # do not leave reference modules importable by name: the product package ships
        # modules with the same names and tests must not confuse the two.
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return _loaded
