"""TEST INFRASTRUCTURE ONLY -- import the real (unmodified) Optiland reference.

The reference (``/root/reference``, pure Python) imports matplotlib / vtk /
seaborn at module import time on the trace path
(``optiland/backend/numpy_backend.py:13``, ``optiland/physical_apertures/base.py:18``),
none of which is installed in this image.  ``import_reference()`` installs a
``sys.meta_path`` finder that serves inert stub modules for those plotting
packages and puts the reference on ``sys.path``.  Nothing of the reference enters the
repository's history.  ``/root/reference`` does not exist on the GPU box; ``scripts/make_ref.sh`` stages
an unmodified copy under ``oracle/_ref/`` (git-ignored, shipped by gpurun), which is what this helper
finds there.  Used by ``oracle/make_golden.py``, the plugin tests (CPU: oracle engine; ``-m gpu``: the real
CUDA engine under live Optiland objects) and ``bench.py``'s reference arm / ``e2e_optic_trace`` key.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    env = os.environ.get("OPTILAND_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "optiland")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
REFERENCE_TESTS = os.path.join(REFERENCE_ROOT, "tests")
_STUB_ROOTS = ("matplotlib", "mpl_toolkits", "vtk", "vtkmodules", "seaborn")


class _StubModule(types.ModuleType):
    """A module whose every missing attribute is a MagicMock (classes usable as bases)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        value = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, value)
        return value


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []  # behave as a package so sub-imports resolve
        return mod

    def exec_module(self, module):
        return None


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "optiland"))


def import_reference():
    """Return the imported ``optiland`` package of the unmodified reference."""
    if not reference_available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        for root in _STUB_ROOTS:
            try:
                __import__(root)
            except ImportError:
                pass
        sys.meta_path.append(_StubFinder())  # after the real finders: only fills gaps
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import optiland  # noqa: F401

    return optiland
