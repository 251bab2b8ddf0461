"""Import the real reference (``/root/reference``) in the build container.

Test infrastructure only (see ``oracle/__init__.py``).  The reference cannot be
imported as shipped: seven third-party modules that the hot path never touches
are absent.  We insert empty stub modules for exactly those and
put ``/root/reference`` on ``sys.path``.  Nothing here is available on the GPU
box (``/root/reference`` does not travel); callers must check ``available()``.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DLPM_REFERENCE_ROOT", "/root/reference")
_STUBS = ["torchquad", "matplotlib", "matplotlib.pyplot", "matplotlib.animation",
          "imageio", "prdc", "pyemd"]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dlpm", "methods"))


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless dummy callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        dummy = type(name, (), {"__init__": lambda self, *a, **k: None,
                                "__call__": lambda self, *a, **k: None})
        setattr(self, name, dummy)
        return dummy


def load():
    """Return a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                mod = _Anything(name)
                mod.__path__ = []  # behave like a package
                sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.glp = importlib.import_module("dlpm.methods.GenerativeLevyProcess")
    ns.dlpm = importlib.import_module("dlpm.methods.dlpm")
    ns.Data = importlib.import_module("bem.datasets.Data")
    ns.Distributions = importlib.import_module("bem.datasets.Distributions")
    ns.unet = importlib.import_module("dlpm.models.unet")
    ns.Model = importlib.import_module("dlpm.models.Model")
    ns.sampler = importlib.import_module("dlpm.methods.LIM.functions.sampler")
    ns.sde = importlib.import_module("dlpm.methods.LIM.functions.sde")
    return ns
