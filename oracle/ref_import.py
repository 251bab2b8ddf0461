"""Import the real reference: ``/root/reference`` in the build container, ``oracle/_ref`` on the GPU box.

Test / bench infrastructure only (see ``oracle/__init__.py``).  The reference cannot be imported as shipped: seven
third-party modules that the hot path never touches are absent.  We insert empty stub modules
for exactly those and put the reference root on ``sys.path``.  ``/root/reference`` does not travel to the GPU box; the
verbatim, git-ignored copy that ``oracle/install_ref.py`` places under ``oracle/_ref`` does.  Callers must check
``available()``.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = ["torchquad", "matplotlib", "matplotlib.pyplot", "matplotlib.animation",
          "imageio", "prdc", "pyemd"]


def _has_ref(root):
    return bool(root) and os.path.isdir(os.path.join(root, "dlpm", "methods"))


def reference_root():
    """First existing of: $DLPM_REFERENCE_ROOT, /root/reference (build container), oracle/_ref (travels with gpurun)."""
    for cand in (os.environ.get("DLPM_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if _has_ref(cand):
            return cand
    return None


REFERENCE_ROOT = reference_root()


def available() -> bool:
    return reference_root() is not None


def kind() -> str:
    """'reference' = the tree under /root/reference, '_ref' = the installed verbatim copy, '' = none."""
    root = reference_root()
    if root is None:
        return ""
    return "_ref" if os.path.abspath(root) == os.path.abspath(os.path.join(_HERE, "_ref")) else "reference"


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless dummy callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        dummy = type(name, (), {"__init__": lambda self, *a, **k: None,
                                "__call__": lambda self, *a, **k: None})
        setattr(self, name, dummy)
        return dummy


_ns = None


def load():
    """Return a namespace with the reference's hot-path modules."""
    global _ns
    if _ns is not None:
        return _ns
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (neither /root/reference nor oracle/_ref)")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                mod = _Anything(name)
                mod.__path__ = []  # behave like a package
                sys.modules[name] = mod
    if root not in sys.path:
        sys.path.insert(0, root)
    ns = types.SimpleNamespace()
    ns.root = root
    ns.glp = importlib.import_module("dlpm.methods.GenerativeLevyProcess")
    ns.dlpm = importlib.import_module("dlpm.methods.dlpm")
    ns.Data = importlib.import_module("bem.datasets.Data")
    ns.Distributions = importlib.import_module("bem.datasets.Distributions")
    ns.unet = importlib.import_module("dlpm.models.unet")
    ns.Model = importlib.import_module("dlpm.models.Model")
    ns.sampler = importlib.import_module("dlpm.methods.LIM.functions.sampler")
    ns.sde = importlib.import_module("dlpm.methods.LIM.functions.sde")
    ns.GenerationManager = importlib.import_module("bem.GenerationManager")
    ns.utils_ema = importlib.import_module("bem.utils_ema")
    _ns = ns
    return ns
