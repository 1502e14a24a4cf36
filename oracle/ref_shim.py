"""Import shim that makes the reference's hot-path modules importable in the
build container (never on the GPU box: ``/root/reference`` does not exist there).

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` to generate
golden vectors from the reference's own modules, and by optional cross-check
tests that skip when ``/root/reference`` is absent.

What it does:
* puts ``/root/reference`` on ``sys.path``;
* stubs the non-numeric packages that are missing from this image
  (``jsonpickle``, ``matplotlib*``, ``simple_parsing*``, ``typing_inspect``,
  ``h5py``, ``xmltodict``, ``colorspacious``, ``pydensecrf``, ``higher``);
* registers ``oracle/normflows_restated.py`` as ``normflows``;
* wraps ``ReduceLROnPlateau``/``StepLR`` to accept-and-drop the ``verbose`` kwarg
  that torch 2.11 removed (reference ``awesome/model/path_connected_net.py:932-933``).
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AWESOME_REFERENCE_ROOT", "/root/reference")


class _Anything:
    """Stands in for any class/function/decorator of a stubbed package."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]          # used as a bare decorator
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __getitem__(self, item):
        return self


class _StubModule(types.ModuleType):
    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []      # looks like a package, so submodule imports resolve
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return type(name, (_Anything,), {})


class _StubFinder:
    """Resolves ``import stubbed.pkg.sub`` for every registered stub root."""

    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


_STUB_ROOTS = ("jsonpickle", "matplotlib", "mpl_toolkits", "simple_parsing", "typing_inspect",
               "h5py", "xmltodict", "colorspacious", "pydensecrf", "higher")

_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "awesome"))


def install() -> None:
    """Idempotent.  Raises RuntimeError when the reference tree is absent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    missing = []
    for root in _STUB_ROOTS:
        try:
            __import__(root)
        except Exception:
            missing.append(root)
    if missing:
        sys.meta_path.append(_StubFinder(missing))
    try:
        import normflows  # noqa: F401
    except Exception:
        from oracle import normflows_restated
        nf = normflows_restated.as_module()
        sys.modules["normflows"] = nf
        sys.modules["normflows.nets"] = nf.nets
        sys.modules["normflows.flows"] = nf.flows
        sys.modules["normflows.distributions"] = nf.distributions
        sys.modules["normflows.distributions.base"] = nf.distributions.base
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    import torch.optim.lr_scheduler as lrs

    def _drop_verbose(cls):
        orig = cls.__init__
        if getattr(orig, "_awb_wrapped", False):
            return

        def __init__(self, *a, verbose=None, **k):
            orig(self, *a, **k)
        __init__._awb_wrapped = True
        cls.__init__ = __init__

    _drop_verbose(lrs.ReduceLROnPlateau)
    _drop_verbose(lrs.StepLR)
    _installed = True
