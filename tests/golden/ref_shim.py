"""Import shim for the UNMODIFIED reference (ankitdhall/learning_embeddings).

Only usable where /root/reference exists (the build container). It is used by
tests/golden/make_golden.py to produce the committed golden vectors; nothing
under tests/ -m gpu, bench.py or smoke() imports it (the GPU box has no
/root/reference).

The reference imports four modules that are absent here and carry no
arithmetic (tensorboardX, matplotlib, skimage, lime); they are replaced by
inert stand-ins so `import network.order_embeddings` etc. succeed.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("LEC_REFERENCE_ROOT", "/root/reference")


class _Inert:
    """Object whose every attribute/call is another inert object."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __iter__(self):
        return iter(())


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert


_STUB_ROOTS = ("tensorboardX", "matplotlib", "skimage", "lime")


class _InertFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serves an inert module for any (sub)module of the absent plotting/logging packages."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _InertModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "network"))


def install():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not any(isinstance(f, _InertFinder) for f in sys.meta_path):
        missing = []
        for root in _STUB_ROOTS:
            try:
                importlib.import_module(root)
            except Exception:
                missing.append(root)
        if missing:
            sys.meta_path.append(_InertFinder())
    for p in (REF_ROOT, os.path.join(REF_ROOT, "network")):
        if p not in sys.path:
            sys.path.insert(0, p)


def load(module: str):
    """e.g. load('network.order_embeddings_h')"""
    install()
    return importlib.import_module(module)
