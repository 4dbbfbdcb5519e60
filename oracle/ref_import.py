"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *real* reference package from /root/reference (authoring container
only; that path does not exist on the GPU box) so that golden vectors can be
generated from the reference's own layer classes.

The snapshot of the reference cannot be imported as-is:
  * libfewshot_core/__init__.py pulls in trainer.py/test.py which import the
    missing packages libfewshot_core.data, libfewshot_core.data.collates and
    libfewshot_core.config (reference trainer.py:16, test.py:15,31);
  * backbone/vit_class_aware.py:23 imports timm, which is not installed.
We pre-seed sys.modules with empty stand-ins for exactly those four things.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AFS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "libfewshot_core"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (libfewshot_core.model, libfewshot_core.utils, libfewshot_core.audio_augmentations)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "libfewshot_core.model" not in sys.modules:
        data = _stub("libfewshot_core.data", get_dataloader=None)
        data.__path__ = []
        coll = _stub("libfewshot_core.data.collates", get_mean_std=None)
        data.collates = coll
        _stub("libfewshot_core.config", Config=None)
        timm = _stub("timm")
        timm.__path__ = []
        tm = _stub("timm.models")
        tm.__path__ = []
        _stub("timm.models.registry", register_model=lambda f: f)
        _stub("timm.models.layers", trunc_normal_=None, DropPath=None, to_2tuple=None)
        timm.models = tm
    arch = importlib.import_module("libfewshot_core.model")
    utils = importlib.import_module("libfewshot_core.utils")
    aug = importlib.import_module("libfewshot_core.audio_augmentations")
    return arch, utils, aug
