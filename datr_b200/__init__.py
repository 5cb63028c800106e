"""datr_b200 -- B200-native (sm_100a) implementation of DATR's data-parallel hot path.

The package holds the CUDA kernels + C ABI (csrc/, include/datr_msda.h) and the host-side mirror
of the reference interface for that path (models/dino/..., MultiScaleDeformableAttention shim).
"""
__version__ = "0.1.0"

from . import native  # noqa: F401


def install_dropin():
    """Make the reference's import names resolve to this package:
    `MultiScaleDeformableAttention` (native op module) and `models` (models.registry, models.dino.*).

    Every module of datr_b200.models is imported once and registered under its `models.*` name as the SAME module
    object, so `from models.registry import MODULE_BUILD_FUNCS` (main.py:81) sees the 'dino' entry registered by
    datr_b200.models.dino.dino, whatever the import order, and classes are not duplicated under two names."""
    import importlib
    import pkgutil
    import sys
    from . import MultiScaleDeformableAttention as _msda
    _msda.install()
    from . import models as _models
    for info in pkgutil.walk_packages(_models.__path__, prefix=_models.__name__ + "."):
        importlib.import_module(info.name)
    prefix = _models.__name__
    for name, module in list(sys.modules.items()):
        if name == prefix or name.startswith(prefix + "."):
            sys.modules.setdefault("models" + name[len(prefix):], module)
    return _models
