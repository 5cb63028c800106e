"""datr_b200 -- B200-native (sm_100a) implementation of DATR's data-parallel hot path.

The package holds the CUDA kernels + C ABI (csrc/, include/datr_msda.h) and the host-side mirror
of the reference interface for that path (models/dino/..., MultiScaleDeformableAttention shim).
"""
__version__ = "0.1.0"

from . import native  # noqa: F401


def install_dropin():
    """Make the reference's import names resolve to this package:
    `MultiScaleDeformableAttention` (native op module) and `models` (models.registry, models.dino.*)."""
    import sys
    from . import MultiScaleDeformableAttention as _msda
    _msda.install()
    from . import models as _models
    sys.modules.setdefault("models", _models)
    return _models
