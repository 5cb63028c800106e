"""`models` package of the drop-in: models.registry.MODULE_BUILD_FUNCS and models.dino.* with the
reference's import paths (reference models/__init__.py:8 imports build_dino the same way)."""
from .dino import build_dino  # noqa: F401  (registers 'dino' in MODULE_BUILD_FUNCS)
