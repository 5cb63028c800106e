from .dino import build_dino  # noqa: F401
