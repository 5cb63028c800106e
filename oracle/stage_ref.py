"""oracle/stage_ref.py -- TEST / BASELINE INFRASTRUCTURE ONLY.

Stages the reference's Python tree (models/, util/, config/, engine.py -- unmodified) from where it lies under
/root/reference into the git-ignored `baseline/_ref/`, the one place the task contract reserves for the
reference install.  `baseline/_ref/` is NOT gpurun-ignored, so it travels to the GPU box, where
/root/reference does not exist; `tests/ref_loader.py` imports the reference model from there, and
`tools/bench_reference_gpu.py` / `bench.py --impl reference` run it as the baseline.  Nothing is staged into
tracked paths and no product module imports anything from here.

The reference is not a pip package (no setup.py / pyproject at its root; the only setup.py is the native op's,
which refuses to build without a visible GPU, ops/setup.py:48-49), so `pip install --target baseline/_ref`
does not apply: this copy IS the install.  The native op travels separately as oracle/_ref/*.so (build_ref.py).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
WHAT = ("models", "util", "config", "engine.py", "LICENSE")


def staged() -> bool:
    return os.path.isdir(os.path.join(DST, "models", "dino"))


def stage(force: bool = False) -> str | None:
    """Returns the staged root, or None when neither /root/reference nor an earlier staging exists."""
    if not os.path.isdir(os.path.join(SRC, "models", "dino")):
        return DST if staged() else None
    if staged() and not force:
        return DST
    os.makedirs(DST, exist_ok=True)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "build", "*.egg-info", "src")
    for name in WHAT:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=ignore, dirs_exist_ok=True)
        elif os.path.exists(s):
            shutil.copy2(s, d)
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
