"""Bookkeeping of LIBRARY code taken on the hot path (cuBLAS / cuDNN / ATen kernels standing where a hand-written
kernel does not exist yet).  Nothing here changes behaviour: call sites `note(name)` when they hand work to a library,
bench.py snapshots the counters over eager steps and lists them in its JSON line (`library_calls_per_step`), so the
"no silent fallback" rule is checkable: every library GEMM / convolution of the step is named there."""
from __future__ import annotations

import collections

_counts = collections.Counter()
enabled = False          # bench.py switches this on around the steps it inspects


def note(name: str, n: int = 1) -> None:
    if enabled:
        _counts[name] += n


def reset() -> None:
    _counts.clear()


def snapshot() -> dict:
    return dict(sorted(_counts.items()))
