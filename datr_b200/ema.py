"""Multi-tensor EMA (host side of include/datr_ema.h): `ema = ema * d + (1 - d) * model` for a whole state dict.

`StateDictEMA(ema_tensors, model_tensors)` pairs the floating-point tensors of two state dicts once; `update(d)` then
runs ONE kernel over all CUDA fp32 pairs (pointer / chunk tables live in device memory) and the reference's two
in-place ops (`v *= d; v += (1 - d) * m`, models/dino/EMA.py:47-50) for anything else -- CPU tensors, other dtypes --
so the numbers are those of the reference loop in every case (bit-identical: same two fp32 roundings)."""
from __future__ import annotations

import numpy as np
import torch

from . import native

CHUNK = 16384       # DATR_EMA_CHUNK


class StateDictEMA:
    def __init__(self, ema_tensors, model_tensors):
        pairs = [(e, m) for e, m in zip(ema_tensors, model_tensors) if e.dtype.is_floating_point]
        for e, m in pairs:
            if e.shape != m.shape:
                raise ValueError("EMA and model state dicts do not line up")
        fast = lambda e, m: (e.is_cuda and m.is_cuda and e.device == m.device and e.dtype == torch.float32 and m.dtype == torch.float32
                             and e.is_contiguous() and m.is_contiguous() and e.numel() > 0)
        self.fast = [(e, m) for e, m in pairs if fast(e, m)]
        self.slow = [(e, m) for e, m in pairs if not fast(e, m)]
        self.tables = {}
        for dev in {e.device for e, _ in self.fast}:
            mine = [(e, m) for e, m in self.fast if e.device == dev]
            segs = np.array([[e.data_ptr(), m.data_ptr(), e.numel()] for e, m in mine], dtype=np.int64)
            chunks = np.array([[i, off] for i, (e, _) in enumerate(mine) for off in range(0, e.numel(), CHUNK)], dtype=np.int64)
            self.tables[dev] = (torch.from_numpy(segs).to(dev), torch.from_numpy(chunks).to(dev), len(chunks),
                                [(e.data_ptr(), m.data_ptr()) for e, m in mine], mine)

    def stale(self) -> bool:
        """True if a paired tensor moved (load_state_dict keeps storages; .to(device) or re-assignment does not)."""
        return any((e.data_ptr(), m.data_ptr()) != p for _, _, _, ptrs, mine in self.tables.values() for (e, m), p in zip(mine, ptrs))

    @torch.no_grad()
    def update(self, d: float):
        lib = native.lib() if self.tables else None
        for dev, (segs, chunks, n, _, _) in self.tables.items():
            with torch.cuda.device(dev):
                rc = lib.datr_ema_update(segs.data_ptr(), chunks.data_ptr(), n, float(d), float(1.0 - d),
                                         torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise RuntimeError(f"datr_ema_update failed (code {rc}): {lib.datr_ema_last_error().decode()}")
        for e, m in self.slow:
            e *= d
            e += (1.0 - d) * m.detach()
