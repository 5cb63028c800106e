"""Multi-tensor EMA (host side of include/datr_ema.h): `ema = ema * d + (1 - d) * model` for a whole state dict.

`StateDictEMA(ema_tensors, model_tensors)` pairs the floating-point tensors of two state dicts once; `update(d)` then
runs ONE kernel over all CUDA fp32 pairs (plus one short launch per extra alias of shared tensors, see below) (pointer / chunk tables live in device memory) and the reference's two
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
        # A state dict may list one storage under several names (DINO with dec_pred_*_embed_share: bbox_embed.0..5,
        # class_embed.0..5 and transformer.decoder.*_embed.* are the same tensors, 12 names each).  The reference loop
        # (models/dino/EMA.py:47-50) then applies the update once PER NAME, i.e. k times to such a tensor.  The kernel
        # must not see an aliased tensor twice in one launch (concurrent read-modify-write), so identical
        # (ema, model, numel) pairs are folded into one table entry with a multiplicity k and the update is launched in
        # k rounds, round r covering the entries with multiplicity > r: the same k sequential applications, bit for bit.
        # Pairs whose EMA storages overlap in any other way take the sequential path below.
        uniq, mult = {}, {}
        self.slow = []
        for e, m in pairs:
            if not fast(e, m):
                self.slow.append((e, m))
                continue
            key = (e.device, e.data_ptr(), m.data_ptr(), e.numel())
            uniq.setdefault(key, (e, m))
            mult[key] = mult.get(key, 0) + 1
        spans = sorted((k[1], k[1] + k[3] * 4, k) for k in uniq)
        clash = set()
        for (a0, a1, ka), (b0, b1, kb) in zip(spans, spans[1:]):
            if b0 < a1 and ka[0] == kb[0]:
                clash.update((ka, kb))
        for key in clash:
            self.slow += [uniq[key]] * mult[key]
        self.fast = [(uniq[k], mult[k]) for k in uniq if k not in clash]
        self.rounds = {}         # device -> list of (segs, chunks, n_chunks) per round
        self._ptrs = [(e, m, e.data_ptr(), m.data_ptr()) for (e, m), _ in self.fast]
        for dev in {e.device for (e, _), _ in self.fast}:
            mine = [(e, m, k) for (e, m), k in self.fast if e.device == dev]
            tables = []
            for r in range(max(k for _, _, k in mine)):
                live = [(e, m) for e, m, k in mine if k > r]
                segs = np.array([[e.data_ptr(), m.data_ptr(), e.numel()] for e, m in live], dtype=np.int64)
                chunks = np.array([[i, off] for i, (e, _) in enumerate(live) for off in range(0, e.numel(), CHUNK)], dtype=np.int64)
                tables.append((torch.from_numpy(segs).to(dev), torch.from_numpy(chunks).to(dev), len(chunks)))
            self.rounds[dev] = tables

    def stale(self) -> bool:
        """True if a paired tensor moved (load_state_dict keeps storages; .to(device) or re-assignment does not)."""
        return any(e.data_ptr() != pe or m.data_ptr() != pm for e, m, pe, pm in self._ptrs)

    @torch.no_grad()
    def update(self, d: float):
        lib = native.lib() if self.rounds else None
        for dev, tables in self.rounds.items():
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream().cuda_stream
                for segs, chunks, n in tables:
                    rc = lib.datr_ema_update(segs.data_ptr(), chunks.data_ptr(), n, float(d), float(1.0 - d), stream)
                    if rc != 0:
                        raise RuntimeError(f"datr_ema_update failed (code {rc}): {lib.datr_ema_last_error().decode()}")
        for e, m in self.slow:
            e *= d
            e += (1.0 - d) * m.detach()
