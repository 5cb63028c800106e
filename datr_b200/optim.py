"""One-launch gradient clipping + AdamW for the training step (host side of include/datr_adamw.h).

`FlatAdamW(param_groups, grads)` takes the parameter groups of torch.optim.AdamW ({"params", "lr"[, "weight_decay"]}) and
the step's FlatGradients buffer (datr_b200.parallel).  `clip_and_step(max_norm)` = torch.nn.utils.clip_grad_norm_ followed
by optimizer.step() of the reference's iteration (engine.py:108-111): one norm reduction, a few scalar kernels for the
clipping coefficient (no host sync) and ONE kernel that scales each gradient on the fly and updates parameter and both
moments.  Moments live in two flat buffers laid out like the gradients.  Same arithmetic as torch's fused AdamW
(tests/test_optim_gpu.py compares a few hundred steps)."""
from __future__ import annotations

import math
import struct

import numpy as np
import torch

from . import native

CHUNK = 16384   # DATR_ADAMW_CHUNK


class FlatAdamW:
    def __init__(self, param_groups, grads, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.grads = grads
        self.betas, self.eps = betas, eps
        self.param_groups = [dict(g) for g in param_groups]
        for g in self.param_groups:
            g.setdefault("weight_decay", weight_decay)
            g["params"] = list(g["params"])
        view_of = {id(p): v for p, v in zip(grads.params, grads.views)}
        params = [p for g in self.param_groups for p in g["params"]]
        assert all(id(p) in view_of for p in params), "every optimised parameter must own a slice of the flat gradient buffer"
        assert all(p.is_cuda and p.dtype == torch.float32 for p in params), "FlatAdamW covers CUDA fp32 parameters"
        dev = params[0].device
        self.device = dev
        self.exp_avg = torch.zeros_like(grads.flat)
        self.exp_avg_sq = torch.zeros_like(grads.flat)
        self.step_count = 0
        base = grads.flat.data_ptr()
        self._entries = []      # (param, grad slice address, byte offset into the flat buffers, group)
        for g in self.param_groups:
            for p in g["params"]:
                v = view_of[id(p)]
                # the gradient slice shares the parameter's (dense) strides, so element i of the parameter's storage pairs
                # with element i of its slice
                self._entries.append((p, v.data_ptr(), v.data_ptr() - base, g))
        chunks = np.array([[i, off] for i, (p, _, _, _) in enumerate(self._entries) for off in range(0, p.numel(), CHUNK)],
                          dtype=np.int64)
        self._chunks = torch.from_numpy(chunks).to(dev)
        self._n_chunks = len(chunks)
        self._segs = None
        self._seg_key = None

    def _table(self):
        key = tuple((p.data_ptr(), g["lr"], g["weight_decay"]) for p, _, _, g in self._entries)
        if key != self._seg_key:          # first step, a changed learning rate (lr_drop) or re-allocated parameters
            m0, v0 = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
            rows = []
            for p, gptr, off, g in self._entries:
                packed = struct.unpack("<q", struct.pack("<ff", float(g["lr"]), float(g["weight_decay"])))[0]
                rows.append([p.data_ptr(), gptr, m0 + off, v0 + off, p.numel(), packed])
            self._segs = torch.from_numpy(np.array(rows, dtype=np.int64)).to(self.device)
            self._seg_key = key
        return self._segs

    @torch.no_grad()
    def clip_and_step(self, max_norm=None):
        """Clip the (all-reduced) flat gradient to `max_norm` (None / <= 0: no clipping) and apply one AdamW step.
        Returns the total gradient norm (0-dim tensor) or None."""
        g = self.grads
        g.collect()
        inv = 1.0 / g.world_size if getattr(g, "_unscaled", False) else 1.0
        g._unscaled = False
        norm = scale = None
        if max_norm is not None and max_norm > 0:
            norm = torch.linalg.vector_norm(g.flat) * inv
            scale = (torch.clamp(max_norm / (norm + 1e-6), max=1.0) * inv).reshape(1)
        elif inv != 1.0:
            scale = torch.full((1,), inv, dtype=torch.float32, device=self.device)
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1 ** self.step_count
        bc2_sqrt = math.sqrt(1.0 - b2 ** self.step_count)
        lib = native.lib()
        segs = self._table()
        with torch.cuda.device(self.device):
            rc = lib.datr_adamw_step(segs.data_ptr(), self._chunks.data_ptr(), self._n_chunks,
                                     scale.data_ptr() if scale is not None else None, b1, b2, self.eps, bc1, bc2_sqrt,
                                     torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"datr_adamw_step failed (code {rc}): {lib.datr_adamw_last_error().decode()}")
        return norm

    def step(self):
        return self.clip_and_step(None)

    def zero_grad(self, set_to_none=False):
        self.grads.zero()
