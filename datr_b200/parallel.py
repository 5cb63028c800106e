"""Data parallelism for the DINO hot path: one process per GPU, one flat fp32 gradient buffer, ONE NCCL
all-reduce per step.

Replaces the reference's `DistributedDataParallel(model, find_unused_parameters=True)` (main.py:156) for the
training step: every trainable parameter's .grad is a view into one contiguous buffer, so
  * zeroing the gradients is one memset,
  * parameters that did not take part in the step contribute zeros (what find_unused_parameters achieves),
  * the gradient exchange is a single 191 MB (DINO-4scale) sum-all-reduce over NVLink followed by a scale by
    1/world_size -- DDP's averaging semantics,
  * gradient clipping (engine.py:110, max_norm 0.1) is one norm + one scale over the flat buffer.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).

Two ways of getting the gradients into the buffer:
  * gather=False: every .grad IS its slice of the buffer during the backward; autograd then runs one tiny `grad += new`
    kernel per gradient arrival (1 200 of them in a DINO step: 640 parameters, the transformer's used by two passes);
  * gather=True: .grad is None during the backward, so autograd simply keeps the first gradient tensor of a
    parameter (no kernel) and adds in place only for a second arrival; collect() then moves everything into the buffer
    with ONE multi-tensor copy and re-points .grad at the slices for the optimizer.
Measured on the B200 DINO step both take the same time (68.0 vs 67.9 ms: with CUDA graphs the per-parameter
accumulation is a small part of the ~5 000 short kernels of a step), so the simpler gather=False stays the default."""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradients:
    def __init__(self, model: torch.nn.Module, process_group=None, gather: bool = False):
        self.gather, self._pending = gather, False
        self.params = [p for p in model.parameters() if p.requires_grad]
        assert self.params, "model has no trainable parameters"
        dev, dt = self.params[0].device, self.params[0].dtype
        assert all(p.device == dev and p.dtype == dt for p in self.params)
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        off = 0
        self.views = []
        for p in self.params:
            # the view takes the parameter's own (dense) strides, e.g. channels_last convolution weights
            assert p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last), "parameters must be dense"
            self.views.append(self.flat[off:off + p.numel()].as_strided(p.size(), p.stride()))
            p.grad = self.views[-1]
            off += p.numel()
        self.group = process_group

    @property
    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        """Start of a step: zero gradients (gather mode: detach .grad from the buffer until collect())."""
        if self.gather:
            for p in self.params:
                p.grad = None
            self._pending = True
        else:
            self.flat.zero_()

    def collect(self):
        """Gather mode, after the backward: zero the buffer (parameters that took no part keep zeros), copy every
        gradient autograd produced into its slice with one multi-tensor copy, point .grad back at the slices."""
        if not self._pending:
            return
        self._pending = False
        self.flat.zero_()
        dst = [v for p, v in zip(self.params, self.views) if p.grad is not None]
        src = [p.grad for p in self.params if p.grad is not None]
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def check_views(self):
        """True while every .grad still aliases the flat buffer (an optimizer.zero_grad(set_to_none=True) breaks it)."""
        base = self.flat.data_ptr()
        return all(p.grad is not None and base <= p.grad.data_ptr() < base + self.numel * self.flat.element_size()
                   for p in self.params)

    def all_reduce(self):
        """Average the gradients over the ranks: one collective on the flat buffer."""
        self.collect()
        ws = self.world_size
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / ws)

    def clip_(self, max_norm: float):
        """torch.nn.utils.clip_grad_norm_ on the flat buffer; returns the total norm (0-dim tensor, no host sync)."""
        self.collect()
        norm = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm


def broadcast_parameters(model: torch.nn.Module, src: int = 0, group=None):
    """Initial parameter/buffer sync (the broadcast DDP does at construction)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def param_groups(model, lr=1e-4, lr_backbone=1e-5):
    """The two parameter groups of the reference (util/get_param_dicts.py:23-31): 'backbone' in the name -> lr_backbone."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    return [{"params": [p for n, p in named if "backbone" not in n], "lr": lr},
            {"params": [p for n, p in named if "backbone" in n], "lr": lr_backbone}]
