"""Data parallelism for the DINO hot path: one process per GPU, one flat fp32 gradient buffer, ONE NCCL
all-reduce per step.

Replaces the reference's `DistributedDataParallel(model, find_unused_parameters=True)` (main.py:156) for the
training step: every trainable parameter's .grad is a view into one contiguous buffer, so
  * zeroing the gradients is one memset,
  * parameters that did not take part in the step contribute zeros (what find_unused_parameters achieves),
  * the gradient exchange is a 191 MB (DINO-4scale) sum-all-reduce over NVLink with DDP's averaging semantics (the
    1/world_size rides in the clipping pass); the transformer / head part of the buffer is reduced from a backward hook
    while the ResNet backward is still running (reduce_early), the backbone part right after the backward,
  * gradient clipping (engine.py:110, max_norm 0.1) is one norm + one scale over the flat buffer.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).

Two ways of getting the gradients into the buffer:
  * gather=False: every .grad IS its slice of the buffer during the backward; eagerly, autograd runs one tiny
    `grad += new` kernel per gradient arrival (640 parameters in a DINO step); under CUDA graphs the captured backward of
    each segment puts its gradients into the slices itself -- the weight-gradient kernels of parameter-owning Linears reduce
    straight into them, the rest is added with multi-tensor launches (graphs._TrainingGraph, linear.GradSinks);
  * gather=True: .grad is None during the backward, so autograd simply keeps the first gradient tensor of a
    parameter (no kernel) and adds in place only for a second arrival; collect() then moves everything into the buffer
    with ONE multi-tensor copy and re-points .grad at the slices for the optimizer.
Measured on the B200 DINO step in round 1 both took the same time (68.0 vs 67.9 ms); gather=False is the default and what
the in-graph accumulation builds on (the views must stay in place between steps: zero(), not `p.grad = None`)."""
from __future__ import annotations

import torch
import torch.distributed as dist


import os as _os
_EARLY = _os.environ.get("DATR_EARLY_REDUCE", "1") != "0"     # 0: one all-reduce after the whole backward


class FlatGradients:
    def __init__(self, model: torch.nn.Module, process_group=None, gather: bool = False, late=None):
        """`late(name) -> bool` marks the parameters whose gradients are produced LAST by the backward pass (the backbone in
        DINO: it runs first in the forward).  They are laid out at the end of the buffer, so `reduce_early()` -- called
        from a backward hook once every other gradient is final -- can start the all-reduce of the first part while the
        backward of the late part is still running (gradient exchange overlapped with compute)."""
        self.gather, self._pending = gather, False
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if late is not None:
            named = [(n, p) for n, p in named if not late(n)] + [(n, p) for n, p in named if late(n)]
        self.params = [p for _, p in named]
        self.split = sum(p.numel() for n, p in named if late is None or not late(n))     # first element of the late part
        self._early_handle = None
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
        if not gather and dev.type == "cuda":
            # CUDA-graph segments captured from now on add their parameter gradients into these views inside the captured
            # backward (datr_b200.graphs._TrainingGraph) instead of one eager `grad += new` kernel per parameter
            from . import graphs
            if graphs.ACTIVE is not None:
                graphs.ACTIVE.set_grad_sinks(self)

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

    def reduce_early(self):
        """Start the sum-all-reduce of the early part of the buffer (asynchronously: the collective is ordered after the
        work already enqueued on the current stream and runs beside what is enqueued next).  Call it from a backward hook
        once the gradients of every non-late parameter are final; all_reduce() finishes the job.  No-op on one rank, in
        gather mode, without a late part, or if it already ran in this step."""
        if self.world_size == 1 or self.gather or self._early_handle is not None or self.split in (0, self.numel) or not _EARLY:
            return
        self._early_handle = dist.all_reduce(self.flat[:self.split], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def all_reduce(self):
        """Sum the gradients over the ranks: one collective on the flat buffer, or -- after reduce_early() -- one on the
        late part plus the wait for the early one.  The division by the world size (DDP's averaging) rides in clip_()'s
        single scaling pass; call average_() instead if the step does not clip."""
        self.collect()
        ws = self.world_size
        self._unscaled = ws > 1
        if ws > 1:
            if self._early_handle is not None:
                dist.all_reduce(self.flat[self.split:], op=dist.ReduceOp.SUM, group=self.group)
                self._early_handle.wait()
                self._early_handle = None
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)

    def average_(self):
        """Apply the pending 1 / world_size of all_reduce() (a no-op after clip_(), which folds it into its own pass)."""
        if getattr(self, "_unscaled", False):
            self.flat.mul_(1.0 / self.world_size)
            self._unscaled = False

    def clip_(self, max_norm: float):
        """torch.nn.utils.clip_grad_norm_ on the flat buffer; returns the total norm of the averaged gradient (0-dim
        tensor, no host sync).  One norm + one scaling pass, which also applies the 1 / world_size left pending by
        all_reduce()."""
        self.collect()
        inv = 1.0 / self.world_size if getattr(self, "_unscaled", False) else 1.0
        self._unscaled = False
        norm = torch.linalg.vector_norm(self.flat) * inv
        self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0) * inv)
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
