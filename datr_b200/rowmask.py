"""Padding-mask fill of a token matrix on the CUDA kernel of csrc/rowmask.cu (host side of include/datr_rowmask.h).

`zero_masked_rows(x, mask)` == `x.masked_fill(mask[..., None], 0.0)` (reference models/dino/ops/modules/
ms_deform_attn.py:96-97) for a freshly produced CUDA fp32 `x` [..., C]: the rows are zeroed IN PLACE (only the masked
rows are written), and the backward zeroes the same rows of the incoming gradient in place.  `x` must not be needed
unmasked by anyone else -- the caller passes the output of the value projection, which has no other consumer."""
from __future__ import annotations

import torch

from . import native


def _zero(x: torch.Tensor, mask: torch.Tensor) -> None:
    lib = native.lib()
    rows = mask.numel()
    with torch.cuda.device(x.device):
        rc = lib.datr_zero_masked_rows(x.data_ptr(), mask.data_ptr(), rows, x.numel() // rows,
                                       torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"datr_zero_masked_rows failed (code {rc}): {lib.datr_rowmask_last_error().decode()}")


class _ZeroMaskedRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        _zero(x, mask)
        ctx.mark_dirty(x)
        ctx.save_for_backward(mask)
        return x

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        # the masked tensor has a single consumer (the MSDeformAttn op), so `g` is that op's freshly allocated
        # grad_value (possibly viewed) and nobody else reads it: zero its padding rows in place
        if not g.is_contiguous() or g.data_ptr() % 16:
            g = g.contiguous().clone() if g.data_ptr() % 16 else g.contiguous()
        _zero(g, mask)
        return g, None


def zero_masked_rows(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x [*mask.shape, ...] with mask (bool, True = padding): every x[i] with mask[i] set becomes zero.
    Falls back to masked_fill when the kernel does not apply."""
    if (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and mask.dtype == torch.bool and mask.is_contiguous()
            and mask.shape == x.shape[:mask.dim()] and mask.numel() > 0 and x.numel() > 0
            and (x.numel() // mask.numel()) % 4 == 0 and x.data_ptr() % 16 == 0
            and not (x.requires_grad and x.is_leaf)):
        return _ZeroMaskedRows.apply(x, mask)
    return x.masked_fill(mask.view(mask.shape + (1,) * (x.dim() - mask.dim())), 0.0)
