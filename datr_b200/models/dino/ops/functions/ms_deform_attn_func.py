"""Autograd binding of the native op -- mirror of the reference's
models/dino/ops/functions/ms_deform_attn_func.py:21-38 (same class name, argument order, saved
tensors, `once_differentiable`, and the (grad_value, None, None, grad_loc, grad_attn, None) return).

The reference file also carries a pure-PyTorch implementation for debugging
(ms_deform_attn_core_pytorch, :41-61).  This package deliberately has none: the product path is the
CUDA library only and raises when it is unavailable; the CPU restatement lives under oracle/ and is
test infrastructure.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from datr_b200 import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = MSDA.ms_deform_attn_backward(
            value, shapes, level_start, loc, attn, grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


class MSDeformAttnFusedFunction(Function):
    """Extension (no counterpart class in the reference): the elementwise prologue of
    MSDeformAttn.forward (ops/modules/ms_deform_attn.py:99-111 -- softmax of the attention logits, sampling
    locations from reference points and offsets) evaluated inside the CUDA kernels, so neither tensor is
    materialised; the backward returns the gradients of the raw offsets and logits."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_offsets, attention_logits,
                reference_points):
        output = MSDA.ms_deform_attn_fused_forward(value, value_spatial_shapes, value_level_start_index,
                                                   sampling_offsets, attention_logits, reference_points)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_offsets,
                              attention_logits, reference_points)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, offsets, logits, ref = ctx.saved_tensors
        grad_value, grad_off, grad_logits = MSDA.ms_deform_attn_fused_backward(
            value, shapes, level_start, offsets, logits, ref, grad_output.contiguous())
        return grad_value, None, None, grad_off, grad_logits, None


class MSDeformAttnMergedFunction(Function):
    """MSDeformAttnFusedFunction fed by ONE Linear: `merged` [N, Lq, 3*M*L*P] holds the sampling offsets in columns
    [0, 2*M*L*P) and the attention logits in [2*M*L*P, 3*M*L*P) (the rows of `sampling_offsets.weight` and
    `attention_weights.weight` stacked), and the backward returns one gradient of the same layout, so the two
    projections of the query cost one GEMM forward and one dgrad + one wgrad backward."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, merged, reference_points, M, L, P,
                pair_dtype=None):
        """`pair_dtype` (torch.bfloat16 / torch.float16, optional): the forward gathers from 16-bit pair rows packed
        from `value` (include/datr_msda.h, two line gathers per sample instead of four); the backward still reads the
        fp32 rows, so the gradients are those of the fp32 op at the rounded-value forward point."""
        N, Lq = merged.shape[:2]
        T = M * L * P
        ctx.dims = (M, L, P)
        offsets = merged[..., :2 * T].view(N, Lq, M, L, P, 2)
        logits = merged[..., 2 * T:].view(N, Lq, M, L * P)
        pairs = None
        if pair_dtype is not None and P == 4:
            pairs = MSDA.pack_value_pairs(value, value_spatial_shapes, value_level_start_index, pair_dtype)
        output = MSDA.ms_deform_attn_fused_forward(value, value_spatial_shapes, value_level_start_index,
                                                   offsets, logits, reference_points, pairs=pairs)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, merged, reference_points)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, merged, ref = ctx.saved_tensors
        M, L, P = ctx.dims
        N, Lq = merged.shape[:2]
        T = M * L * P
        grad_merged = torch.empty_like(merged)
        grad_value, _, _ = MSDA.ms_deform_attn_fused_backward(
            value, shapes, level_start, merged[..., :2 * T].view(N, Lq, M, L, P, 2),
            merged[..., 2 * T:].view(N, Lq, M, L * P), ref, grad_output.contiguous(), merged_grad=grad_merged)
        return grad_value, None, None, grad_merged, None, None, None, None, None
