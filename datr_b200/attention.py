"""Decoder self-attention with the score matrix kept in HBM (host side of include/datr_attn.h).

`self_attention(q, k, v, blocked)` == softmax(q k^T / sqrt(d) with -inf where blocked) v for q, k, v [N, H, T, d]
(what nn.MultiheadAttention computes in the reference's decoder layer, models/dino/deformable_transformer.py:880-897):
two library batched GEMMs around the in-place masked-softmax kernel of csrc/attn_softmax.cu, and in the backward
    dV = P^T dO,   dP = dO V^T,   dS = softmax-backward(P, dP) (in place, one kernel),   dQ = dS K,   dK = dS^T Q.
At DINO's sizes (T <= 1100, 8 heads x 32 channels) this is ~2.5x faster than PyTorch's memory-efficient SDPA kernel on
B200 (tools/bench_sdpa.py).  CUDA fp32 only, no dropout; callers keep F.scaled_dot_product_attention otherwise."""
from __future__ import annotations

import math

import torch

from . import fallbacks, native


def applicable(q: torch.Tensor, blocked, dropout_p: float) -> bool:
    return (q.is_cuda and q.dtype == torch.float32 and q.dim() == 4 and q.shape[2] <= 2048 and dropout_p == 0.0
            and (blocked is None or (blocked.dtype == torch.bool and tuple(blocked.shape) == (q.shape[2], q.shape[2]))))


def _raise(lib, rc, what):
    raise RuntimeError(f"{what} failed (code {rc}): {lib.datr_attn_last_error().decode()}")


class _SelfAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, blocked):
        N, H, T, d = q.shape
        scale = 1.0 / math.sqrt(d)
        q3, k3, v3 = (t.reshape(N * H, T, d) for t in (q, k, v))          # copies the strided head views (2 MB each)
        fallbacks.note("torch.bmm (cuBLAS) decoder self-attention Q K^T and P V", 2)
        p = torch.bmm(q3, k3.transpose(1, 2))                             # [N*H, T, T] scores, then probabilities
        lib = native.lib()
        mask = blocked.contiguous() if blocked is not None else None
        with torch.cuda.device(q.device):
            rc = lib.datr_attn_softmax_forward(p.data_ptr(), mask.data_ptr() if mask is not None else None, scale,
                                               N * H * T, T, T, torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            _raise(lib, rc, "datr_attn_softmax_forward")
        o = torch.bmm(p, v3)
        ctx.save_for_backward(q3, k3, v3, p)
        ctx.scale, ctx.shape = scale, (N, H, T, d)
        return o.view(N, H, T, d)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        q3, k3, v3, p = ctx.saved_tensors
        N, H, T, d = ctx.shape
        go3 = go.reshape(N * H, T, d)
        fallbacks.note("torch.bmm (cuBLAS) decoder self-attention backward (dV, dP, dQ, dK)", 4)
        dv = torch.bmm(p.transpose(1, 2), go3)
        ds = torch.bmm(go3, v3.transpose(1, 2))                           # dP, turned into dS in place
        lib = native.lib()
        with torch.cuda.device(go.device):
            rc = lib.datr_attn_softmax_backward(p.data_ptr(), ds.data_ptr(), ctx.scale, N * H * T, T,
                                                torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            _raise(lib, rc, "datr_attn_softmax_backward")
        dq = torch.bmm(ds, k3)
        dk = torch.bmm(ds.transpose(1, 2), q3)
        return dq.view(N, H, T, d), dk.view(N, H, T, d), dv.view(N, H, T, d), None


def self_attention(q, k, v, blocked=None):
    """q, k, v [N, H, T, d] fp32 CUDA; blocked [T, T] bool, True = may NOT attend (nn.MultiheadAttention's convention)."""
    return _SelfAttention.apply(q, k, v, blocked)
