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


def fused_applicable(qk: torch.Tensor, v: torch.Tensor, num_heads: int, blocked, dropout_p: float) -> bool:
    """The tensor-core kernel of csrc/attn_fused.cu covers fp32 CUDA tensors, 32 channels per head, no dropout and a
    [T, T] boolean mask (or none)."""
    return (qk.is_cuda and qk.dtype == torch.float32 and v.dtype == torch.float32 and qk.dim() == 3 and dropout_p == 0.0
            and v.shape[-1] == 32 * num_heads and qk.shape[-1] == 2 * v.shape[-1] and qk.is_contiguous() and v.is_contiguous()
            and (blocked is None or (blocked.dtype == torch.bool and tuple(blocked.shape) == (qk.shape[1], qk.shape[1]))))


def pack_mask(blocked, T: int, device, transposed: bool = True):
    """[T, T] bool (True = may NOT attend) or None -> bit-packed int32 [T, words] for the fused kernels (indices >= T are
    packed as blocked).  Returns (bits, bits_t): bits along the keys (forward, dQ pass) and -- with `transposed` -- along the
    queries (dK / dV pass), else None."""
    lib = native.lib()
    words = lib.datr_attn_mask_words(T)
    bits = torch.empty((2 if transposed else 1, T, words), dtype=torch.int32, device=device)
    src = blocked.contiguous() if blocked is not None else None
    with torch.cuda.device(device):
        rc = lib.datr_attn_pack_mask(src.data_ptr() if src is not None else None, T, bits[0].data_ptr(),
                                     bits[1].data_ptr() if transposed else None, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"datr_attn_pack_mask failed (code {rc}): {lib.datr_attn_fused_last_error().decode()}")
    return bits[0], (bits[1] if transposed else None)


# "fused": both directions on the tcgen05 kernels; "gemm": fused forward that also writes the probabilities, backward as
# library batched GEMMs around the softmax-backward kernel (the round-2 intermediate, kept for comparison)
import os
_BACKWARD = os.environ.get("DATR_ATTENTION_BACKWARD", "fused")


class _FusedSelfAttention(torch.autograd.Function):
    """qk [N, T, 2C] (queries in columns [0, C), keys in [C, 2C)), v [N, T, C], packed masks -> [N, T, C].  Forward = ONE
    tcgen05 kernel, backward = two launches of a second one (dQ; dK + dV) that rebuild the score tiles in tensor memory:
    nothing of size T x T is ever written to HBM."""

    @staticmethod
    def forward(ctx, qk, v, bits, bits_t, H):
        N, T, C = v.shape
        d = C // H
        scale = 1.0 / math.sqrt(d)
        need_grad = qk.requires_grad or v.requires_grad
        gemm_bwd = need_grad and (_BACKWARD == "gemm" or bits_t is None)
        lib = native.lib()
        with torch.cuda.device(qk.device):
            out = torch.empty((N, T, C), dtype=torch.float32, device=qk.device)
            lse = torch.empty((N, H, T), dtype=torch.float32, device=qk.device)
            p = torch.empty((N * H, T, T), dtype=torch.float32, device=qk.device) if gemm_bwd else None
            rc = lib.datr_attn_fused_forward(qk.data_ptr(), 2 * C, qk.data_ptr() + 4 * C, 2 * C, v.data_ptr(), C,
                                             bits.data_ptr(), N, H, T, scale, out.data_ptr(), lse.data_ptr(),
                                             p.data_ptr() if p is not None else None, torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"datr_attn_fused_forward failed (code {rc}): {lib.datr_attn_fused_last_error().decode()}")
        if gemm_bwd:
            ctx.save_for_backward(qk, v, p)
        elif need_grad:
            ctx.save_for_backward(qk, v, bits, bits_t, out, lse)
        ctx.gemm_bwd = gemm_bwd
        ctx.scale, ctx.shape = scale, (N, H, T, d)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        N, H, T, d = ctx.shape
        C = H * d
        lib = native.lib()
        if not ctx.gemm_bwd:
            qk, v, bits, bits_t, out, lse = ctx.saved_tensors
            go = go.contiguous()
            with torch.cuda.device(go.device):
                dqk, dv = torch.empty_like(qk), torch.empty_like(v)
                delta = torch.empty_like(lse)
                rc = lib.datr_attn_fused_backward(qk.data_ptr(), 2 * C, qk.data_ptr() + 4 * C, 2 * C, v.data_ptr(), C,
                                                  bits.data_ptr(), bits_t.data_ptr(), N, H, T, ctx.scale, out.data_ptr(),
                                                  lse.data_ptr(), go.data_ptr(), delta.data_ptr(), dqk.data_ptr(), 2 * C,
                                                  dqk.data_ptr() + 4 * C, 2 * C, dv.data_ptr(), C,
                                                  torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise RuntimeError(f"datr_attn_fused_backward failed (code {rc}): {lib.datr_attn_fused_last_error().decode()}")
            return dqk, dv, None, None, None
        qk, v, p = ctx.saved_tensors
        heads = lambda t: t.reshape(N, T, H, d).permute(0, 2, 1, 3).reshape(N * H, T, d)     # noqa: E731
        q3, k3, v3, go3 = heads(qk[..., :C]), heads(qk[..., C:]), heads(v), heads(go)
        fallbacks.note("torch.bmm (cuBLAS) decoder self-attention backward (dV, dP, dQ, dK)", 4)
        dv = torch.bmm(p.transpose(1, 2), go3)
        ds = torch.bmm(go3, v3.transpose(1, 2))                           # dP, turned into dS in place
        with torch.cuda.device(go.device):
            rc = lib.datr_attn_softmax_backward(p.data_ptr(), ds.data_ptr(), ctx.scale, N * H * T, T,
                                                torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            _raise(lib, rc, "datr_attn_softmax_backward")
        dqk = torch.empty_like(qk).view(N, T, 2, H, d)
        dqk[:, :, 0] = torch.bmm(ds, k3).view(N, H, T, d).permute(0, 2, 1, 3)
        dqk[:, :, 1] = torch.bmm(ds.transpose(1, 2), q3).view(N, H, T, d).permute(0, 2, 1, 3)
        return dqk.view(N, T, 2 * C), dv.view(N, H, T, d).permute(0, 2, 1, 3).reshape(N, T, C), None, None, None


def fused_self_attention(qk, v, num_heads, blocked=None, bits=None):
    """softmax(q k^T / sqrt(d), -inf where blocked) v for the packed projections qk [N, T, 2C], v [N, T, C]; returns
    [N, T, C] (heads concatenated: the input of out_proj).  `bits` = pack_mask(blocked, T, device) if the caller already
    has it (a decoder pass packs once for its six layers)."""
    if bits is None:
        bits = pack_mask(blocked, qk.shape[1], qk.device, transposed=qk.requires_grad or v.requires_grad)
    return _FusedSelfAttention.apply(qk, v, bits[0], bits[1], num_heads)


def self_attention(q, k, v, blocked=None):
    """q, k, v [N, H, T, d] fp32 CUDA; blocked [T, T] bool, True = may NOT attend (nn.MultiheadAttention's convention)."""
    return _SelfAttention.apply(q, k, v, blocked)
