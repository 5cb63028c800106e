"""Drop-in for the reference's native extension module `MultiScaleDeformableAttention`.

Same two callables, argument order and error behaviour as the pybind module built from
models/dino/ops/src/vision.cpp:13-16 (dispatch ms_deform_attn.h:21-60, host code
cuda/ms_deform_attn_cuda.cu:20-153), implemented as a thin ctypes shim over the C ABI in
include/datr_msda.h.  `install()` registers it under the reference's import name so that
`import MultiScaleDeformableAttention as MSDA` (ms_deform_attn_func.py:18) resolves to it.

Thread-safety: no Python-side state; forward runs on the caller's thread, backward on autograd's
worker thread, both on the *current* torch CUDA stream of the tensors' device.
"""
from __future__ import annotations

import sys

import torch

from . import native

_DTYPES = {torch.float32: 0, torch.float64: 1}

# Optional launch timers: bench.py sets this to a list and each launch appends
# (kind, shape_key, start_event, end_event) recorded on the launching stream.
_timers = None

# Host copies of (spatial_shapes, level_start_index) per device tensor pair: the backward builds its TMA tensor maps
# from them (include/datr_msda.h, *_hs entry points).  Filled by one device->host read per distinct pair, never while a
# CUDA graph is being captured; a miss during capture just selects the vector-reduction scatter.
_HOST_GEOMETRY = {}


def host_geometry(spatial_shapes, level_start_index):
    """(int64[L,2] array, int64[L] array) ctypes copies of the two device tensors, or (None, None)."""
    key = (spatial_shapes.data_ptr(), spatial_shapes._version, level_start_index.data_ptr(), level_start_index._version,
           str(spatial_shapes.device))
    hit = _HOST_GEOMETRY.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            return None, None
        import ctypes
        sh = [int(v) for v in spatial_shapes.reshape(-1).tolist()]
        st = [int(v) for v in level_start_index.reshape(-1).tolist()]
        if len(_HOST_GEOMETRY) > 256:
            _HOST_GEOMETRY.clear()
        hit = _HOST_GEOMETRY[key] = ((ctypes.c_int64 * len(sh))(*sh), (ctypes.c_int64 * len(st))(*st))
    return hit


def _check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, extra=()):
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), *extra]
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")           # ms_deform_attn.h:38,60
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")  # cu:28-32, :93-98
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")        # cu:34-38, :100-105
        if t.device != value.device:
            raise RuntimeError(f"{name} must be on the same device as value")
    if value.dtype not in _DTYPES:
        # AT_DISPATCH_FLOATING_TYPES (cu:64,134): float and double only
        raise RuntimeError(f'"ms_deform_attn_cuda" not implemented for \'{value.dtype}\'')
    for name, t in named[3:]:
        if t.dtype != value.dtype:
            raise RuntimeError(f"{name} must have the dtype of value ({value.dtype}), got {t.dtype}")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64")
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5:
        raise RuntimeError("expected value[N,S,M,D], sampling_loc[N,Lq,M,L,P,2], attn_weight[N,Lq,M,L,P]")
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    if tuple(sampling_loc.shape) != (N, Lq, M, L, P, 2) or tuple(attn_weight.shape) != (N, Lq, M, L, P) \
            or level_start_index.numel() != L or tuple(spatial_shapes.shape) != (L, 2):
        raise RuntimeError("inconsistent MSDeformAttn argument shapes")
    return N, S, M, D, L, Lq, P


def _check_step(batch, im2col_step):
    step = min(batch, int(im2col_step))
    if step <= 0 or batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")  # cu:52, :119


def _raise(rc, what):
    raise RuntimeError(f"{what} failed (code {rc}): {native.lib().datr_last_error().decode()}")


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    N, S, M, D, L, Lq, P = _check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    _check_step(N, im2col_step)
    lib = native.lib()
    with torch.cuda.device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        rc = lib.datr_msda_forward(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                   sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P,
                                   _DTYPES[value.dtype], out.data_ptr(), stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("fwd", (N, S, M, D, L, Lq, P, value.element_size()), e0, e1))
    if rc != 0:
        _raise(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    N, S, M, D, L, Lq, P = _check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                  extra=(("grad_output", grad_output),))
    _check_step(N, im2col_step)
    if grad_output.numel() != N * Lq * M * D:
        raise RuntimeError("grad_output has the wrong number of elements")
    lib = native.lib()
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)            # zero-filled by the library on the stream
        grad_loc = torch.empty_like(sampling_loc)
        grad_attn = torch.empty_like(attn_weight)
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        hs, hl = host_geometry(spatial_shapes, level_start_index)
        rc = lib.datr_msda_backward_hs(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), hs, hl,
                                    sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
                                    N, S, M, D, L, Lq, P, _DTYPES[value.dtype],
                                    grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
                                    stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("bwd", (N, S, M, D, L, Lq, P, value.element_size()), e0, e1))
    if rc != 0:
        _raise(rc, "ms_deform_attn_backward")
    return [grad_value, grad_loc, grad_attn]


def fused_config_supported(value, reference_points, L: int, P: int) -> bool:
    """True if datr_msda_fused_forward / _backward (include/datr_msda.h) cover a call with this value map
    [N,S,M,D], these reference points and L levels x P points."""
    if not (value.is_cuda and value.dtype == torch.float32 and value.dim() == 4 and value.shape[-1] == 32):
        return False
    return P in (1, 2, 4, 8) and L * P <= 32 and reference_points.shape[-1] in (2, 4) \
        and reference_points.dtype == torch.float32 and not reference_points.requires_grad


def fused_supported(value, sampling_offsets, reference_points) -> bool:
    """fused_config_supported for given sampling offsets [N,Lq,M,L,P,2]."""
    return fused_config_supported(value, reference_points, sampling_offsets.shape[3], sampling_offsets.shape[4])


def _row_stride(name, t):
    """Elements between consecutive queries of a [N, Lq, ...] tensor whose per-query block is contiguous (the tensor
    may be a column slice of a wider [N, Lq, R] GEMM output)."""
    inner, want = 1, []
    for d in reversed(t.shape[2:]):
        want.append(inner)
        inner *= d
    if list(t.stride()[2:]) != want[::-1] or t.stride(1) < inner or (t.shape[0] > 1 and t.stride(0) != t.shape[1] * t.stride(1)):
        raise RuntimeError(f"{name}: each query's block must be contiguous and the rows uniformly strided")
    return t.stride(1)


def _check_fused(value, spatial_shapes, level_start_index, offsets, logits, ref, extra=()):
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_offsets", offsets), ("attn_logits", logits), ("reference_points", ref), *extra]
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for name, t in named:
        if name not in ("sampling_offsets", "attn_logits") and not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda or t.device != value.device:
            raise RuntimeError(f"{name} must be a CUDA tensor on the device of value")
    for name, t in named[3:]:
        if t.dtype != torch.float32 or value.dtype != torch.float32:
            raise RuntimeError(f"fused MSDeformAttn is fp32 only ({name} is {t.dtype})")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64")
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    if offsets.dim() != 6 or ref.dim() != 4 or logits.dim() < 3:
        raise RuntimeError("expected sampling_offsets[N,Lq,M,L,P,2], attn_logits[N,Lq,M,L*P], reference_points[N,Lq,L,2|4]")
    Lq, P, R = offsets.shape[1], offsets.shape[4], ref.shape[-1]
    if tuple(offsets.shape) != (N, Lq, M, L, P, 2) or logits.numel() != N * Lq * M * L * P \
            or tuple(ref.shape) != (N, Lq, L, R) or R not in (2, 4) or level_start_index.numel() != L:
        raise RuntimeError("inconsistent fused MSDeformAttn argument shapes")
    if tuple(logits.shape[:2]) != (N, Lq):
        raise RuntimeError("attn_logits must be [N, Lq, ...]")
    return N, S, M, D, L, Lq, P, R, _row_stride("sampling_offsets", offsets), _row_stride("attn_logits", logits)


PAIR_STORAGE = {torch.bfloat16: 1, torch.float16: 2}     # include/datr_msda.h: DATR_STORE_BF16_PAIRS / _FP16_PAIRS


def pack_value_pairs(value, spatial_shapes, level_start_index, dtype=torch.bfloat16):
    """value [N,S,M,32] fp32 -> pair rows [N,S,M,64] of `dtype` (bf16 / fp16): per (pixel, head) line and per 16-byte
    lane segment, 4 channels of the pixel and the same 4 channels of its right-hand neighbour (include/datr_msda.h).
    No gradient flows through the result; the backward of the fused op still produces the fp32 grad_value."""
    if not (value.is_cuda and value.dtype == torch.float32 and value.dim() == 4 and value.shape[-1] == 32
            and value.is_contiguous()):
        raise RuntimeError("pack_value_pairs expects a contiguous CUDA fp32 value map [N,S,M,32]")
    if dtype not in PAIR_STORAGE:
        raise RuntimeError("pair rows are stored as torch.bfloat16 or torch.float16")
    N, S, M, _ = value.shape
    L = spatial_shapes.shape[0]
    with torch.cuda.device(value.device):
        pairs = torch.empty((N, S, M, 64), dtype=dtype, device=value.device)
        rc = native.lib().datr_msda_pack_value_pairs(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                                     N, S, M, L, PAIR_STORAGE[dtype], pairs.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        _raise(rc, "pack_value_pairs")
    return pairs


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                                 reference_points, pairs=None):
    """Extension (not in the reference's module): MSDeformAttn.forward's prologue + the op in one kernel.
    `pairs` (optional, from pack_value_pairs(value, ...)): gather from the 16-bit pair rows instead of `value`."""
    N, S, M, D, L, Lq, P, R, so, sl = _check_fused(value, spatial_shapes, level_start_index, sampling_offsets,
                                                   attn_logits, reference_points)
    if pairs is not None and (tuple(pairs.shape) != (N, S, M, 64) or pairs.dtype not in PAIR_STORAGE
                              or not pairs.is_contiguous() or pairs.device != value.device or D != 32):
        raise RuntimeError("pairs must be the [N,S,M,64] bf16 / fp16 tensor of pack_value_pairs(value, ...)")
    lib = native.lib()
    with torch.cuda.device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        if pairs is None:
            rc = lib.datr_msda_fused_forward(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                             sampling_offsets.data_ptr(), so, attn_logits.data_ptr(), sl,
                                             reference_points.data_ptr(), R, N, S, M, D, L, Lq, P, 0, out.data_ptr(),
                                             stream.cuda_stream)
        else:
            rc = lib.datr_msda_fused_forward_pairs(pairs.data_ptr(), PAIR_STORAGE[pairs.dtype], spatial_shapes.data_ptr(),
                                                   level_start_index.data_ptr(), sampling_offsets.data_ptr(), so,
                                                   attn_logits.data_ptr(), sl, reference_points.data_ptr(), R,
                                                   N, S, M, D, L, Lq, P, out.data_ptr(), stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("fwd", (N, S, M, D, L, Lq, P, value.element_size() if pairs is None else 2), e0, e1))
    if rc != 0:
        _raise(rc, "ms_deform_attn_fused_forward")
    return out


def ms_deform_attn_fused_backward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                                  reference_points, grad_output, merged_grad=None):
    """Returns [grad_value, grad_sampling_offsets, grad_attn_logits].  The two last gradients mirror the memory layout
    of their inputs; `merged_grad` (optional, [N, Lq, R] with R = the common row stride) is the buffer they are
    carved from when offsets and logits are column slices [0, 2T) and [2T, 3T) of one [N, Lq, R] GEMM output."""
    N, S, M, D, L, Lq, P, R, so, sl = _check_fused(value, spatial_shapes, level_start_index, sampling_offsets,
                                                   attn_logits, reference_points, extra=(("grad_output", grad_output),))
    if grad_output.numel() != N * Lq * M * D:
        raise RuntimeError("grad_output has the wrong number of elements")
    lib = native.lib()
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)            # zero-filled by the library on the stream
        T = M * L * P
        if merged_grad is not None:
            if so != sl or tuple(merged_grad.shape) != (N, Lq, so) or not merged_grad.is_contiguous() or so < 3 * T:
                raise RuntimeError("merged_grad must be a contiguous [N, Lq, row_stride] buffer")
            grad_off = merged_grad[..., :2 * T].view(sampling_offsets.shape)
            grad_logits = merged_grad[..., 2 * T:3 * T].view(attn_logits.shape)
        else:
            grad_off = torch.empty(sampling_offsets.shape, dtype=value.dtype, device=value.device)
            grad_logits = torch.empty(attn_logits.shape, dtype=value.dtype, device=value.device)
        go, gl = _row_stride("grad_offsets", grad_off), _row_stride("grad_logits", grad_logits)
        if merged_grad is None and (so != go or sl != gl):   # strided inputs, dense gradients: not expressible
            raise RuntimeError("strided sampling_offsets / attn_logits need merged_grad")
        stream = torch.cuda.current_stream()
        if _timers is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        hs, hl = host_geometry(spatial_shapes, level_start_index)
        rc = lib.datr_msda_fused_backward_hs(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), hs, hl,
                                          sampling_offsets.data_ptr(), so, attn_logits.data_ptr(), sl,
                                          reference_points.data_ptr(), R, grad_output.data_ptr(),
                                          N, S, M, D, L, Lq, P, 0, grad_value.data_ptr(), grad_off.data_ptr(),
                                          grad_logits.data_ptr(), stream.cuda_stream)
        if _timers is not None:
            e1.record(stream)
            _timers.append(("bwd", (N, S, M, D, L, Lq, P, value.element_size()), e0, e1))
    if rc != 0:
        _raise(rc, "ms_deform_attn_fused_backward")
    return [grad_value, grad_off, grad_logits]


def install(name: str = "MultiScaleDeformableAttention"):
    """Register this module under the reference's extension name."""
    sys.modules[name] = sys.modules[__name__]
    return sys.modules[__name__]
