"""oracle/msda.py -- TEST INFRASTRUCTURE ONLY (never imported by datr_b200/).

CPU oracle for multi-scale deformable attention.  Two independent restatements:

* ``fwd`` / ``bwd``: ctypes front-end of ``msda_oracle.c`` (plain C, OpenMP), the
  scalar arithmetic of the reference CUDA kernels
  (models/dino/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-159, :237-403).
* ``core_torch``: the ``grid_sample`` formulation of the reference's only CPU
  path, ``ms_deform_attn_core_pytorch`` (models/dino/ops/functions/
  ms_deform_attn_func.py:41-61); differentiable, so autograd gives the
  gradient oracle.  It is also the "port" that bench.py times as cpu_baseline.

Parity pinning: both are checked against outputs and autograd gradients of the
reference's own function, generated in the build container by
tests/golden/make_golden.py and committed under tests/golden/.
Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference legs).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "msda_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_OUT_DIR, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile msda_oracle.c with gcc (-O2, OpenMP) into oracle/_build/."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC))
    if force or stale:
        # -ffp-contract=off: keep the reference's mul/add sequence (no FMA fusion)
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
               "-o", _LIB_PATH, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _np(x, dtype=None):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    x = np.ascontiguousarray(x)
    return x.astype(dtype, copy=False) if dtype is not None else x


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, level_start, loc, attn):
    value = _np(value)
    assert value.dtype in (np.float32, np.float64)
    dt = value.dtype
    loc, attn = _np(loc, dt), _np(attn, dt)
    shapes = _np(shapes, np.int64).reshape(-1, 2)
    if level_start is None:
        hw = shapes[:, 0] * shapes[:, 1]
        level_start = np.concatenate([[0], np.cumsum(hw)[:-1]])
    level_start = _np(level_start, np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    assert attn.shape == (N, Lq, M, L, P) and shapes.shape[0] == L
    return value, shapes, level_start, loc, attn, (N, S, M, D, L, Lq, P)


def fwd(value, shapes, level_start, loc, attn) -> np.ndarray:
    """out[N,Lq,M*D] (numpy, dtype of value)."""
    value, shapes, level_start, loc, attn, dims = _prep(value, shapes, level_start, loc, attn)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=value.dtype)
    fn = getattr(_load(), "msda_oracle_fwd_f32" if value.dtype == np.float32 else "msda_oracle_fwd_f64")
    fn(_ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(attn),
       *map(ctypes.c_int, dims), _ptr(out))
    return out


def bwd(value, shapes, level_start, loc, attn, grad_out):
    """(grad_value, grad_loc, grad_attn) as numpy arrays."""
    value, shapes, level_start, loc, attn, dims = _prep(value, shapes, level_start, loc, attn)
    grad_out = _np(grad_out, value.dtype)
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(attn)
    fn = getattr(_load(), "msda_oracle_bwd_f32" if value.dtype == np.float32 else "msda_oracle_bwd_f64")
    fn(_ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(attn), _ptr(grad_out),
       *map(ctypes.c_int, dims), _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga


def core_torch(value: torch.Tensor, shapes, loc: torch.Tensor, attn: torch.Tensor) -> torch.Tensor:
    """grid_sample formulation (func.py:41-61): bilinear, zeros padding,
    align_corners=False on grid 2*loc-1, then the attention-weighted sum over
    (level, point).  Returns [N, Lq, M*D]."""
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = [(int(h), int(w)) for h, w in (shapes.tolist() if hasattr(shapes, "tolist") else shapes)]
    # heads become the batch of grid_sample: [N*M, D, S]
    per_head = value.permute(0, 2, 3, 1).reshape(N * M, D, S)
    grid = (loc * 2 - 1).permute(0, 2, 1, 3, 4, 5).reshape(N * M, Lq, L, P, 2)
    taps, start = [], 0
    for lvl, (h, w) in enumerate(hw):
        fmap = per_head[:, :, start:start + h * w].reshape(N * M, D, h, w)
        start += h * w
        taps.append(F.grid_sample(fmap, grid[:, :, lvl], mode="bilinear",
                                  padding_mode="zeros", align_corners=False))  # [N*M, D, Lq, P]
    sampled = torch.stack(taps, dim=3).reshape(N * M, D, Lq, L * P)
    wts = attn.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)
    out = (sampled * wts).sum(-1)                                              # [N*M, D, Lq]
    return out.reshape(N, M * D, Lq).transpose(1, 2).contiguous()


def module_prologue_torch(offsets: torch.Tensor, logits: torch.Tensor, reference_points: torch.Tensor, shapes, P: int):
    """The elementwise prologue of the reference's MSDeformAttn.forward (ops/modules/ms_deform_attn.py:99-111):
    attention weights = softmax of the logits over levels*points (:100-101); sampling locations = reference point +
    offset / (W_l, H_l) for 2-d reference points (:102-105) or reference centre + offset / n_points * box size * 0.5
    for 4-d reference boxes (:106-108).  offsets [N,Lq,M,L,P,2], logits [N,Lq,M,L*P], reference_points [N,Lq,L,2|4];
    returns (sampling_locations [N,Lq,M,L,P,2], attention_weights [N,Lq,M,L,P]).  Oracle of the fused entry points."""
    N, Lq, M, L = offsets.shape[:4]
    attn = torch.softmax(logits, -1).view(N, Lq, M, L, P)
    hw = torch.as_tensor(shapes.tolist() if hasattr(shapes, "tolist") else shapes, device=offsets.device)
    if reference_points.shape[-1] == 2:
        normalizer = torch.stack([hw[..., 1], hw[..., 0]], -1).to(offsets.dtype)                # (W_l, H_l), :103
        loc = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:
        loc = reference_points[:, :, None, :, None, :2] + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError("Last dim of reference_points must be 2 or 4")                         # :109-111
    return loc, attn


def fused_torch(value, shapes, offsets, logits, reference_points, P):
    """module_prologue_torch followed by core_torch: what datr_msda_fused_forward computes."""
    loc, attn = module_prologue_torch(offsets, logits, reference_points, shapes, P)
    return core_torch(value, shapes, loc, attn)
