"""Seeded synthetic inputs for the MSDeformAttn parity tests.

numpy's PCG64 stream is stable across numpy/torch versions, so the same inputs
are reproduced in the build container (where tests/golden/make_golden.py feeds
them to the reference) and on the GPU box (where /root/reference is absent).
"""
from __future__ import annotations

import numpy as np

# (name, N, M, D, Lq, P, levels[(H,W)...], loc_mode, seed)
SMALL_CASES = [
    ("testpy_shape",      1, 2, 2,  2, 2, [(6, 4), (3, 2)],                          "uniform", 11),
    ("d32_l4",            2, 8, 32, 37, 4, [(7, 9), (4, 5), (2, 3), (1, 2)],         "uniform", 12),
    ("d32_l4_outside",    1, 8, 32, 29, 4, [(7, 9), (4, 5), (2, 3), (1, 2)],         "outside", 13),
    ("d32_l4_integer",    1, 8, 32, 31, 4, [(8, 8), (4, 4), (2, 2), (1, 1)],         "integer", 14),
    ("d32_l5",            1, 8, 32, 41, 4, [(6, 7), (3, 4), (2, 2), (1, 2), (1, 1)], "encoder", 15),
    ("d32_l4_encoder",    2, 8, 32, -1, 4, [(6, 8), (3, 4), (2, 2), (1, 1)],         "encoder", 16),
    ("d30_generic",       1, 2, 30, 5, 2, [(6, 4), (3, 2)],                          "uniform", 17),
    ("d64_generic",       1, 2, 64, 5, 2, [(6, 4), (3, 2)],                          "outside", 18),
    ("d71_generic",       1, 2, 71, 3, 2, [(6, 4), (3, 2)],                          "uniform", 19),
    ("d16_m4_p3",         2, 4, 16, 9, 3, [(7, 5), (4, 3), (2, 2)],                  "outside", 20),
    ("single_pixel_lvls", 1, 8, 32, 6, 4, [(1, 1), (1, 1)],                          "outside", 21),
]

CFG1_LEVELS = [(100, 100), (50, 50), (25, 25), (13, 13)]           # 800x800, S=13294
CFG2_LEVELS = [(100, 167), (50, 84), (25, 42), (13, 21)]           # 1333x800, S=22223
CFG4_LEVELS = [(200, 334), (100, 167), (50, 84), (25, 42), (13, 21)]  # 5-scale, S=89023


def level_start_index(levels):
    hw = np.array([h * w for h, w in levels], dtype=np.int64)
    return np.concatenate([[0], np.cumsum(hw)[:-1]]).astype(np.int64)


def encoder_reference_points(levels):
    """Per-token centres ((x+0.5)/W, (y+0.5)/H) for all levels, as
    get_reference_points gives for valid_ratio 1 (deformable_transformer.py:477-489)."""
    pts = []
    for h, w in levels:
        ys, xs = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
        pts.append(np.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    return np.concatenate(pts, 0)  # [S, 2]


def make_inputs(N, M, D, Lq, P, levels, loc_mode="uniform", seed=0, dtype=np.float32,
                value_scale=1.0):
    """Returns dict(value, shapes, level_start, loc, attn, grad_out)."""
    rng = np.random.default_rng(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    if Lq < 0:
        Lq = S
    value = (rng.standard_normal((N, S, M, D)) * value_scale).astype(dtype)
    if loc_mode == "uniform":
        loc = rng.random((N, Lq, M, L, P, 2))
    elif loc_mode == "outside":            # ~25 % of samples outside [0,1], some far outside
        loc = rng.random((N, Lq, M, L, P, 2)) * 1.5 - 0.25
        far = rng.random((N, Lq, M, L, P, 1)) < 0.03
        loc = np.where(far, loc * 9.0 - 4.0, loc)
    elif loc_mode == "integer":            # exact pixel centres / edges / half-integers
        hw = np.array([[w, h] for h, w in levels], dtype=np.float64)[None, None, None, :, None, :]
        k = rng.integers(-3, 12, size=(N, Lq, M, L, P, 2)).astype(np.float64)
        # pixel coord = k/2 - 0.5.  k == -1 (coord exactly -1) is excluded: there the reference's
        # CUDA kernel drops the sample (guard `> -1`, cuh:288) while grid_sample's backward still
        # reports a one-sided d/dloc -- the two reference paths disagree on that measure-zero set.
        k = np.where(k == -1, -3, k)
        loc = (k * 0.5) / hw
    elif loc_mode == "encoder":            # reference grid + small learned-offset-like jitter
        ref = encoder_reference_points(levels)
        ref = ref[np.arange(Lq) % S][None, :, None, None, None, :]
        inv = np.array([[1.0 / w, 1.0 / h] for h, w in levels])[None, None, None, :, None, :]
        loc = ref + rng.standard_normal((N, Lq, M, L, P, 2)) * 2.0 * inv
    else:
        raise ValueError(loc_mode)
    logits = rng.standard_normal((N, Lq, M, L * P))
    e = np.exp(logits - logits.max(-1, keepdims=True))
    attn = (e / e.sum(-1, keepdims=True)).reshape(N, Lq, M, L, P)
    grad_out = rng.standard_normal((N, Lq, M * D))
    return dict(
        value=value,
        shapes=np.array(levels, dtype=np.int64),
        level_start=level_start_index(levels),
        loc=loc.astype(dtype),
        attn=attn.astype(dtype),
        grad_out=grad_out.astype(dtype),
    )


def init_pattern_offsets(N, Lq, M, L, P, jitter, rng):
    """Offsets in pixels as MSDeformAttn._reset_parameters leaves them (weight 0, bias = direction grid), + jitter."""
    th = np.arange(M) * (2.0 * np.pi / M)
    grid = np.stack([np.cos(th), np.sin(th)], -1)
    grid = grid / np.abs(grid).max(-1, keepdims=True)
    off = np.tile(grid[:, None, None, :], (1, L, P, 1)) * (np.arange(P) + 1)[None, None, :, None]
    off = np.broadcast_to(off[None, None], (N, Lq, M, L, P, 2)).copy()
    return off + rng.standard_normal(off.shape) * jitter


def coherent_inputs(N, M, P, levels, jitter, seed, Lq=-1):
    rng = np.random.default_rng(seed)
    L, S = len(levels), sum(h * w for h, w in levels)
    if Lq < 0:
        Lq = S
    ref = encoder_reference_points(levels)[np.arange(Lq) % S]
    inv = np.array([[1.0 / w, 1.0 / h] for h, w in levels])[None, None, None, :, None, :]
    off = init_pattern_offsets(N, Lq, M, L, P, jitter, rng)
    loc = ref[None, :, None, None, None, :] + off * inv
    logits = rng.standard_normal((N, Lq, M, L * P))
    e = np.exp(logits - logits.max(-1, keepdims=True))
    return dict(value=rng.standard_normal((N, S, M, 32)).astype(np.float32), shapes=np.array(levels, dtype=np.int64),
                level_start=level_start_index(levels), loc=loc.astype(np.float32),
                attn=(e / e.sum(-1, keepdims=True)).reshape(N, Lq, M, L, P).astype(np.float32),
                grad_out=rng.standard_normal((N, Lq, M * 32)).astype(np.float32))


def small_case(name, dtype=np.float32):
    for c in SMALL_CASES:
        if c[0] == name:
            _, N, M, D, Lq, P, levels, mode, seed = c
            return make_inputs(N, M, D, Lq, P, levels, mode, seed, dtype)
    raise KeyError(name)


def rel_err(a, b):
    """max|a-b| / max|b|  -- the per-tensor relative error of SURVEY.md §8c."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
