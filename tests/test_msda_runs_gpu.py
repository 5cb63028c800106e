"""GPU parity of the MSDeformAttn RUN kernels (csrc/msda.cu: msda_fwd_runs_f32_d32 / msda_bwd_runs_f32_d32;
include/datr_msda.h: datr_msda_set_strategy), the kernels the encoder self-attention calls take by default.

A group of 8 lanes walks consecutive queries of one (head, level, point) and keeps the 2x2 footprint in registers
(same anchor: nothing moves; anchor + 1: the right column becomes the left one; else reload / flush), so these tests
feed it every footprint transition: run-coherent locations (the reference's initial offset pattern,
ops/modules/ms_deform_attn.py:62-76: point k of head m sits k pixels along direction m on EVERY level), the same with
sub-pixel jitter, fully random / out-of-range / exact-integer locations, levels one pixel wide or high, query counts
that end inside a run, batch boundaries.  Oracles: the C restatement of the reference CUDA arithmetic
(oracle/msda_oracle.c) for the op, the reference module prologue in fp64 + core_torch for the fused entry points;
and the row kernels (strategy 1), an independent code path for the same arithmetic."""
import numpy as np
import pytest
import torch

import msda_cases as mc
from oracle import msda as om
from test_msda_fused_gpu import make as make_fused, oracle as fused_oracle

pytestmark = pytest.mark.gpu
TIGHT = 5e-5


@pytest.fixture()
def MSDA():
    assert torch.cuda.is_available()
    from datr_b200 import MultiScaleDeformableAttention as M
    yield M
    M.set_strategy(0)


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
    ref = mc.encoder_reference_points(levels)[np.arange(Lq) % S]
    inv = np.array([[1.0 / w, 1.0 / h] for h, w in levels])[None, None, None, :, None, :]
    off = init_pattern_offsets(N, Lq, M, L, P, jitter, rng)
    loc = ref[None, :, None, None, None, :] + off * inv
    logits = rng.standard_normal((N, Lq, M, L * P))
    e = np.exp(logits - logits.max(-1, keepdims=True))
    return dict(value=rng.standard_normal((N, S, M, 32)).astype(np.float32), shapes=np.array(levels, dtype=np.int64),
                level_start=mc.level_start_index(levels), loc=loc.astype(np.float32),
                attn=(e / e.sum(-1, keepdims=True)).reshape(N, Lq, M, L, P).astype(np.float32),
                grad_out=rng.standard_normal((N, Lq, M * 32)).astype(np.float32))


def run_op(MSDA, inp):
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in inp.items()}
    out = MSDA.ms_deform_attn_forward(d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"], 64)
    gv, gl, ga = MSDA.ms_deform_attn_backward(d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"], d["grad_out"], 64)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in (out, gv, gl, ga)]


def oracle_op(inp):
    return [om.fwd(inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["attn"]),
            *om.bwd(inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["attn"], inp["grad_out"])]


LEVELS = {
    "l4": [(12, 17), (6, 9), (3, 5), (2, 3)],
    "l5": [(16, 22), (8, 11), (4, 6), (2, 3), (1, 2)],
    "thin": [(9, 1), (1, 7), (1, 1), (5, 2)],          # one pixel wide / high / single pixel levels
    "wide": [(3, 70), (2, 35), (1, 18), (1, 9)],        # long rows: runs of +1 steps
}


@pytest.mark.parametrize("jitter", [0.0, 0.02, 0.6], ids=["init", "subpixel", "halfpixel"])
@pytest.mark.parametrize("lv", list(LEVELS))
def test_run_kernels_on_coherent_locations_match_oracle_and_row_kernels(MSDA, lv, jitter):
    inp = coherent_inputs(2, 8, 4, LEVELS[lv], jitter, seed=list(LEVELS).index(lv) * 10 + int(jitter * 100))
    want = oracle_op(inp)
    MSDA.set_strategy(2)
    got = run_op(MSDA, inp)
    MSDA.set_strategy(1)
    rows = run_op(MSDA, inp)
    for g, r, w, key in zip(got, rows, want, ("out", "grad_value", "grad_loc", "grad_attn")):
        assert mc.rel_err(g, w) < TIGHT, f"{key} vs oracle"
        assert mc.rel_err(g, r) < TIGHT, f"{key} vs row kernels"


@pytest.mark.parametrize("mode", ["uniform", "outside", "integer", "encoder"])
@pytest.mark.parametrize("Lq", [1, 31, 64, 65, 203])
def test_run_kernels_forced_on_arbitrary_queries(MSDA, mode, Lq):
    """Strategy 2 on decoder-like calls: no run coherence, run ends inside a tile, two batches."""
    inp = mc.make_inputs(2, 8, 32, Lq, 4, LEVELS["l4"], mode, 300 + Lq, np.float32)
    want = oracle_op(inp)
    MSDA.set_strategy(2)
    got = run_op(MSDA, inp)
    for g, w, key in zip(got, want, ("out", "grad_value", "grad_loc", "grad_attn")):
        assert mc.rel_err(g, w) < TIGHT, key


def test_auto_strategy_picks_run_kernels_only_for_pixel_queries(MSDA):
    assert MSDA.get_strategy() == 0
    inp = coherent_inputs(1, 8, 4, LEVELS["l4"], 0.3, 5)
    auto = run_op(MSDA, inp)
    MSDA.set_strategy(2)
    forced = run_op(MSDA, inp)
    assert np.array_equal(auto[0], forced[0])            # forward is deterministic: same kernel => same bits
    with pytest.raises(RuntimeError):
        MSDA.set_strategy(7)


@pytest.mark.parametrize("lv,N", [("l4", 2), ("l5", 1), ("thin", 2)])
@pytest.mark.parametrize("spread", [0.0, 0.4, 3.0])
def test_fused_run_kernels_match_reference_prologue_oracle(MSDA, lv, N, spread):
    """Encoder mode of the fused entry points (2-d reference points = pixel centres, Lq == S) on the run kernels."""
    levels = LEVELS[lv]
    inp = make_fused(N, 8, -1, 4, levels, 2, 77, spread=spread)
    if spread == 0.0:       # exactly the initial pattern
        rng = np.random.default_rng(1)
        inp["offsets"] = init_pattern_offsets(N, inp["offsets"].shape[1], 8, len(levels), 4, 0.0, rng).astype(np.float32)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    want = fused_oracle(inp, 4)
    res = {}
    for st in (2, 1):
        MSDA.set_strategy(st)
        out = MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
        gv, goff, glg = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"],
                                                           d["logits"], d["ref"], d["grad_out"])
        torch.cuda.synchronize()
        res[st] = [t.cpu().numpy() for t in (out, gv, goff, glg)]
    for g, r, w, key in zip(res[2], res[1], want, ("out", "grad_value", "grad_offsets", "grad_logits")):
        assert mc.rel_err(g.reshape(w.shape), w) < TIGHT, f"{key} vs oracle"
        assert mc.rel_err(g, r) < TIGHT, f"{key} vs row kernels"


def test_fused_run_kernels_with_strided_merged_rows(MSDA):
    """Offsets and logits as column slices of one [rows, 384] tensor (the merged projection of the module)."""
    levels = LEVELS["l4"]
    inp = make_fused(2, 8, -1, 4, levels, 2, 78, spread=0.5)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    N, Lq = inp["offsets"].shape[:2]
    merged = torch.cat([d["offsets"].reshape(N, Lq, 256), d["logits"].reshape(N, Lq, 128)], -1).contiguous()
    off_v, lg_v = merged[..., :256].view(N, Lq, 8, 4, 4, 2), merged[..., 256:].view(N, Lq, 8, 16)
    MSDA.set_strategy(2)
    a = MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
    b = MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], off_v, lg_v, d["ref"])
    assert torch.equal(a, b)
    ga = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"], d["grad_out"])
    gb = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], off_v, lg_v, d["ref"], d["grad_out"])
    torch.cuda.synchronize()
    assert torch.equal(ga[1].reshape(-1), gb[1].reshape(-1)) and torch.equal(ga[2].reshape(-1), gb[2].reshape(-1))
    assert mc.rel_err(ga[0].cpu().numpy(), gb[0].cpu().numpy()) < TIGHT


@pytest.mark.parametrize("levels,N", [(mc.CFG2_LEVELS, 2), (mc.CFG4_LEVELS, 1)], ids=["cfg2", "cfg4_5scale"])
@pytest.mark.parametrize("jitter", [0.0, 0.5])
def test_full_size_encoder_runs_vs_rows_and_properties(MSDA, levels, N, jitter):
    """BASELINE.json configs[1] / [3] encoder shapes: run kernels against the row kernels, plus linearity in value
    (size-independent property: out(value_a + value_b) = out(value_a) + out(value_b))."""
    inp = coherent_inputs(N, 8, 4, levels, jitter, seed=9)
    MSDA.set_strategy(2)
    got = run_op(MSDA, inp)
    MSDA.set_strategy(1)
    rows = run_op(MSDA, inp)
    for g, r, key in zip(got, rows, ("out", "grad_value", "grad_loc", "grad_attn")):
        assert np.isfinite(g).all(), key
        assert mc.rel_err(g, r) < TIGHT, key
    MSDA.set_strategy(2)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in inp.items()}
    vb = torch.randn_like(d["value"])
    f = lambda v: MSDA.ms_deform_attn_forward(v, d["shapes"], d["level_start"], d["loc"], d["attn"], 64)
    lhs, rhs = f(d["value"] + vb), f(d["value"]) + f(vb)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-5
