"""GPU parity of the module-level fused MSDeformAttn entry points (datr_msda_fused_forward / _backward,
include/datr_msda.h) and of the padding-mask kernel (include/datr_rowmask.h).

Oracle: the reference's prologue (ops/modules/ms_deform_attn.py:99-111: softmax of the logits, locations from
reference points + offsets) written with torch in fp64 on the CPU, feeding the oracle's `core_torch`
(ms_deform_attn_func.py:41-61); autograd through that composition supplies the gradients of value, offsets and
logits.  Bar: fp32 1e-3 relative per tensor (north_star), regression guard 5e-5."""
import numpy as np
import pytest
import torch

import msda_cases as mc
from oracle import msda as om

pytestmark = pytest.mark.gpu

TIGHT = 5e-5


def make(N, M, Lq, P, levels, ref_dim, seed, spread=3.0):
    rng = np.random.default_rng(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    if Lq < 0:
        Lq = S
    value = rng.standard_normal((N, S, M, 32)).astype(np.float32)
    offsets = (rng.standard_normal((N, Lq, M, L, P, 2)) * spread).astype(np.float32)
    logits = (rng.standard_normal((N, Lq, M, L * P)) * 2).astype(np.float32)
    if ref_dim == 2:
        if Lq == S:
            ref = np.broadcast_to(mc.encoder_reference_points(levels)[None, :, None, :], (N, Lq, L, 2))
        else:
            ref = rng.random((N, Lq, L, 2)) * 1.2 - 0.1            # a few centres outside the map
    else:
        ref = np.concatenate([rng.random((N, Lq, L, 2)), rng.random((N, Lq, L, 2)) * 0.6 + 0.01], -1)
    grad_out = rng.standard_normal((N, Lq, M * 32)).astype(np.float32)
    shapes = np.array(levels, dtype=np.int64)
    return dict(value=value, offsets=offsets, logits=logits, ref=np.ascontiguousarray(ref, dtype=np.float32),
                grad_out=grad_out, shapes=shapes, level_start=mc.level_start_index(levels))


def oracle(inp, P):
    v, off, lg = (torch.from_numpy(inp[k]).double().requires_grad_(True) for k in ("value", "offsets", "logits"))
    out = om.fused_torch(v, inp["shapes"], off, lg, torch.from_numpy(inp["ref"]).double(), P)
    out.backward(torch.from_numpy(inp["grad_out"]).double().view_as(out))
    return [t.detach().numpy() for t in (out, v.grad, off.grad, lg.grad)]


CASES = [
    ("enc_l4", 2, 8, -1, 4, [(9, 12), (5, 6), (3, 3), (2, 2)], 2, 41),
    ("enc_l5", 1, 8, -1, 4, [(8, 11), (4, 6), (2, 3), (1, 2), (1, 1)], 2, 42),
    ("dec_ref4", 2, 8, 77, 4, [(9, 12), (5, 6), (3, 3), (2, 2)], 4, 43),
    ("dec_ref2", 1, 8, 33, 4, [(7, 9), (4, 5), (2, 3), (1, 2)], 2, 44),
    ("p8_l4", 1, 4, 19, 8, [(7, 9), (4, 5), (2, 3), (1, 2)], 4, 45),
    ("p1_l3", 1, 8, 21, 1, [(7, 9), (4, 5), (2, 3)], 2, 46),
    ("p2_l2", 2, 2, 13, 2, [(6, 4), (3, 2)], 4, 47),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_entry_points_match_the_reference_prologue_plus_oracle(case):
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    _, N, M, Lq, P, levels, ref_dim, seed = case
    inp = make(N, M, Lq, P, levels, ref_dim, seed)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    assert MSDA.fused_supported(d["value"], d["offsets"], d["ref"])
    out = MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
    gv, goff, glg = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"],
                                                       d["logits"], d["ref"], d["grad_out"])
    torch.cuda.synchronize()
    want = oracle(inp, P)
    for g, w, key in zip((out, gv, goff, glg), want, ("out", "grad_value", "grad_offsets", "grad_logits")):
        assert g.shape == w.shape or g.numel() == w.size, key
        assert mc.rel_err(g.cpu().numpy().reshape(w.shape), w) < TIGHT, key


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_fused_equals_unfused_cuda_composition(ref_dim):
    """Same CUDA kernels with and without the in-kernel prologue: the fused path must reproduce the op fed with
    torch-computed locations / softmax (identical operation order => differences at rounding level only)."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    levels = [(20, 31), (10, 16), (5, 8), (3, 4)]
    inp = make(2, 8, 333 if ref_dim == 4 else -1, 4, levels, ref_dim, 50 + ref_dim, spread=4.0)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    off = d["offsets"].clone().requires_grad_(True); lg = d["logits"].clone().requires_grad_(True)
    N, Lq, M, L, P = off.shape[:5]
    attn = torch.softmax(lg, -1).view(N, Lq, M, L, P)
    if ref_dim == 2:
        loc = d["ref"][:, :, None, :, None, :] + off / d["shapes"].flip(-1)[None, None, None, :, None, :]
    else:
        loc = d["ref"][:, :, None, :, None, :2] + off / P * d["ref"][:, :, None, :, None, 2:] * 0.5
    out_u = MSDA.ms_deform_attn_forward(d["value"], d["shapes"], d["level_start"], loc.detach().contiguous(),
                                        attn.detach().contiguous(), 64)
    gv_u, gl_u, ga_u = MSDA.ms_deform_attn_backward(d["value"], d["shapes"], d["level_start"], loc.detach().contiguous(),
                                                    attn.detach().contiguous(), d["grad_out"], 64)
    torch.autograd.backward([loc, attn], [gl_u, ga_u])
    out_f = MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
    gv_f, goff_f, glg_f = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"],
                                                             d["logits"], d["ref"], d["grad_out"])
    for a, b, key in ((out_f, out_u, "out"), (gv_f, gv_u, "gv"), (goff_f, off.grad, "goff"), (glg_f, lg.grad, "glg")):
        assert mc.rel_err(a.cpu().numpy(), b.cpu().numpy().reshape(a.shape)) < 1e-5, key


def test_merged_projection_layout_matches_separate_tensors():
    """Offsets and logits as column slices of one [N, Lq, 3*M*L*P] GEMM output (row-strided views) and their gradients
    carved from one merged buffer: bit-identical to the densely packed call (same kernels, same arithmetic)."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    from datr_b200.models.dino.ops.functions import MSDeformAttnMergedFunction
    levels = [(9, 12), (5, 6), (3, 3), (2, 2)]
    inp = make(2, 8, 77, 4, levels, 4, 61)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    N, Lq, M, L, P = d["offsets"].shape[:5]
    T = M * L * P
    merged = torch.cat((d["offsets"].reshape(N, Lq, 2 * T), d["logits"].reshape(N, Lq, T)), -1).contiguous()
    out_d = MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
    gv_d, go_d, gl_d = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"],
                                                          d["logits"], d["ref"], d["grad_out"])
    v = d["value"].clone().requires_grad_(True); mg = merged.clone().requires_grad_(True)
    out_m = MSDeformAttnMergedFunction.apply(v, d["shapes"], d["level_start"], mg, d["ref"], M, L, P)
    out_m.backward(d["grad_out"])
    assert torch.equal(out_m, out_d)
    assert torch.equal(mg.grad[..., :2 * T].reshape(go_d.shape), go_d) and torch.equal(mg.grad[..., 2 * T:].reshape(gl_d.shape), gl_d)
    assert mc.rel_err(v.grad.cpu().numpy(), gv_d.cpu().numpy()) < 1e-5        # atomics: order-dependent rounding


@pytest.mark.parametrize("gemm", ["fp32", "tf32"])
def test_module_fused_and_unfused_agree_with_padding_mask(gemm):
    """MSDeformAttn module: fused kernels + in-place padding mask vs the reference-shaped composition (masked_fill,
    torch softmax / location arithmetic, the plain op), in both GEMM modes (same GEMM kernels on both sides)."""
    from datr_b200 import linear as dl
    from datr_b200.models.dino.ops.modules import ms_deform_attn as mod
    dl.set_mode(gemm)
    levels = [(12, 17), (6, 9), (3, 5), (2, 3)]
    S = sum(h * w for h, w in levels)
    torch.manual_seed(7)
    m = mod.MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():   # move the offsets / logits off their symmetric initialisation
        m.sampling_offsets.weight.normal_(0, 0.05); m.attention_weights.weight.normal_(0, 0.05)
    shapes = torch.tensor(levels, dtype=torch.long, device="cuda")
    start = torch.from_numpy(mc.level_start_index(levels)).cuda()
    src = torch.randn(2, S, 256, device="cuda")
    ref = torch.from_numpy(mc.encoder_reference_points(levels)).float().cuda()[None, :, None, :].expand(2, S, 4, 2).contiguous()
    mask = torch.zeros(2, S, dtype=torch.bool, device="cuda"); mask[1, 150:] = True; mask[0, ::7] = True
    gout = torch.randn(2, S, 256, device="cuda")
    res = []
    for fused in (True, False):
        mod.set_fused(fused)
        m.zero_grad()
        x = src.clone().requires_grad_(True)
        y = m(x, ref, x, shapes, start, mask)
        y.backward(gout)
        res.append([y.detach(), x.grad] + [p.grad.clone() for p in m.parameters()])
    mod.set_fused(True)
    dl.set_mode("fp32")
    # fp32: the same arithmetic on both sides; tf32: the merged offsets+logits GEMM sums its K blocks in another order
    # than the two separate GEMMs, and the sampling positions amplify that rounding (bar for this mode: 1e-2)
    for a, b in zip(*res):
        assert mc.rel_err(a.cpu().numpy(), b.cpu().numpy()) < (1e-4 if gemm == "fp32" else 2e-3)
    # the reference composition with torch's own masked_fill on the same weights
    x = src.clone().requires_grad_(True)
    value = torch.nn.functional.linear(x, m.value_proj.weight, m.value_proj.bias).masked_fill(mask[..., None], 0.0)
    assert gemm == "tf32" or torch.equal(value[mask], torch.zeros_like(value[mask]))


def test_zero_masked_rows_matches_masked_fill():
    from datr_b200 import native
    from datr_b200.rowmask import zero_masked_rows
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(3, 1001, 256, generator=g).cuda()
    mask = (torch.rand(3, 1001, generator=g) < 0.3).cuda()
    w = torch.randn(3, 1001, 256, generator=g).cuda()
    a = x.clone().requires_grad_(True)
    n0 = native.rowmask_launch_count()
    ya = zero_masked_rows(a * 1.0, mask)
    (ya * w).sum().backward()
    assert native.rowmask_launch_count() == n0 + 2, "padding-mask kernel did not launch"
    b = x.clone().requires_grad_(True)
    yb = (b * 1.0).masked_fill(mask[..., None], 0.0)
    (yb * w).sum().backward()
    assert torch.equal(ya, yb) and torch.equal(a.grad, b.grad)
    # 4-d activation with a 2-d mask, ragged tail (rows not a multiple of 32)
    v = torch.randn(2, 77, 8, 32, generator=g).cuda()
    mk = (torch.rand(2, 77, generator=g) < 0.5).cuda()
    assert torch.equal(zero_masked_rows(v.clone(), mk), v.masked_fill(mk[..., None, None], 0.0))


def _torch_prologue(d, P):
    """ops/modules/ms_deform_attn.py:99-111 with torch ops on the device."""
    off, lg, ref = d["offsets"], d["logits"], d["ref"]
    N, Lq, M, L = off.shape[:4]
    attn = torch.softmax(lg, -1).view(N, Lq, M, L, P)
    if ref.shape[-1] == 2:
        loc = ref[:, :, None, :, None, :] + off / d["shapes"].flip(-1)[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + off / P * ref[:, :, None, :, None, 2:] * 0.5
    return loc.contiguous(), attn.contiguous()


@pytest.mark.parametrize("levels,tag", [(mc.CFG2_LEVELS, "config2 N=2 S=22223"), (mc.CFG4_LEVELS, "config4 5-scale N=1 S=89023")])
def test_full_size_fused_vs_reference_cuda_kernels_and_properties(levels, tag):
    """BASELINE.json configs[1] / configs[3] encoder calls through the fused entry points: against the reference's
    own CUDA kernels (oracle/_ref, when built) fed with the torch prologue, plus size-independent properties --
    linearity in value, zero grad_out -> zero grads, sum(grad_logits) == 0 per (query, head) (softmax), Euler identity
    <grad_value, value> == <grad_out, out>."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    N = 2 if len(levels) == 4 else 1
    inp = make(N, 8, -1, 4, levels, 2, 71 + len(levels), spread=3.0)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    f = lambda v: MSDA.ms_deform_attn_fused_forward(v, d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
    out = f(d["value"])
    gv, goff, glg = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"],
                                                       d["ref"], d["grad_out"])
    v2 = torch.randn_like(d["value"])
    assert mc.rel_err(f(d["value"] + 2 * v2).cpu().numpy(), (out + 2 * f(v2)).cpu().numpy()) < 1e-5
    rhs = (d["grad_out"].double() * out.double()).sum().item()
    assert abs((gv.double() * d["value"].double()).sum().item() - rhs) / abs(rhs) < 1e-4
    assert float(glg.double().sum(-1).abs().max()) < 1e-3 * float(glg.abs().max())
    zero = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"],
                                              torch.zeros_like(d["grad_out"]))
    assert all(not t.any().item() for t in zero)
    from oracle import build_ref
    ref = build_ref.load()
    if ref is None:
        return
    off = d["offsets"].clone().requires_grad_(True); lg = d["logits"].clone().requires_grad_(True)
    loc, attn = _torch_prologue(dict(d, offsets=off, logits=lg), 4)
    args = (d["value"], d["shapes"], d["level_start"], loc.detach(), attn.detach())
    out_ref = ref.ms_deform_attn_forward(*args, 64)
    gv_ref, gl_ref, ga_ref = ref.ms_deform_attn_backward(*args, d["grad_out"], 64)
    torch.autograd.backward([loc, attn], [gl_ref, ga_ref])
    assert mc.rel_err(out.cpu().numpy(), out_ref.cpu().numpy()) < TIGHT
    for a, b, key in ((gv, gv_ref, "gv"), (goff, off.grad, "goff"), (glg, lg.grad, "glg")):
        assert mc.rel_err(a.cpu().numpy(), b.cpu().numpy().reshape(a.shape)) < TIGHT * 5, key


def test_fused_error_behaviour():
    from datr_b200 import MultiScaleDeformableAttention as MSDA, native
    inp = make(1, 8, 9, 4, [(4, 5), (2, 3)], 2, 80)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_fused_forward(d["value"].cpu(), d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
    with pytest.raises(RuntimeError, match="uniformly strided|contiguous"):
        padded = torch.zeros(d["offsets"].shape[:-1] + (3,), device="cuda")[..., :2]      # same shape, gaps inside a row
        MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], padded, d["logits"], d["ref"])
    with pytest.raises(RuntimeError, match="fp32 only"):
        MSDA.ms_deform_attn_fused_forward(d["value"], d["shapes"], d["level_start"], d["offsets"].double(), d["logits"], d["ref"])
    assert not MSDA.fused_supported(d["value"].double(), d["offsets"], d["ref"])
    assert not MSDA.fused_supported(d["value"], d["offsets"], d["ref"].clone().requires_grad_(True))
    lib = native.lib()
    p = d["value"].data_ptr()
    rc = lib.datr_msda_fused_forward(p, p, p, p, 0, p, 0, p, 3, 1, 1, 1, 32, 1, 1, 4, 0, p, None)
    assert rc == -1 and b"2 or 4" in lib.datr_last_error()
    rc = lib.datr_msda_fused_forward(p, p, p, p, 0, p, 0, p, 2, 1, 1, 1, 16, 1, 1, 4, 0, p, None)
    assert rc == -4 and b"fused entry points cover" in lib.datr_last_error()
    rc = lib.datr_msda_fused_forward(p, p, p, p, 5, p, 0, p, 2, 1, 1, 8, 32, 4, 1, 4, 0, p, None)
    assert rc == -1 and b"row strides" in lib.datr_last_error()


# --------------------------------------------------------------------------------------------------------------------
# Backward scatter variants (include/datr_msda.h: datr_msda_set_backward_stages) and pair-row value maps
# --------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def scatter_variant():
    from datr_b200 import native
    lib = native.lib()
    yield lib.datr_msda_set_backward_stages
    lib.datr_msda_set_backward_stages(-1)


TMA_CASES = [c for c in CASES if c[4] == 4 and len(c[5]) <= 4 and all(h >= 2 and w >= 2 for h, w in c[5])]


@pytest.mark.parametrize("stages", [0, 1, 2])
@pytest.mark.parametrize("case", TMA_CASES, ids=[c[0] for c in TMA_CASES])
def test_backward_scatter_variants_match_the_oracle(case, stages, scatter_variant):
    """Vector reductions (0) and the TMA reduce with 1 / 2 staging buffers per warp: same sums, different order."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    _, N, M, Lq, P, levels, ref_dim, seed = case
    inp = make(N, M, Lq, P, levels, ref_dim, seed)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    scatter_variant(stages)
    assert MSDA.host_geometry(d["shapes"], d["level_start"])[0] is not None     # the TMA variants need host geometry
    gv, goff, glg = MSDA.ms_deform_attn_fused_backward(d["value"], d["shapes"], d["level_start"], d["offsets"],
                                                       d["logits"], d["ref"], d["grad_out"])
    torch.cuda.synchronize()
    want = oracle(inp, P)[1:]
    for g, w, key in zip((gv, goff, glg), want, ("grad_value", "grad_offsets", "grad_logits")):
        assert mc.rel_err(g.cpu().numpy().reshape(w.shape), w) < TIGHT, (key, stages)


@pytest.mark.parametrize("stages", [1, 2])
def test_tma_scatter_at_the_map_borders_and_beyond(stages, scatter_variant):
    """Samples far outside the map (rejected), exactly on its border rows / columns (two of the four corners dropped)
    and in the interior: the anchored 2x2 box of the TMA reduce carries zero weights where the reference skips a corner
    (cuh:56-79, :288).  Compared with the vector-reduction kernel on the same inputs and with the oracle."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    levels = [(6, 7), (3, 4), (2, 3), (2, 2)]
    inp = make(2, 8, 40, 4, levels, 2, 77, spread=9.0)
    inp["ref"][:, :10] = 0.0                   # corner of the map
    inp["ref"][:, 10:20] = 1.0
    inp["ref"][:, 20:25] = -3.0                # every sample rejected
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    args = (d["value"], d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"], d["grad_out"])
    scatter_variant(0)
    base = MSDA.ms_deform_attn_fused_backward(*args)
    scatter_variant(stages)
    got = MSDA.ms_deform_attn_fused_backward(*args)
    torch.cuda.synchronize()
    want = oracle(inp, 4)[1:]
    for g, b, w, key in zip(got, base, want, ("grad_value", "grad_offsets", "grad_logits")):
        assert mc.rel_err(g.cpu().numpy().reshape(w.shape), w) < TIGHT, key
        assert mc.rel_err(g.cpu().numpy(), b.cpu().numpy()) < 1e-5, key
    assert torch.equal(got[1], base[1]) and torch.equal(got[2], base[2])   # only grad_value's summation order changes


def test_op_backward_takes_the_tma_scatter_and_matches_the_vector_reductions(scatter_variant):
    """The reference-ABI op (materialised locations / weights) through datr_msda_backward_hs."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    levels = [(9, 12), (5, 6), (3, 3), (2, 2)]
    inp = mc.make_inputs(2, 8, 32, 50, 4, levels, "uniform", 5, np.float32)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    args = (d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"], d["grad_out"], 64)
    scatter_variant(0)
    base = MSDA.ms_deform_attn_backward(*args)
    scatter_variant(2)
    got = MSDA.ms_deform_attn_backward(*args)
    for g, b in zip(got, base):
        assert mc.rel_err(g.cpu().numpy(), b.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("dtype,bar", [(torch.bfloat16, 1e-2), (torch.float16, 1e-3)])
@pytest.mark.parametrize("case", [c for c in CASES if c[4] == 4], ids=[c[0] for c in CASES if c[4] == 4])
def test_pair_row_forward(case, dtype, bar):
    """Forward on 16-bit pair rows: (a) inside the storage type's precision class against the fp64 oracle on the fp32
    values (bf16: BASELINE's 1e-2 bar, fp16: 1e-3), (b) equal -- up to fp32 summation order -- to the fp32-row kernel
    on values rounded to the storage type, i.e. the rounding of the value map is the ONLY difference."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    _, N, M, Lq, P, levels, ref_dim, seed = case
    inp = make(N, M, Lq, P, levels, ref_dim, seed)
    d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    pairs = MSDA.pack_value_pairs(d["value"], d["shapes"], d["level_start"], dtype)
    assert pairs.shape == (N, d["value"].shape[1], M, 64) and pairs.dtype == dtype
    rest = (d["shapes"], d["level_start"], d["offsets"], d["logits"], d["ref"])
    got = MSDA.ms_deform_attn_fused_forward(d["value"], *rest, pairs=pairs)
    rounded = MSDA.ms_deform_attn_fused_forward(d["value"].to(dtype).float(), *rest)
    torch.cuda.synchronize()
    want = oracle(inp, P)[0]
    assert mc.rel_err(got.cpu().numpy().reshape(want.shape), want) < bar
    assert mc.rel_err(got.cpu().numpy(), rounded.cpu().numpy()) < 2e-6


def test_pair_rows_layout():
    """Lane j of a (pixel, head) line: channels 4j..4j+3 of the pixel, then of its right-hand neighbour (zeros in the
    last column of a level row)."""
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    levels = [(3, 4), (2, 2), (1, 3)]
    S = sum(h * w for h, w in levels)
    value = torch.randn(2, S, 4, 32, device="cuda")
    shapes = torch.tensor(levels, dtype=torch.int64, device="cuda")
    start = torch.from_numpy(mc.level_start_index(levels)).cuda()
    pairs = MSDA.pack_value_pairs(value, shapes, start, torch.bfloat16).float().view(2, S, 4, 8, 2, 4)
    want = value.bfloat16().float().view(2, S, 4, 8, 4)
    assert torch.equal(pairs[..., 0, :], want)
    right = torch.zeros_like(want)
    s0 = 0
    for h, w in levels:
        for y in range(h):
            a = s0 + y * w
            right[:, a:a + w - 1] = want[:, a + 1:a + w]
        s0 += h * w
    assert torch.equal(pairs[..., 1, :], right)
