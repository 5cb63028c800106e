"""GPU parity of the decoder self-attention path (datr_b200.attention: batched GEMMs around the in-place masked-softmax
kernels of include/datr_attn.h) against torch's own attention arithmetic in fp64 on the same inputs, and against
nn.MultiheadAttention as the reference's decoder layer calls it (models/dino/deformable_transformer.py:880-897).
Bar: fp32 1e-3 relative per tensor (north_star); measured ~1e-6 with fp32 GEMMs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def dn_mask(T, pad, groups, g):
    """The de-noising mask of dn_components.py:105-121: matching queries cannot see the de-noising part, de-noising
    groups cannot see each other (True = blocked)."""
    m = torch.zeros(T, T, dtype=torch.bool)
    m[pad:, :pad] = True
    single = pad // groups
    for i in range(groups):
        m[single * i:single * (i + 1), :single * i] = True
        m[single * i:single * (i + 1), single * (i + 1):pad] = True
    return m


@pytest.mark.parametrize("T,masked", [(1100, True), (900, False), (37, True), (1300, True)])
def test_matches_fp64_attention(T, masked):
    from datr_b200 import attention, native
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(T)
    q, k, v = (torch.randn(2, T, 8, 32, generator=g).cuda().transpose(1, 2).requires_grad_(True) for _ in range(3))
    go = torch.randn(2, 8, T, 32, generator=g).cuda()
    blocked = dn_mask(T, 200 if T > 300 else 20, 10, g).cuda() if masked else None
    assert attention.applicable(q, blocked, 0.0)
    n0 = native.attn_launch_count()
    o = attention.self_attention(q, k, v, blocked)
    o.backward(go)
    assert native.attn_launch_count() == n0 + 2
    got = [o.detach(), q.grad, k.grad, v.grad]
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    s = qd @ kd.transpose(-1, -2) / math.sqrt(32)
    if blocked is not None:
        s = s.masked_fill(blocked, float("-inf"))
    od = torch.softmax(s, -1) @ vd
    od.backward(go.double())
    for a, b, name in zip(got, [od.detach(), qd.grad, kd.grad, vd.grad], ("out", "dq", "dk", "dv")):
        assert rel(a, b) < 2e-5, name


def test_packed_self_attention_matches_nn_multihead_attention():
    """The decoder's PackedSelfAttention (own attention path) vs torch.nn.MultiheadAttention with the same weights."""
    from datr_b200.models.dino import deformable_transformer as dt
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    T, C = 300, 256
    ours = dt.PackedSelfAttention(C, 8, dropout=0.0).cuda()
    ref = torch.nn.MultiheadAttention(C, 8, dropout=0.0).cuda()
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(2, T, C, device="cuda"); pos = torch.randn(2, T, C, device="cuda")
    mask = dn_mask(T, 100, 5, None).cuda()
    res = []
    for own in (True, False):
        dt._OWN_ATTENTION = own
        xa = x.clone().requires_grad_(True)
        y = ours(xa + pos, xa, attn_mask=mask)
        y.sum().backward()
        res.append((y.detach(), xa.grad))
    dt._OWN_ATTENTION = True
    xb = x.clone().requires_grad_(True)
    qk = (xb + pos).transpose(0, 1)
    yr = ref(qk, qk, xb.transpose(0, 1), attn_mask=mask)[0].transpose(0, 1)
    yr.sum().backward()
    for y, gx in res:
        assert rel(y, yr.detach()) < 1e-4 and rel(gx, xb.grad) < 1e-4


# --------------------------------------------------------------------------------------------------------------------
# Fused tensor-core attention (csrc/attn_fused.cu)
# --------------------------------------------------------------------------------------------------------------------
def _fp64_attention(qk, v, H, blocked):
    N, T, C2 = qk.shape
    C, d = C2 // 2, C2 // 2 // H
    q, k = (t.reshape(N, T, H, d).transpose(1, 2) for t in (qk[..., :C], qk[..., C:]))
    vv = v.reshape(N, T, H, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(d)
    if blocked is not None:
        s = s.masked_fill(blocked, float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ vv).transpose(1, 2).reshape(N, T, C), p, torch.logsumexp(s, -1)


@pytest.mark.parametrize("N,H,T,masked", [(2, 8, 1100, True), (1, 8, 900, False), (2, 4, 37, True), (1, 8, 128, True),
                                           (1, 2, 129, False), (3, 8, 1300, True)])
def test_fused_forward_matches_fp64_attention(N, H, T, masked):
    """Output, log-sum-exp and probabilities of the one-kernel attention against fp64 torch arithmetic.  TF32 products
    (10-bit mantissa) on unit-variance inputs: bar 1e-2 (BASELINE's reduced-precision class), measured ~1e-3."""
    from datr_b200 import attention, native
    g = torch.Generator(device="cpu").manual_seed(T + H)
    C = 32 * H
    qk = torch.randn(N, T, 2 * C, generator=g).cuda()
    v = torch.randn(N, T, C, generator=g).cuda()
    blocked = dn_mask(T, 200 if T > 300 else 20, 10, g).cuda() if masked else None
    assert attention.fused_applicable(qk, v, H, blocked, 0.0)
    n0 = native.attn_launch_count()
    both = attention.pack_mask(blocked, T, qk.device)
    bits = both[0]
    out = attention.fused_self_attention(qk, v, H, blocked, bits=both)
    assert native.attn_launch_count() == n0 + 2
    want, p_want, lse_want = _fp64_attention(qk.double(), v.double(), H, blocked)
    assert rel(out, want) < 1e-2
    # probabilities / log-sum-exp through the raw entry point
    lib = native.lib()
    o2 = torch.empty_like(out)
    lse = torch.empty(N, H, T, device="cuda")
    p = torch.full((N * H, T, T), float("nan"), device="cuda")
    rc = lib.datr_attn_fused_forward(qk.data_ptr(), 2 * C, qk.data_ptr() + 4 * C, 2 * C, v.data_ptr(), C, bits.data_ptr(),
                                     N, H, T, 1.0 / math.sqrt(32), o2.data_ptr(), lse.data_ptr(), p.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(o2, out)
    assert rel(p.view(N, H, T, T), p_want) < 1e-2
    assert float((lse.double() - lse_want).abs().max()) < 2e-2
    assert torch.allclose(p.sum(-1), torch.ones_like(p[..., 0]), atol=1e-5)
    if blocked is not None:
        assert float(p.view(N, H, T, T)[:, :, blocked].abs().max()) == 0.0


@pytest.mark.parametrize("backward", ["fused", "gemm"])
@pytest.mark.parametrize("N,H,T,masked", [(2, 8, 333, True), (1, 8, 1100, True), (2, 4, 128, False), (1, 2, 37, True),
                                           (1, 8, 900, False)])
def test_fused_attention_gradients_match_fp64(N, H, T, masked, backward, monkeypatch):
    """dQ, dK, dV of the tensor-core backward (two launches, score tiles rebuilt in tensor memory) and of the GEMM-based
    backward on the stored probabilities, against fp64 autograd.  TF32 products: bar 1e-2."""
    from datr_b200 import attention, native
    torch.backends.cuda.matmul.allow_tf32 = False
    monkeypatch.setattr(attention, "_BACKWARD", backward)
    g = torch.Generator(device="cpu").manual_seed(9 + T)
    C = 32 * H
    qk = torch.randn(N, T, 2 * C, generator=g).cuda().requires_grad_(True)
    v = torch.randn(N, T, C, generator=g).cuda().requires_grad_(True)
    go = torch.randn(N, T, C, generator=g).cuda()
    blocked = dn_mask(T, 40 if T > 100 else 20, 10, g).cuda() if masked else None
    n0 = native.attn_launch_count()
    attention.fused_self_attention(qk, v, H, blocked).backward(go)
    assert native.attn_launch_count() == n0 + (4 if backward == "fused" else 3)
    qd, vd = qk.detach().double().requires_grad_(True), v.detach().double().requires_grad_(True)
    _fp64_attention(qd, vd, H, blocked)[0].backward(go.double())
    assert rel(qk.grad[..., :C], qd.grad[..., :C]) < 1e-2, "dq"
    assert rel(qk.grad[..., C:], qd.grad[..., C:]) < 1e-2, "dk"
    assert rel(v.grad, vd.grad) < 1e-2, "dv"
