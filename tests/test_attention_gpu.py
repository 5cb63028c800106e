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
