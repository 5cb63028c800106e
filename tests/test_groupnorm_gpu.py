"""GPU parity of the NHWC GroupNorm kernels (include/datr_groupnorm.h) against torch's group_norm in fp64 -- the
nn.GroupNorm(32, 256) of the reference's input projections (models/dino/dino.py:111-126).  Bar: fp32 1e-3 relative per
tensor (north_star); measured ~1e-6."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("N,C,G,H,W", [(4, 256, 32, 100, 167), (2, 256, 32, 13, 21), (1, 256, 32, 1, 1), (3, 128, 8, 7, 9),
                                       (2, 64, 16, 5, 33)])
def test_groupnorm_nhwc_matches_fp64(N, C, G, H, W):
    from datr_b200 import groupnorm as gn, native
    torch.manual_seed(N * 1000 + H)
    m = torch.nn.GroupNorm(G, C).cuda()
    with torch.no_grad():
        m.weight.normal_(1.0, 0.3)
        m.bias.normal_(0.0, 0.3)
    x = (torch.randn(N, C, H, W, device="cuda") * 2 + 3).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gy = torch.randn(N, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    assert gn.applicable(m, x)
    n0 = native.all_launch_count()
    y = gn.group_norm_nhwc(m, x)
    y.backward(gy)
    assert native.all_launch_count() == n0 + 2
    assert y.is_contiguous(memory_format=torch.channels_last) and x.grad.is_contiguous(memory_format=torch.channels_last)
    xd = x.detach().double().requires_grad_(True)
    wd, bd = m.weight.detach().double().requires_grad_(True), m.bias.detach().double().requires_grad_(True)
    yd = F.group_norm(xd, G, wd, bd, m.eps)
    yd.backward(gy.double())
    assert rel(y.detach(), yd.detach()) < 1e-5
    assert rel(x.grad, xd.grad) < 1e-4
    assert rel(m.weight.grad, wd.grad) < 1e-4 and rel(m.bias.grad, bd.grad) < 1e-4


def test_input_projection_takes_the_nhwc_kernel():
    """DINO._group_norm: same result as nn.GroupNorm on the NCHW view, channels_last out, no layout copies."""
    from datr_b200.models.dino.dino import DINO
    m = torch.nn.GroupNorm(32, 256).cuda()
    x = torch.randn(2, 256, 25, 42, device="cuda").contiguous(memory_format=torch.channels_last)
    got = DINO._group_norm(m, x)
    want = m(x)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert rel(got, want) < 1e-5
