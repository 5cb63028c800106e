"""GPU parity of the LayerNorm(256) forward and one-pass backward kernels (include/datr_layernorm.h) against torch's
own LayerNorm in fp64 on the same inputs.  Bar: fp32 1e-3 relative per tensor (measured ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("shape", [(2, 22223, 256), (2, 1100, 256), (1, 7, 256), (3, 1, 256), (44446, 256)])
def test_backward_matches_torch_fp64(shape):
    from datr_b200 import native
    from datr_b200.layernorm import layer_norm
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 3 + 1).cuda()
    gy = torch.randn(shape, generator=g).cuda()
    norm = torch.nn.LayerNorm(256).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(256, generator=g)); norm.bias.copy_(torch.randn(256, generator=g))
    n0 = native.layernorm_launch_count()
    xa = x.clone().requires_grad_(True)
    y = layer_norm(norm, xa)
    y.backward(gy)
    assert native.layernorm_launch_count() == n0 + 2, "CUDA LayerNorm forward + backward did not launch"
    got = (y.detach(), xa.grad, norm.weight.grad.clone(), norm.bias.grad.clone())
    ref = torch.nn.LayerNorm(256).cuda().double()
    with torch.no_grad():
        ref.weight.copy_(norm.weight.double()); ref.bias.copy_(norm.bias.double())
    xb = x.double().requires_grad_(True)
    yb = ref(xb)
    yb.backward(gy.double())
    want = (yb.detach(), xb.grad, ref.weight.grad, ref.bias.grad)
    for a, b, name in zip(got, want, ("y", "dx", "dgamma", "dbeta")):
        assert rel(a, b) < 1e-4, name


def test_column_sum_output():
    """dx_colsum = sum over rows of dx (used as a fused bias gradient), via the C ABI directly."""
    from datr_b200 import native
    lib = native.lib()
    rows = 5000
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(rows, 256, generator=g).cuda(); gy = torch.randn(rows, 256, generator=g).cuda()
    w = torch.randn(256, generator=g).cuda(); b = torch.zeros(256).cuda()
    y, mean, rstd = torch.native_layer_norm(x, (256,), w, b, 1e-5)
    dx = torch.empty_like(x); out = torch.empty(3, 256, device="cuda")
    rc = lib.datr_layernorm256_backward(gy.data_ptr(), x.data_ptr(), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                        dx.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), rows,
                                        torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    want = dx.double().sum(0)
    assert float((out[2].double() - want).abs().max()) < 1e-3 * max(1.0, float(want.abs().max()))


def test_forward_statistics_and_no_grad_path():
    """mean / rstd written by the forward kernel match ATen's, and the kernel also serves torch.no_grad() callers."""
    from datr_b200 import native
    from datr_b200.layernorm import layer_norm
    lib = native.lib()
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(1237, 256, generator=g) * 2 - 0.5).cuda()
    w = torch.randn(256, generator=g).cuda(); b = torch.randn(256, generator=g).cuda()
    y = torch.empty_like(x); stats = torch.empty(2, 1237, device="cuda")
    rc = lib.datr_layernorm256_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), 1e-5, y.data_ptr(), stats[0].data_ptr(),
                                       stats[1].data_ptr(), 1237, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    yr, mr, rr = torch.native_layer_norm(x.double(), (256,), w.double(), b.double(), 1e-5)
    assert rel(y, yr) < 1e-5 and rel(stats[0], mr.view(-1)) < 1e-5 and rel(stats[1], rr.view(-1)) < 1e-5
    norm = torch.nn.LayerNorm(256).cuda()
    n0 = native.layernorm_launch_count()
    with torch.no_grad():
        z = layer_norm(norm, x)
    assert native.layernorm_launch_count() == n0 + 1
    assert rel(z, norm(x).double()) < 1e-5
