"""CPU: host-side dispatch of the op wrappers (datr_b200/linear.py, rowmask.py, layernorm.py, attention.py): on CPU
tensors / in the default 'fp32' mode they must reduce to exactly the torch expressions of the reference's modules
(no native call, no approximation), so the model-level CPU parity tests exercise the reference arithmetic."""
import pytest
import torch
import torch.nn.functional as F

from datr_b200 import attention, linear as dl
from datr_b200.layernorm import layer_norm
from datr_b200.rowmask import zero_masked_rows


def test_linear_modes_equal_the_torch_expressions():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 32, generator=g); w = torch.randn(16, 32, generator=g); b = torch.randn(16, generator=g)
    r = torch.randn(3, 5, 16, generator=g)
    assert dl.get_mode() == "fp32"
    assert torch.equal(dl.linear(x, w, b), F.linear(x, w, b))
    assert torch.equal(dl.linear(x, w, b, relu=True), F.relu(F.linear(x, w, b)))
    assert torch.equal(dl.linear(x, w, b, residual=r), F.linear(x, w, b) + r)
    assert torch.equal(dl.linear(x, w, b, relu=1, residual=r), F.relu(F.linear(x, w, b)) + r)
    assert torch.equal(dl.linear(x, w, b, relu=2, residual=r), F.relu(F.linear(x, w, b) + r))
    w2 = torch.randn(32, 16, generator=g); b2 = torch.randn(32, generator=g)
    assert torch.equal(dl.ffn(x, w, b, w2, b2), F.linear(F.relu(F.linear(x, w, b)), w2, b2) + x)
    with pytest.raises(ValueError):
        dl.set_mode("bf16")


def test_linear_zero_rows_equals_masked_fill_forward_and_backward():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 7, 32, generator=g); w = torch.randn(16, 32, generator=g); b = torch.randn(16, generator=g)
    mask = torch.rand(2, 7, generator=g) < 0.4
    gy = torch.randn(2, 7, 16, generator=g)
    res = []
    for ours in (True, False):
        xa, wa, ba = (t.clone().requires_grad_(True) for t in (x, w, b))
        y = dl.linear(xa, wa, ba, zero_rows=mask) if ours else F.linear(xa, wa, ba).masked_fill(mask[..., None], 0.0)
        y.backward(gy)
        res.append((y.detach(), xa.grad, wa.grad, ba.grad))
    for a, c in zip(*res):
        assert torch.equal(a, c)
    with pytest.raises(ValueError):
        dl.linear(x, w, b, relu=True, zero_rows=mask)


def test_zero_masked_rows_falls_back_to_masked_fill_on_cpu():
    g = torch.Generator().manual_seed(2)
    v = torch.randn(2, 9, 4, 8, generator=g); mask = torch.rand(2, 9, generator=g) < 0.5
    assert torch.equal(zero_masked_rows(v, mask), v.masked_fill(mask[..., None, None], 0.0))
    assert torch.equal(zero_masked_rows(v.view(2, 9, 32), mask), v.view(2, 9, 32).masked_fill(mask[..., None], 0.0))


def test_layer_norm_and_attention_dispatch_on_cpu():
    norm = torch.nn.LayerNorm(256)
    x = torch.randn(4, 256)
    assert torch.equal(layer_norm(norm, x), norm(x))
    q = torch.randn(1, 2, 5, 8)
    assert not attention.applicable(q, None, 0.0)                       # CPU tensors keep the SDPA path
    assert not attention.applicable(q, torch.zeros(5, 5), 0.0)
