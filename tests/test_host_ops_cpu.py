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


def test_backward_fusion_flags_need_the_tensor_core_path():
    """skip / mask / chain flags change what the backward returns to autograd: on the plain torch path they must not be
    silently ignored."""
    x = torch.randn(4, 32); w = torch.randn(16, 32)
    for kw in (dict(skip_in=dl.GradCarrier()), dict(skip_out=dl.GradCarrier(), residual=torch.randn(4, 16)),
               dict(mask_input_grad=True), dict(grad_premasked=True), dict(chain=dl.GradChain(2))):
        with pytest.raises(ValueError):
            dl.linear(x, w, None, **kw)


def test_grad_sinks_take_parameters_only():
    """linear.GradSinks.lookup: a weight-gradient kernel may reduce into a registered buffer only if the Linear's weight (and
    bias) ARE the registered parameters -- slices / products of parameters are post-processed by autograd and keep the
    ordinary path; shape, density and alignment of the buffer are checked."""
    lin = torch.nn.Linear(32, 16)
    flat = torch.zeros(16 * 32 + 16 + 4)
    vw, vb = flat[:512].view(16, 32), flat[512:528]
    sinks = dl.GradSinks({id(lin.weight): vw, id(lin.bias): vb})
    sw, sb = sinks.lookup(lin.weight, lin.bias, True)
    assert sw.data_ptr() == vw.data_ptr() and sb.data_ptr() == vb.data_ptr()
    assert sinks.lookup(lin.weight, lin.bias, False)[1] is None            # no bias gradient wanted
    assert sinks.lookup(lin.weight[:8], lin.bias[:8], True) is None        # a slice is not the parameter
    assert sinks.lookup(lin.weight * 2.0, lin.bias, True) is None          # nor is a BN-folded product
    assert sinks.lookup(lin.weight, None, True) is None
    other = torch.nn.Linear(32, 16)
    assert sinks.lookup(other.weight, other.bias, True) is None            # not registered
    assert dl.GradSinks({id(lin.weight): flat[1:513].view(16, 32)}).lookup(lin.weight, None, False) is None   # misaligned
    assert dl.GradSinks({id(lin.weight): flat[:512].view(32, 16)}).lookup(lin.weight, None, False) is None    # other shape


def test_step_graphs_drop_training_captures_when_the_gradient_buffer_moves():
    """graphs.StepGraphs.set_grad_sinks: captures made against an earlier FlatGradients buffer would add into memory nobody
    reads; registering a new buffer for the same parameters forgets them (they are captured again on next use)."""
    from datr_b200 import graphs
    from datr_b200.parallel import FlatGradients
    net = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    sg = graphs.StepGraphs()
    a = FlatGradients(net)
    sg.set_grad_sinks(a)
    assert len(sg.grad_sinks) == 4 and all(sg.grad_sinks[id(p)].data_ptr() == p.grad.data_ptr() for p in net.parameters())
    sg.cache[("enc", 0, True, None, ())] = ("training graph", 3)
    sg.cache[("enc", 0, False, None, ())] = ("inference graph", 3)
    sg.per_segment[("enc", 0)] = 2
    sg.set_grad_sinks(a)                                   # same buffer: nothing to forget
    assert len(sg.cache) == 2
    b = FlatGradients(net)                                 # new buffer for the same parameters
    sg.set_grad_sinks(b)
    assert list(sg.cache) == [("enc", 0, False, None, ())] and sg.per_segment[("enc", 0)] == 1
    assert all(sg.grad_sinks[id(p)].data_ptr() == p.grad.data_ptr() for p in net.parameters())
    sg.set_grad_sinks(FlatGradients(net, gather=True))     # gather mode has no persistent views: ignored
    assert all(sg.grad_sinks[id(p)].data_ptr() == v.data_ptr() for p, v in zip(b.params, b.views))


def test_resnet_stage_runner_is_the_sequential_on_the_plain_path():
    """backbone.run_stage only re-routes gradients on the tensor-core path; elsewhere it is stage(x) -- same output, same
    gradients."""
    from datr_b200.models.dino import backbone as bb
    torch.manual_seed(0)
    down = torch.nn.Sequential(torch.nn.Conv2d(16, 32, 1, stride=2, bias=False), bb.FrozenBatchNorm2d(32))
    stage = torch.nn.Sequential(bb.Bottleneck(16, 8, stride=2, downsample=down), bb.Bottleneck(32, 8), bb.Bottleneck(32, 8))
    x = torch.randn(2, 16, 12, 10)
    outs = []
    for fn in (lambda t: bb.run_stage(stage, t), stage):
        xa = x.clone().requires_grad_(True)
        stage.zero_grad()
        y = fn(xa)
        y.square().sum().backward()
        outs.append([y.detach(), xa.grad] + [p.grad.clone() for p in stage.parameters()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
