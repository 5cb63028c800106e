"""GPU parity of the implicit-GEMM 3x3 convolution (include/datr_conv.h) against torch's convolution in fp64 on the
same NHWC inputs.  Bar: tensor-core (TF32) class, 2e-3 relative per tensor forward; gradients (cuDNN backward on the
same operands, ReLU mask taken from the kernel's own output) 2e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
REL = 2e-3

CASES = [  # N, Cin, Cout, H, W, stride
    (2, 64, 64, 50, 84, 1),
    (1, 128, 128, 25, 42, 1),
    (2, 128, 128, 51, 83, 2),      # odd sizes, stride 2 (first block of a ResNet stage)
    (1, 256, 256, 13, 21, 1),
    (1, 512, 512, 8, 16, 1),
    (3, 64, 128, 9, 7, 1),         # patch larger than the map
    (1, 32, 4, 17, 33, 2),
]


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("N,Cin,Cout,H,W,stride", CASES)
def test_forward_backward(N, Cin, Cout, H, W, stride):
    from datr_b200 import native
    from datr_b200.conv import conv3x3_bias_relu
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + Cin + H)
    x = torch.randn(N, Cin, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda().contiguous(memory_format=torch.channels_last)
    b = torch.randn(Cout, generator=g).cuda()
    xa, wa, ba = (t.clone().requires_grad_(True) for t in (x, w, b))
    n0 = native.conv_launch_count()
    y = conv3x3_bias_relu(xa, wa, ba, stride)
    assert native.conv_launch_count() == n0 + 1
    assert y.is_contiguous(memory_format=torch.channels_last) or y.numel() == y.shape[1]
    xd, wd, bd = (t.double().clone().requires_grad_(True) for t in (x, w, b))
    z = F.conv2d(xd, wd, bd, stride=stride, padding=1)
    assert y.shape == z.shape
    assert rel(y.detach(), z.detach().clamp_min(0)) < REL
    gy = torch.randn(y.shape, generator=g).cuda()
    y.backward(gy)
    (z * (y.detach() > 0)).backward(gy.double())        # same active set as the kernel's output
    torch.backends.cudnn.allow_tf32 = False
    for got, want in ((xa.grad, xd.grad), (wa.grad, wd.grad), (ba.grad, bd.grad)):
        assert rel(got, want) < 5e-3
