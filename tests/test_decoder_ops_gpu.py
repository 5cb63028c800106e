"""GPU parity of the small fused kernels of the decoder layer loop (include/datr_decoder_ops.h) against the torch
expressions they replace (reference models/dino/utils.py:gen_sineembed_for_position)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _torch_sine_embed(pos):
    k = pos.size(-1)
    idx = torch.arange(128, dtype=torch.float32, device=pos.device)
    dim_t = 10000 ** (2 * torch.div(idx, 2, rounding_mode="floor") / 128)
    ang = pos.unsqueeze(-1) * (2 * math.pi) / dim_t
    emb = torch.stack((ang[..., 0::2].sin(), ang[..., 1::2].cos()), dim=-1).flatten(-2)
    return torch.cat((emb[..., 1:2, :], emb[..., 0:1, :], emb[..., 2:, :]), dim=-2).flatten(-2)


@pytest.mark.parametrize("shape", [(4, 1100, 4), (2, 900, 2), (1, 7, 4)])
def test_sine_embed_kernel_matches_the_torch_expression(shape):
    from datr_b200 import native
    from datr_b200.models.dino.utils import gen_sineembed_for_position
    g = torch.Generator(device="cpu").manual_seed(shape[1])
    pos = torch.rand(shape, generator=g).cuda()
    n0 = native.all_launch_count()
    got = gen_sineembed_for_position(pos)
    assert native.all_launch_count() == n0 + 1                      # the kernel ran, not the ATen chain
    want = _torch_sine_embed(pos)
    assert got.shape == want.shape
    # same operation order (x * 2 pi, / dim_t, sinf / cosf); allow the last bit of the device sin / cos
    assert float((got - want).abs().max()) <= 2e-6
    # a tensor that requires grad keeps the differentiable torch path
    pos_g = pos.clone().requires_grad_(True)
    out = gen_sineembed_for_position(pos_g)
    assert out.requires_grad and torch.equal(out.detach(), want)


def test_position_embedding_kernel_matches_the_torch_expression():
    """PositionEmbeddingSineHW (reference position_encoding.py:62-107) with ragged padding masks."""
    from datr_b200.models.dino.position_encoding import PositionEmbeddingSineHW
    from datr_b200.util.misc import NestedTensor
    pe = PositionEmbeddingSineHW(128, temperatureH=20, temperatureW=20, normalize=True)
    mask = torch.zeros(3, 25, 42, dtype=torch.bool, device="cuda")
    mask[1, 20:, :] = True
    mask[2, :, 30:] = True
    x = torch.zeros(3, 256, 25, 42, device="cuda")
    got = pe(NestedTensor(x, mask))
    keep = ~mask
    yy = keep.cumsum(1, dtype=torch.float32)
    xx = keep.cumsum(2, dtype=torch.float32)
    yy = yy / (yy[:, -1:, :] + 1e-6) * pe.scale
    xx = xx / (xx[:, :, -1:] + 1e-6) * pe.scale
    want = torch.cat((pe._axis(yy, 20), pe._axis(xx, 20)), dim=3).permute(0, 3, 1, 2)
    assert got.shape == want.shape == (3, 256, 25, 42)
    assert float((got - want).abs().max()) <= 2e-6


@pytest.mark.parametrize("shape", [(2, 64, 400, 667), (1, 64, 7, 9), (3, 8, 5, 4)])
def test_stem_tail_kernel_matches_bn_relu_maxpool(shape):
    """FrozenBatchNorm2d -> ReLU -> MaxPool2d(3, 2, 1) as one kernel on the NHWC stem output (odd sizes, negative scales)."""
    import torch.nn.functional as F
    from datr_b200 import native
    g = torch.Generator(device="cpu").manual_seed(shape[2])
    n, c, h, w = shape
    x = torch.randn(shape, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    scale = torch.randn(c, generator=g).cuda()
    shift = torch.randn(c, generator=g).cuda()
    out = torch.empty((n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1), device="cuda").contiguous(memory_format=torch.channels_last)
    rc = native.lib().datr_bn_relu_maxpool_nhwc(x.data_ptr(), scale.data_ptr(), shift.data_ptr(), n, h, w, c, out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    want = F.max_pool2d(F.relu(torch.addcmul(shift.view(1, -1, 1, 1), x, scale.view(1, -1, 1, 1))), 3, stride=2, padding=1)
    assert out.shape == want.shape
    assert float((out - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))
