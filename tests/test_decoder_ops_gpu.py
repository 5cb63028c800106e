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
