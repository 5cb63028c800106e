"""CPU: the oracle of the fused MSDeformAttn entry points (oracle/msda.py::module_prologue_torch / fused_torch) against
the REFERENCE's own MSDeformAttn module run live in the build container (its forward with the pure-PyTorch op) and
against the committed module-level goldens (mod.msda_2d / mod.msda_4d of tests/golden/model_golden.npz)."""
import os

import numpy as np
import pytest
import torch

import ref_loader
from oracle import msda as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(ref_dim, seed):
    rng = np.random.default_rng(seed)
    levels = [(8, 10), (4, 5), (2, 3), (1, 2)]
    S = sum(h * w for h, w in levels)
    q = torch.from_numpy(rng.standard_normal((2, 7, 256)))
    src = torch.from_numpy(rng.standard_normal((2, S, 256)))
    ref = torch.from_numpy(rng.uniform(0.1, 0.9, (2, 7, 4, ref_dim)))
    mask = torch.zeros(2, S, dtype=torch.bool); mask[1, -3:] = True
    shapes = torch.tensor(levels)
    return q, src, ref, mask, shapes, torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("ref_dim", [2, 4])
def test_fused_oracle_equals_the_reference_module(ref_dim):
    ns = ref_loader.load()
    torch.manual_seed(ref_dim)
    m = ns.msda_module.MSDeformAttn(256, 4, 8, 4).double()
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.05); m.attention_weights.weight.normal_(0, 0.05)
    q, src, ref, mask, shapes, lstart = _inputs(ref_dim, 3 + ref_dim)
    with torch.no_grad():
        want = m(q, ref, src, shapes, lstart, mask)
        value = m.value_proj(src).masked_fill(mask[..., None], 0.0).view(2, -1, 8, 32)
        off = m.sampling_offsets(q).view(2, 7, 8, 4, 4, 2)
        lg = m.attention_weights(q).view(2, 7, 8, 16)
        got = m.output_proj(om.fused_torch(value, shapes, off, lg, ref, 4))
    assert float((got - want).abs().max()) < 1e-12 * max(1.0, float(want.abs().max()))


def test_prologue_matches_hand_computed_values():
    """2-d: loc = ref + off / (W, H); 4-d: loc = ref.xy + off / P * ref.wh / 2; weights sum to one per (query, head)."""
    shapes = [(3, 5), (2, 2)]
    off = torch.zeros(1, 1, 1, 2, 2, 2, dtype=torch.float64)
    off[0, 0, 0, 0, 1] = torch.tensor([2.5, -1.5]); off[0, 0, 0, 1, 0] = torch.tensor([1.0, 1.0])
    lg = torch.tensor([[[[0.0, 1.0, 2.0, 3.0]]]], dtype=torch.float64)
    ref2 = torch.tensor([[[[0.5, 0.5], [0.25, 0.75]]]], dtype=torch.float64)
    loc, attn = om.module_prologue_torch(off, lg, ref2, shapes, 2)
    assert torch.allclose(loc[0, 0, 0, 0, 1], torch.tensor([0.5 + 2.5 / 5, 0.5 - 1.5 / 3], dtype=torch.float64))
    assert torch.allclose(loc[0, 0, 0, 1, 0], torch.tensor([0.25 + 1.0 / 2, 0.75 + 1.0 / 2], dtype=torch.float64))
    assert torch.allclose(attn.sum((-1, -2)), torch.ones(1, 1, 1, dtype=torch.float64))
    assert torch.allclose(attn.reshape(-1), torch.softmax(lg.reshape(-1), 0))
    ref4 = torch.tensor([[[[0.5, 0.5, 0.2, 0.4], [0.5, 0.5, 0.2, 0.4]]]], dtype=torch.float64)
    loc4, _ = om.module_prologue_torch(off, lg, ref4, shapes, 2)
    assert torch.allclose(loc4[0, 0, 0, 0, 1], torch.tensor([0.5 + 2.5 / 2 * 0.2 * 0.5, 0.5 - 1.5 / 2 * 0.4 * 0.5], dtype=torch.float64))
    with pytest.raises(ValueError):
        om.module_prologue_torch(off, lg, ref4[..., :3], shapes, 2)
