"""CPU: the 5-scale model (BASELINE.json configs[3]: C2..C5 + one extra level, 5 MSDeformAttn levels) against goldens
produced by the reference's own model code (tests/golden/make_5scale_golden.py): state-dict layout, eval-mode
outputs, training losses.  The MSDeformAttn op is served by the oracle's grid_sample port (host-logic test)."""
import os

import numpy as np
import pytest
import torch

import model_cases as mcase
from test_model_cpu import cpu_op, finite_rel, rel  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "fivescale_golden.npz"))


@pytest.fixture(scope="module")
def five(cpu_op):  # noqa: F811
    from datr_b200.models.dino.dino import build_dino
    torch.manual_seed(0)
    model, crit, post = build_dino(mcase.small_args(**mcase.FIVE_SCALE))
    model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
    return model, crit, post


def test_5scale_state_dict_and_eval_outputs_match_reference(G, five):
    model, _, _ = five
    assert [f"{k}|{','.join(map(str, v.shape))}" for k, v in model.state_dict().items()] == list(G["keys"])
    assert model.transformer.encoder.layers[0].self_attn.n_levels == 5
    model.eval()
    with torch.no_grad():
        out = model(mcase.images())
    flat = mcase.flatten(out)
    keys = [k[5:] for k in G.files if k.startswith("eval.")]
    assert sorted(flat) == sorted(keys)
    for k in keys:
        assert rel(flat[k].numpy(), G["eval." + k]) < 5e-5, k


def test_5scale_training_losses_match_reference(G, five):
    model, crit, _ = five
    model.train(); crit.train()
    model.global_proto = None
    torch.manual_seed(7)
    out = model(mcase.images(), mcase.targets())
    losses = crit(out, mcase.targets())
    for k in ("pred_logits", "pred_boxes"):
        assert finite_rel(out[k].detach().numpy(), G["train." + k]) < 5e-5, k
    keys = [k[11:] for k in G.files if k.startswith("train.loss.")]
    assert sorted(losses) == sorted(keys)
    for k in keys:
        want = float(G["train.loss." + k])
        assert abs(float(losses[k].detach()) - want) < 1e-4 * max(1.0, abs(want)), k
    total = mcase.total_loss(losses, crit.weight_dict)
    assert abs(float(total.detach()) - float(G["train.total"])) < 1e-4 * abs(float(G["train.total"]))
