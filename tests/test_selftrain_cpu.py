"""CPU: the target-domain half of the mutual-learning step (BASELINE.json configs[4]; reference engine.py:196-260)
against goldens produced by the reference's own model and SetCriterion (tests/golden/make_selftrain_golden.py):
student outputs with self_training_flag=True -> `*_target` dict -> criterion(..., target_domain_flag=True), and the
teacher-side PostProcess(..., not_to_xyxy=True)."""
import os

import numpy as np
import pytest
import torch

import model_cases as mcase
from test_model_cpu import cpu_op, small  # noqa: F401  (fixtures)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "selftrain_golden.npz"))


def _target_domain_step(model, crit, dtype):
    cast = lambda ts: [{k: (v.to(dtype) if v.is_floating_point() else v) for k, v in t.items()} for t in ts]
    model.train(); crit.train()
    model.global_proto = None
    torch.manual_seed(7)
    out = model([i.to(dtype) for i in mcase.images()], cast(mcase.targets()), self_training_flag=True)
    target_out = mcase.split_target_outputs(out)
    keys = sorted(target_out)        # (the criterion renames pred_*_target in place, like the reference: dino.py:727-730)
    losses = crit(target_out, cast(mcase.pseudo_targets()), target_domain_flag=True)
    total = mcase.total_loss(losses, crit.weight_dict)
    model.zero_grad()
    total.backward()
    return keys, losses, total


def _grad_errors(model, G):
    errs = {}
    for k, p in model.named_parameters():
        if p.grad is not None:
            sub = G["grad_sub." + k]
            errs[k] = float(np.abs(p.grad.reshape(-1)[::101].double().numpy() - sub).max() / max(np.abs(sub).max(), 1e-6))
    return errs


def test_target_domain_losses_and_gradients_match_reference(G, small):  # noqa: F811
    model, crit, _ = small
    out_keys, losses, total = _target_domain_step(model, crit, torch.float32)
    assert out_keys == ["aux_outputs_target", "interm_outputs_for_matching_pre_target", "interm_outputs_target",
                                  "pred_boxes_target", "pred_logits_target"]
    keys = [k[5:] for k in G.files if k.startswith("loss.")]
    assert sorted(losses) == sorted(keys)
    for k in keys:
        assert abs(float(losses[k]) - float(G["loss." + k])) < 1e-4 * max(1.0, abs(float(G["loss." + k]))), k
    assert abs(float(total) - float(G["total"])) < 1e-4 * abs(float(G["total"]))
    errs = _grad_errors(model, G)
    assert set(errs) == {k[9:] for k in G.files if k.startswith("grad_sub.")}
    # The target-domain loss reaches the backbone only through GroupNorm layers that shrink the gradient by 1e5, so
    # a single ReLU whose fp32 pre-activation differs in sign from the reference's (|x| ~ 1e-3 after 40 layers of
    # fp32 rounding) moves a max-norm error to the per-cent level in layer3; the fp64 test below pins those tensors.
    for k, e in errs.items():
        assert e < (5e-2 if k.startswith("backbone") else 2e-3), (k, e)


def test_target_domain_gradients_in_fp64_match_reference(G, cpu_op):  # noqa: F811
    """The same step with the model in fp64 (no rounding-induced ReLU flips): every gradient, backbone included,
    agrees with the reference's fp32 golden to 2e-3 (measured <= 1e-4)."""
    from datr_b200.models.dino.dino import build_dino
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(0)
        model, crit, _ = build_dino(mcase.small_args())
        model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
        model.double(); crit.double()
        _, losses, total = _target_domain_step(model, crit, torch.float64)
    finally:
        torch.set_default_dtype(torch.float32)
    assert abs(float(total) - float(G["total"])) < 1e-5 * abs(float(G["total"]))
    for k, e in _grad_errors(model, G).items():
        assert e < 2e-3, (k, e)


def test_teacher_postprocess_in_normalised_cxcywh_matches_reference(G, small):  # noqa: F811
    model, _, post = small
    model.eval()
    with torch.no_grad():
        pred = model(mcase.images()[2:])
        res = post["bbox"](pred, torch.ones(2, 2), not_to_xyxy=True)
    for i, r in enumerate(res):
        assert np.array_equal(r["labels"].numpy(), G[f"teacher[{i}].labels"])          # top-k index work: bit-exact
        assert np.abs(r["scores"].numpy() - G[f"teacher[{i}].scores"]).max() < 1e-5
        assert np.abs(r["boxes"].numpy() - G[f"teacher[{i}].boxes"]).max() < 1e-5
