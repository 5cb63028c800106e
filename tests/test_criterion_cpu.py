"""CPU: host logic of the criterion pipeline -- the batched Hungarian matching and the batched loss
families of datr_b200/models/dino -- against the per-set path that mirrors the reference one to one
(models/dino/dino.py:723-933, matcher.py:47-95)."""
import numpy as np
import pytest
import torch

import model_cases as mcase
from test_model_cpu import cpu_op, small  # noqa: F401  (fixtures)


def _train_outputs(model):
    model.train()
    model.global_proto = None
    torch.manual_seed(7)
    return model(mcase.images(), mcase.targets())


def test_batched_loss_families_equal_the_per_set_losses(small):  # noqa: F811
    model, crit, _ = small
    crit.train()
    tg = mcase.targets()
    res = {}
    for batched in (True, False):
        crit.batched = batched
        model.zero_grad()
        out = _train_outputs(model)
        losses = crit(out, tg)
        total = mcase.total_loss(losses, crit.weight_dict)
        total.backward()
        res[batched] = ({k: float(v) for k, v in losses.items()}, float(total),
                        {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    crit.batched = True
    (la, ta, ga), (lb, tb, gb) = res[True], res[False]
    assert sorted(la) == sorted(lb)
    for k in la:
        assert abs(la[k] - lb[k]) <= 2e-6 * max(1.0, abs(lb[k])), k
    assert abs(ta - tb) <= 1e-6 * abs(tb)
    assert ga.keys() == gb.keys()
    for k in ga:
        scale = float(gb[k].abs().max().clamp_min(1e-8))
        assert float((ga[k] - gb[k]).abs().max()) / scale < 1e-4, k


def test_batched_match_is_bit_identical_to_matching_each_set(small):  # noqa: F811
    from datr_b200.models.dino import matcher as mt
    model, crit, _ = small
    with torch.no_grad():
        out = _train_outputs(model)
    tg = mcase.targets()
    sets = mt.matching_sets(out)
    assert len(sets) == 2 + 1                                        # final + 1 auxiliary layer + intermediate
    assert mt.batchable(crit.matcher, sets)
    handle = mt.BatchedMatch(crit.matcher, sets, tg)
    assert handle.matches(sets, tg) and not handle.matches(sets, mcase.targets())     # same list object required
    pre, nb = handle.result()
    assert nb == float(sum(len(t["labels"]) for t in tg))
    assert not handle.matches(sets, tg)                              # a handle is used once
    for got, o in zip(pre, sets):
        want = crit.matcher(o, tg)
        for (a, b), (c, d) in zip(got, want):
            assert a.dtype == torch.int64 and torch.equal(a, c) and torch.equal(b, d)
    assert [[(a.tolist(), b.tolist()) for a, b in grp] for grp in mt.match_many(crit.matcher, sets, tg)] == \
           [[(a.tolist(), b.tolist()) for a, b in grp] for grp in pre]


def test_degenerate_boxes_are_reported_by_the_deferred_check(small):  # noqa: F811
    """util/box_ops.py:48-49 of the reference asserts inside generalized_box_iou (two device syncs); the batched
    matcher evaluates the same condition with the cost matrix and raises when the result is collected."""
    from datr_b200.models.dino import matcher as mt
    model, crit, _ = small
    with torch.no_grad():
        out = _train_outputs(model)
    bad = {k: v for k, v in out.items()}
    bad["pred_boxes"] = out["pred_boxes"].clone()
    bad["pred_boxes"][0, 0, 2:] = -0.5                               # negative width / height -> x1 < x0
    sets = mt.matching_sets(bad)
    handle = mt.BatchedMatch(crit.matcher, sets, mcase.targets())
    with pytest.raises(AssertionError, match="degenerate boxes"):
        handle.result()
    with pytest.raises(AssertionError):
        crit.matcher(sets[0], mcase.targets())                       # the per-set path asserts like the reference


def test_prefetch_is_a_no_op_on_cpu_and_take_prefetched_clears_stale_handles(small):  # noqa: F811
    from datr_b200.models.dino import matcher as mt
    model, crit, _ = small
    with torch.no_grad():
        out = _train_outputs(model)
    tg = mcase.targets()
    mt.prefetch(crit.matcher, out, tg)
    assert not hasattr(out["pred_logits"], "_datr_match")            # CPU tensors: nothing started
    sets = mt.matching_sets(out)
    out["pred_logits"]._datr_match = mt.BatchedMatch(crit.matcher, sets, tg)
    assert mt.take_prefetched(out, sets, mcase.targets()) is None    # other targets: rejected ...
    assert not hasattr(out["pred_logits"], "_datr_match")            # ... and removed
    assert getattr(model, "_prefetch_matcher", None) is crit.matcher
    assert "_prefetch_matcher" not in dict(model.named_modules())
