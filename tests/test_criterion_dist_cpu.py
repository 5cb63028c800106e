"""CPU, world_size 2 over gloo: the num_boxes normalisation of SetCriterion under data parallelism -- the reference
all-reduces the number of target boxes and divides by the world size (models/dino/dino.py:767-770) -- through both
code paths of datr_b200: riding along with the batched matching (BatchedMatch) and the per-call all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

COUNTS = {0: (3, 5), 1: (1, 2)}          # target boxes per image on each rank


def _sets_and_targets(rank):
    rng = np.random.default_rng(100 + rank)
    mk = lambda: {"pred_logits": torch.from_numpy(rng.standard_normal((2, 12, 9)).astype(np.float32)),
                  "pred_boxes": torch.from_numpy(rng.uniform(0.2, 0.6, (2, 12, 4)).astype(np.float32))}
    sets = [mk(), mk(), mk()]
    targets = []
    for n in COUNTS[rank]:
        targets.append({"labels": torch.from_numpy(rng.integers(0, 9, n)).long(),
                        "boxes": torch.from_numpy(np.concatenate([rng.uniform(0.3, 0.7, (n, 2)), rng.uniform(0.1, 0.3, (n, 2))], 1).astype(np.float32))})
    return sets, targets


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from datr_b200.models.dino import matcher as mt
        from datr_b200.models.dino.dino import SetCriterion
        m = mt.HungarianMatcher(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0)
        sets, targets = _sets_and_targets(rank)
        pre, nb = mt.BatchedMatch(m, sets, targets).result()
        crit = SetCriterion(9, matcher=m, weight_dict={"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0}, focal_alpha=0.25,
                            losses=["labels", "boxes", "cardinality"])
        outputs = dict(sets[0], aux_outputs=sets[1:2], interm_outputs=sets[2], dn_meta=None)
        losses = crit(outputs, targets)                                   # batched path (num_boxes from BatchedMatch)
        crit.batched = False
        per_set = crit(dict(outputs), targets)                            # per-set path of the reference
        target_side = crit({"pred_logits_target": sets[0]["pred_logits"], "pred_boxes_target": sets[0]["pred_boxes"]},
                           targets, target_domain_flag=True)              # no batched handle: all-reduce + .item()
        out[rank] = (nb, {k: float(v) for k, v in losses.items()}, {k: float(v) for k, v in per_set.items()},
                     [[(a.tolist(), b.tolist()) for a, b in grp] for grp in pre], sorted(target_side))
    finally:
        dist.destroy_process_group()


def test_num_boxes_is_averaged_over_ranks_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    want_nb = (sum(COUNTS[0]) + sum(COUNTS[1])) / 2
    for rank in (0, 1):
        nb, losses, per_set, pre, target_keys = out[rank]
        assert nb == want_nb
        assert sorted(losses) == sorted(per_set)
        for k in losses:
            assert abs(losses[k] - per_set[k]) <= 2e-6 * max(1.0, abs(per_set[k])), k
        # single-process cross-check: the same sets scored with the averaged num_boxes given explicitly
        from datr_b200.models.dino import matcher as mt
        from datr_b200.models.dino.dino import SetCriterion
        m = mt.HungarianMatcher(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0)
        sets, targets = _sets_and_targets(rank)
        crit = SetCriterion(9, matcher=m, weight_dict={}, focal_alpha=0.25, losses=["labels", "boxes", "cardinality"])
        idx = m(sets[0], targets)
        assert [(a.tolist(), b.tolist()) for a, b in idx] == pre[0]
        ref = crit.loss_boxes(sets[0], targets, idx, want_nb)
        assert abs(float(ref["loss_bbox"]) - losses["loss_bbox"]) < 1e-6 * max(1.0, abs(losses["loss_bbox"]))
        assert target_keys == []          # a bare final-layer dict has no auxiliary / intermediate sets to score
