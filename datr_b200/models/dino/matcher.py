"""Hungarian assignment between predictions and ground-truth boxes.

Mirrors HungarianMatcher / SimpleMinsumMatcher / build_matcher of the reference's
models/dino/matcher.py (:23-95, :98-169, :172-190): cost = cost_class * focal-style class cost
+ cost_bbox * L1 + cost_giou * (-GIoU), solved per image by scipy's linear_sum_assignment on the host
(the assignment must stay bit-exact, so the solver is the same library call as the reference's).
One D2H copy per call, on the current stream, through pinned memory.
"""
import torch
from scipy.optimize import linear_sum_assignment
from torch import nn

from datr_b200.util.box_ops import box_cxcywh_to_xyxy, generalized_box_iou


def _cost_matrix(outputs, targets, w_class, w_bbox, w_giou, alpha, gamma=2.0):
    bs, nq = outputs["pred_logits"].shape[:2]
    prob = outputs["pred_logits"].flatten(0, 1).sigmoid()
    boxes = outputs["pred_boxes"].flatten(0, 1)
    tgt_ids = torch.cat([t["labels"] for t in targets])
    tgt_box = torch.cat([t["boxes"] for t in targets])
    neg = (1 - alpha) * (prob ** gamma) * (-(1 - prob + 1e-8).log())
    pos = alpha * ((1 - prob) ** gamma) * (-(prob + 1e-8).log())
    c_class = pos[:, tgt_ids] - neg[:, tgt_ids]
    c_bbox = torch.cdist(boxes, tgt_box, p=1)
    c_giou = -generalized_box_iou(box_cxcywh_to_xyxy(boxes), box_cxcywh_to_xyxy(tgt_box))
    C = w_bbox * c_bbox + w_class * c_class + w_giou * c_giou
    return C.view(bs, nq, -1), [len(t["boxes"]) for t in targets]


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal_alpha=0.25):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha = cost_class, cost_bbox, cost_giou, focal_alpha

    @torch.no_grad()
    def forward(self, outputs, targets):
        """-> [(pred_idx int64[K_i], tgt_idx int64[K_i])] per image, K_i = min(num_queries, num_targets_i)."""
        C, sizes = _cost_matrix(outputs, targets, self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha)
        C = C.cpu()
        pairs = [linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))]
        return [(torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)) for i, j in pairs]


def match_many(matcher, outputs_list, targets):
    """Assignments for several prediction sets against the same targets with ONE cost-matrix pass and ONE device->host
    copy: the sets are stacked along the batch axis, which leaves every row of the cost matrix -- and so
    every assignment -- bit-identical to matching them one by one (the reference does 7 matchings with 7 syncs per
    step, dino.py:723-933 / matcher.py:91).  Returns one index list per prediction set."""
    same = all(o["pred_logits"].shape == outputs_list[0]["pred_logits"].shape for o in outputs_list)
    if not isinstance(matcher, HungarianMatcher) or not same or len(outputs_list) == 1:
        return [matcher(o, targets) for o in outputs_list]
    bs = outputs_list[0]["pred_logits"].shape[0]
    stacked = {"pred_logits": torch.cat([o["pred_logits"] for o in outputs_list], 0),
               "pred_boxes": torch.cat([o["pred_boxes"] for o in outputs_list], 0)}
    with torch.no_grad():
        C, sizes = _cost_matrix(stacked, targets, matcher.cost_class, matcher.cost_bbox, matcher.cost_giou, matcher.focal_alpha)
        C = C.cpu()
    pairs = [[linear_sum_assignment(c[g * bs + i]) for i, c in enumerate(C.split(sizes, -1))]
             for g in range(len(outputs_list))]
    # all index vectors travel to the device in ONE pinned, asynchronous copy (the losses index device tensors with
    # them; CPU index tensors would cost a pageable host->device copy per use, ~40 per step)
    import numpy as np
    flat = np.concatenate([np.asarray(v, dtype=np.int64) for grp in pairs for ij in grp for v in ij]) \
        if any(len(ij[0]) for grp in pairs for ij in grp) else np.zeros(0, dtype=np.int64)
    dev = stacked["pred_logits"].device
    host = torch.from_numpy(flat)
    if dev.type == "cuda":
        host = host.pin_memory()
    flat_dev = host.to(dev, non_blocking=True)
    out, off = [], 0
    for grp in pairs:
        cur = []
        for i, j in grp:
            n = len(i)
            cur.append((flat_dev[off:off + n], flat_dev[off + n:off + 2 * n]))
            off += 2 * n
        out.append(cur)
    return out


class SimpleMinsumMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal_alpha=0.25):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha = cost_class, cost_bbox, cost_giou, focal_alpha

    @torch.no_grad()
    def forward(self, outputs, targets):
        C, sizes = _cost_matrix(outputs, targets, self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha)
        out = []
        for i, (c, n) in enumerate(zip(C.split(sizes, -1), sizes)):
            out.append((c[i].min(0)[1].to(torch.int64), torch.arange(n, device=C.device, dtype=torch.int64)))
        return out


def build_matcher(args):
    kinds = {"HungarianMatcher": HungarianMatcher, "SimpleMinsumMatcher": SimpleMinsumMatcher}
    assert args.matcher_type in kinds, f"Unknown args.matcher_type: {args.matcher_type}"
    return kinds[args.matcher_type](cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox,
                                    cost_giou=args.set_cost_giou, focal_alpha=args.focal_alpha)
