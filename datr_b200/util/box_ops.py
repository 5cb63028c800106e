"""Box geometry used by the matcher and the box losses.

Mirrors the functions of the reference's util/box_ops.py that the hot path calls
(box_cxcywh_to_xyxy :9-13, box_xyxy_to_cxcywh :16-20, box_iou :24-39, generalized_box_iou :41-63),
including the +1e-6 stabilisers so the Hungarian costs and GIoU losses agree to fp32 rounding.
"""
import torch


def box_cxcywh_to_xyxy(x):
    c, s = x[..., :2], x[..., 2:] * 0.5
    return torch.cat([c - s, c + s], dim=-1)


def box_xyxy_to_cxcywh(x):
    lo, hi = x[..., :2], x[..., 2:]
    return torch.cat([(lo + hi) / 2, hi - lo], dim=-1)


def _area(b):
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def box_iou(boxes1, boxes2):
    """Pairwise IoU [N,M] and union [N,M] of xyxy boxes."""
    inter_wh = (torch.min(boxes1[:, None, 2:], boxes2[None, :, 2:])
                - torch.max(boxes1[:, None, :2], boxes2[None, :, :2])).clamp(min=0)
    inter = inter_wh[..., 0] * inter_wh[..., 1]
    union = _area(boxes1)[:, None] + _area(boxes2)[None, :] - inter
    return inter / (union + 1e-6), union


def boxes_well_formed(boxes1, boxes2):
    """Device-side bool scalar of the reference's degenerate-box asserts (util/box_ops.py:48-49), for callers that
    check it later on the host instead of synchronising here."""
    return (boxes1[:, 2:] >= boxes1[:, :2]).all() & (boxes2[:, 2:] >= boxes2[:, :2]).all()


def generalized_box_iou(boxes1, boxes2, check=True):
    """Pairwise GIoU [N,M] of xyxy boxes (degenerate boxes are a caller bug, as in the reference).
    check=False skips the two asserts (each is a device synchronisation); pair it with boxes_well_formed()."""
    if check:
        assert (boxes1[:, 2:] >= boxes1[:, :2]).all()
        assert (boxes2[:, 2:] >= boxes2[:, :2]).all()
    iou, union = box_iou(boxes1, boxes2)
    hull_wh = (torch.max(boxes1[:, None, 2:], boxes2[None, :, 2:])
               - torch.min(boxes1[:, None, :2], boxes2[None, :, :2])).clamp(min=0)
    hull = hull_wh[..., 0] * hull_wh[..., 1]
    return iou - (hull - union) / (hull + 1e-6)


def paired_giou(a, b):
    """GIoU of row i of `a` with row i of `b` (the diagonal of generalized_box_iou, without forming
    the N x N matrix the reference builds at dino.py:563-565)."""
    inter_wh = (torch.min(a[:, 2:], b[:, 2:]) - torch.max(a[:, :2], b[:, :2])).clamp(min=0)
    inter = inter_wh[:, 0] * inter_wh[:, 1]
    union = _area(a) + _area(b) - inter
    iou = inter / (union + 1e-6)
    hull_wh = (torch.max(a[:, 2:], b[:, 2:]) - torch.min(a[:, :2], b[:, :2])).clamp(min=0)
    hull = hull_wh[:, 0] * hull_wh[:, 1]
    return iou - (hull - union) / (hull + 1e-6)
