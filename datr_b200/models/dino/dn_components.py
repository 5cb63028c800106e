"""Contrastive de-noising (CDN) queries.

Mirrors prepare_for_cdn (reference models/dino/dn_components.py:20-136) and dn_post_process (:139-155).
The random draws are made in the reference's order, with the reference's shapes and dtypes
(label flip mask, replacement labels, box sign, box magnitude), so a shared generator state gives the
same noised queries.  Differences: tensors are created on the device of `label_enc` instead of a
hard-coded .cuda(), group counts come from tensor shapes (no host sync), and the attention mask is
built from group ids in one comparison instead of a Python loop of slice writes.
"""
import torch

from datr_b200.util.misc import inverse_sigmoid

# random sources (module-level so a test can feed a CPU-generated stream to a CUDA run)
_rand_like = torch.rand_like
_randint_like = torch.randint_like

# CUDA tensors: draw the label noise without a host synchronisation (same distribution, different consumption of the random
# stream than the reference).  DATR_DN_SYNC_FREE=0 / SYNC_FREE = False reproduces the reference's draws one by one (what the
# parity tools and golden tests use); CPU tensors always take the reference's sequence.
import os as _os
SYNC_FREE = _os.environ.get("DATR_DN_SYNC_FREE", "1") != "0"


def prepare_for_cdn(dn_args, training, num_queries, num_classes, hidden_dim, label_enc):
    """-> (input_query_label [B,pad,C], input_query_bbox [B,pad,4] (logits), attn_mask [pad+nq,pad+nq] bool
    (True = blocked), dn_meta {'pad_size','num_dn_group'}); all None when not training."""
    if not training:
        return None, None, None, None
    targets, dn_number, label_noise_ratio, box_noise_scale = dn_args
    device = label_enc.weight.device
    counts = [int(t["labels"].shape[0]) for t in targets]
    batch_size, most = len(targets), max(counts) if counts else 0

    groups = dn_number * 2
    if most == 0:
        groups = 1
    elif groups >= 100:
        groups = groups // (most * 2)
    elif groups < 1:
        groups = 1
    groups = max(groups, 1)

    labels = torch.cat([t["labels"] for t in targets]).to(device)
    boxes = torch.cat([t["boxes"] for t in targets]).to(device)
    total = labels.shape[0]
    image_of = torch.cat([torch.full((n,), i, dtype=torch.long, device=device) for i, n in enumerate(counts)]) \
        if counts else torch.zeros(0, dtype=torch.long, device=device)

    reps = 2 * groups                                   # positive + negative copy per group
    noisy_labels = labels.repeat(reps)
    image_of_rep = image_of.repeat(reps)
    gt_boxes = boxes.repeat(reps, 1)
    noisy_boxes = gt_boxes.clone()

    if label_noise_ratio > 0:
        if SYNC_FREE and noisy_labels.is_cuda:
            # the same noise law (every label replaced by a uniform class with probability ratio / 2) without the
            # device->host synchronisation of nonzero(): that sync sits at the very start of the forward pass and drains the
            # GPU once per step (17.8 ms of host wait per step in tools/host_profile.py), so the host can never enqueue ahead
            flip = _rand_like(noisy_labels.float()) < label_noise_ratio * 0.5
            noisy_labels = torch.where(flip, _randint_like(noisy_labels, 0, num_classes), noisy_labels)
        else:
            # the reference's own sequence of random draws (dn_components.py:60-63): nonzero + a draw per flipped label
            flip = torch.nonzero(_rand_like(noisy_labels.float()) < label_noise_ratio * 0.5).view(-1)
            noisy_labels.scatter_(0, flip, _randint_like(flip, 0, num_classes))

    pad_size = most * reps
    # rows [g*2T, g*2T+T) of the repeated set are positives of group g, the next T rows negatives
    row = torch.arange(total * reps, device=device)
    is_negative = (row // max(total, 1)) % 2 == 1
    if box_noise_scale > 0:
        corners = torch.cat([gt_boxes[:, :2] - gt_boxes[:, 2:] / 2, gt_boxes[:, :2] + gt_boxes[:, 2:] / 2], 1)
        half = (gt_boxes[:, 2:] / 2).repeat(1, 2)
        sign = _randint_like(gt_boxes, low=0, high=2, dtype=torch.float32) * 2.0 - 1.0
        mag = _rand_like(gt_boxes)
        mag = (mag + is_negative[:, None].to(mag.dtype)) * sign
        corners = (corners + mag * half * box_noise_scale).clamp(min=0.0, max=1.0)
        noisy_boxes = torch.cat([(corners[:, :2] + corners[:, 2:]) / 2, corners[:, 2:] - corners[:, :2]], 1)

    label_embed = label_enc(noisy_labels.long())
    bbox_embed = inverse_sigmoid(noisy_boxes)

    query_label = torch.zeros(batch_size, pad_size, hidden_dim, device=device)
    query_bbox = torch.zeros(batch_size, pad_size, 4, device=device)
    if total:
        within = torch.cat([torch.arange(n, device=device) for n in counts])          # index inside its image
        slot = (within[None, :] + most * torch.arange(reps, device=device)[:, None]).reshape(-1)
        query_label[(image_of_rep, slot)] = label_embed
        query_bbox[(image_of_rep, slot)] = bbox_embed

    size = pad_size + num_queries
    group_of = torch.arange(size, device=device) // max(2 * most, 1)               # dn group id; matching part >= groups
    is_dn = torch.arange(size, device=device) < pad_size
    # matching queries never see dn queries; dn queries see only their own group (and all matching queries)
    attn_mask = is_dn[None, :] & ((~is_dn)[:, None] | (group_of[:, None] != group_of[None, :]))
    dn_meta = {"pad_size": pad_size, "num_dn_group": groups}
    return query_label, query_bbox, attn_mask, dn_meta


def dn_post_process(outputs_class, outputs_coord, dn_meta, aux_loss, _set_aux_loss):
    """Split the de-noising part off the decoder outputs [n_layers,B,pad+nq,*] and park it in dn_meta."""
    if dn_meta and dn_meta["pad_size"] > 0:
        pad = dn_meta["pad_size"]
        known_class, known_coord = outputs_class[:, :, :pad], outputs_coord[:, :, :pad]
        outputs_class, outputs_coord = outputs_class[:, :, pad:], outputs_coord[:, :, pad:]
        out = {"pred_logits": known_class[-1], "pred_boxes": known_coord[-1]}
        if aux_loss:
            out["aux_outputs"] = _set_aux_loss(known_class, known_coord)
        dn_meta["output_known_lbs_bboxes"] = out
    return outputs_class, outputs_coord
