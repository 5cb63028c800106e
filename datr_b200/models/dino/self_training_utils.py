"""Pseudo-label plumbing of the mutual-learning step -- host-side mirror of the reference's
models/dino/self_training_utils.py (engine.py:18-20 imports these seven names; the step is engine.py:196-260).

Same function names (the reference's spelling), arguments, return structures and in-place behaviour.  Differences that
do not change results: per-class thresholds and image sizes stay on the device (the reference takes a
`.cpu().numpy()` round trip per image, :36 and :70), and the debug visualiser does not halt the process
(the reference sleeps for 5e6 s after writing its images, :186)."""
import os

import numpy as np
import torch
from torchvision.ops.boxes import batched_nms

from datr_b200.util import box_ops


_CONSTANTS = {}


def _device_constant(array, dtype, device):
    """A small host array as a device tensor, uploaded once per distinct content: a pageable host->device copy is ordered
    behind everything already enqueued on the stream and blocks the host until then (the per-class thresholds and the
    image-index lists of the self-training step were re-uploaded per image and step: two such stalls per step)."""
    a = np.asarray(array)                      # (0-dim arrays stay 0-dim: a scalar threshold)
    key = (a.tobytes(), str(a.dtype), a.shape, dtype, str(device))
    t = _CONSTANTS.get(key)
    if t is None:
        if len(_CONSTANTS) > 256:
            _CONSTANTS.clear()
        t = _CONSTANTS[key] = torch.as_tensor(a, dtype=dtype, device=device)
    return t


def get_unlabel_img(nestedtensor):
    """Target-domain (second) half of a NestedTensor batch: images [B/2, 3, H, W]  (:15-20)."""
    images, _ = nestedtensor.decompose()
    return images[images.shape[0] // 2:]


def get_pseudo_label_via_threshold(results, threshold=0.8):
    """results: per image {'scores','labels','boxes'} (PostProcess output); threshold: per-class array (or scalar).
    Keeps predictions with score >= threshold[label]; returns (indices of images that keep any, labels / boxes / scores
    dicts keyed by image index)  (:23-50).  The comparison runs in float64 like the reference's numpy thresholds."""
    idx_list, labels_d, boxes_d, scores_d = [], {}, {}, {}
    thr = np.asarray(threshold, dtype=np.float64)
    for n, result in enumerate(results):
        scores, labels = result["scores"], result["labels"]
        table = _device_constant(thr, torch.float64, scores.device)
        keep = scores >= (table[labels] if table.dim() else table)
        kept = labels[keep]
        if len(kept) > 0:
            idx_list.append(n)
            labels_d[n], boxes_d[n], scores_d[n] = kept, result["boxes"][keep], scores[keep]
    return idx_list, labels_d, boxes_d, scores_d


def deal_pesudo_label(unlabel_target_list, idx_list, pesudo_labels_dict, pesudo_boxes_dict, scores_dcit):
    """Pseudo labels in the target-dict format of the criterion, keyed by image index  (:52-67)."""
    out = {}
    for i in idx_list:
        t = unlabel_target_list[i]
        out[i] = {"labels": pesudo_labels_dict[i], "boxes": pesudo_boxes_dict[i], "scores": scores_dcit[i],
                  "image_id": t["image_id"], "area": t["area"], "iscrowd": t["iscrowd"], "orig_size": t["orig_size"],
                  "size": t["size"]}
    return out


def rescale_pseudo_targets(unlabel_samples_img, unlabel_pseudo_targets, nms_th=0.7):
    """cxcywh boxes normalised to the padded batch image -> pixels -> class-wise NMS (at most 100 kept) -> cxcywh
    normalised by each image's own size; modifies and returns the dict  (:69-90)."""
    _, _, h, w = unlabel_samples_img.shape
    for k, t in unlabel_pseudo_targets.items():
        size = t["size"]                                    # (h_real, w_real), stays on its device
        boxes = box_ops.box_cxcywh_to_xyxy(t["boxes"])
        boxes[:, [0, 2]] = boxes[:, [0, 2]] * w
        boxes[:, [1, 3]] = boxes[:, [1, 3]] * h
        keep = batched_nms(boxes, t["scores"], t["labels"], nms_th)[:100]
        boxes, t["scores"], t["labels"] = boxes[keep], t["scores"][keep], t["labels"][keep]
        boxes = box_ops.box_xyxy_to_cxcywh(boxes)
        boxes[:, [0, 2]] = boxes[:, [0, 2]] / size[1]
        boxes[:, [1, 3]] = boxes[:, [1, 3]] / size[0]
        t["boxes"] = boxes
    return unlabel_pseudo_targets


def spilt_output(output_dict):
    """(source outputs, target outputs): keys containing 'target' go to the second dict  (:92-100)."""
    source, pseudo = {}, {}
    for k, v in output_dict.items():
        (pseudo if "target" in k else source)[k] = v
    return source, pseudo


def get_valid_output(target_outputs, target_pseudo_labels_dict, idx):
    """Target outputs restricted to the images `idx` that have pseudo labels, and the pseudo labels as a list
    (:103-146)."""
    first = next((v for k, v in target_outputs.items() if "pred" in k), None)
    if first is not None and first.is_cuda and isinstance(idx, (list, tuple)):
        idx = _device_constant(np.asarray(idx, dtype=np.int64), torch.int64, first.device)     # same advanced indexing, no upload
    pick = lambda d: {"pred_logits": d["pred_logits"][idx, :, :], "pred_boxes": d["pred_boxes"][idx, :, :]}
    valid = {}
    for k, v in target_outputs.items():
        if "pred" in k:
            valid[k] = v[idx, :, :]
        elif "aux_outputs_target" in k:
            valid[k] = [pick(d) for d in v]
        elif "interm_outputs_target" in k or "interm_outputs_for_matching_pre_target" in k:
            valid[k] = pick(v)
    return valid, list(target_pseudo_labels_dict.values())


def show_pesudo_label_with_gt(unlabel_img_array, unlabel_pseudo_targets, unlabel_targets, idx_list,
                              unlabel_samples_img_strong_aug_array, save_dir="./show_pseudo"):
    """Debug aid (commented out at engine.py:218): writes, per image with pseudo labels, the weakly augmented image
    with the pseudo boxes and with the ground-truth boxes (cv2)."""
    import cv2
    os.makedirs(save_dir, exist_ok=True)
    mean, std = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1), torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)

    def to_bgr(t):
        img = ((t.detach().cpu() * std + mean) * 255.0).clamp(0, 255).permute(1, 2, 0).numpy().astype(np.uint8)
        return np.ascontiguousarray(img)

    def draw(img, boxes, h, w):
        for cx, cy, bw, bh in boxes:
            cx, cy, bw, bh = int(cx * w), int(cy * h), int(bw * w), int(bh * h)
            cv2.rectangle(img, (cx - bw // 2, cy - bh // 2), (cx + bw // 2, cy + bh // 2), (0, 0, 255), 2)
        return img

    for idx in idx_list:
        h, w = (int(v) for v in unlabel_pseudo_targets[idx]["size"].cpu())
        base = to_bgr(unlabel_img_array[idx])
        cv2.imwrite(os.path.join(save_dir, f"pseudo_{idx}.jpg"), draw(base.copy(), unlabel_pseudo_targets[idx]["boxes"].cpu().tolist(), h, w))
        cv2.imwrite(os.path.join(save_dir, f"label_{idx}.jpg"), draw(base.copy(), unlabel_targets[idx]["boxes"].cpu().tolist(), h, w))
        if unlabel_samples_img_strong_aug_array is not None:
            strong = to_bgr(unlabel_samples_img_strong_aug_array[idx])
            cv2.imwrite(os.path.join(save_dir, f"strong_{idx}.jpg"), draw(strong, unlabel_pseudo_targets[idx]["boxes"].cpu().tolist(), h, w))
