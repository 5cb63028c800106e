"""Seeded small DINO configurations, weights, images and targets for the model-level parity tests.

The same numpy-seeded state dict is loaded into the reference model (in the build container, by
tests/golden/make_model_golden.py) and into datr_b200's model (everywhere), so no weights are stored."""
from __future__ import annotations

import numpy as np
import torch

from datr_b200.config import dino_args

SMALL = dict(enc_layers=2, dec_layers=2, dim_feedforward=64, num_queries=30, num_classes=9, dn_labelbook_size=9,
             dn_number=100, num_select=100)
FIVE_SCALE = dict(return_interm_indices=[0, 1, 2, 3], num_feature_levels=5)       # BASELINE.json configs[3]
IMAGE_SIZES = [(128, 160), (112, 144), (120, 150), (128, 128)]      # 2 source + 2 target images, ragged


def small_args(device="cpu", **over):
    return dino_args(device=device, **{**SMALL, **over})


def seeded_state_dict(model, seed=1234):
    """Deterministic values for every entry of model.state_dict(), keyed by name order (numpy PCG64)."""
    rng = np.random.default_rng(seed)
    out = {}
    sd = model.state_dict()
    for k in sorted(sd):
        v = sd[k]
        if not v.dtype.is_floating_point:
            out[k] = v.clone()
            continue
        shape = tuple(v.shape)
        n = lambda s=1.0: torch.from_numpy((rng.standard_normal(shape) * s).astype(np.float32))
        if k.endswith("running_var"):
            t = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif k.endswith("running_mean"):
            t = n(0.1)
        elif ".bn" in k or "downsample.1" in k or "norm" in k or k.split(".")[-2].isdigit() and "input_proj" in k and ".1." in k:
            t = n(0.1) + (1.0 if k.endswith("weight") else 0.0)
        elif "sampling_offsets.bias" in k:
            t = torch.from_numpy(rng.uniform(-2.0, 2.0, shape).astype(np.float32))
        elif "sampling_offsets.weight" in k:
            t = n(0.02)
        elif "attention_weights" in k:
            t = n(0.05)
        elif k.endswith("level_embed") or "tgt_embed" in k or "label_enc" in k:
            t = n(0.5)
        elif v.dim() <= 1:
            t = n(0.05) - (2.0 if "class_embed" in k else 0.0)
        else:
            fan_in = int(np.prod(shape[1:]))
            t = n((2.0 if "backbone" in k else 1.0) ** 0.5 / fan_in ** 0.5)
        out[k] = t
    return out


def images(seed=3, sizes=IMAGE_SIZES):
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.standard_normal((3, h, w)).astype(np.float32)) for h, w in sizes]


def targets(seed=5, counts=(3, 5), num_classes=9, device="cpu"):
    rng = np.random.default_rng(seed)
    out = []
    for n in counts:
        cxcy = rng.uniform(0.2, 0.8, (n, 2))
        wh = rng.uniform(0.05, 0.35, (n, 2))
        out.append({"labels": torch.from_numpy(rng.integers(0, num_classes, n)).long().to(device),
                    "boxes": torch.from_numpy(np.concatenate([cxcy, wh], 1).astype(np.float32)).to(device)})
    return out


def pseudo_targets(device="cpu"):
    """Seeded pseudo labels for the two target-domain images (what engine.py:208-218 distils from the teacher)."""
    return targets(seed=9, counts=(2, 4), device=device)


def split_target_outputs(out, idx=(0, 1)):
    """The `*_target` half of a self-training output dict restricted to the images `idx` that have pseudo labels:
    spilt_output (self_training_utils.py:99-107) followed by get_valid_output (:110-146)."""
    idx = list(idx)
    pick = lambda d: {"pred_logits": d["pred_logits"][idx, :, :], "pred_boxes": d["pred_boxes"][idx, :, :]}
    res = {}
    for k, v in out.items():
        if "target" not in k:
            continue
        if "pred" in k:
            res[k] = v[idx, :, :]
        elif "aux_outputs_target" in k:
            res[k] = [pick(d) for d in v]
        else:
            res[k] = pick(v)
    return res


def flatten(tree, prefix=""):
    """dict / list / tensor tree -> {dotted name: tensor} (None and non-tensors skipped)."""
    flat = {}
    if isinstance(tree, torch.Tensor):
        flat[prefix] = tree
    elif isinstance(tree, dict):
        for k, v in tree.items():
            flat.update(flatten(v, f"{prefix}.{k}" if prefix else str(k)))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            flat.update(flatten(v, f"{prefix}[{i}]"))
    return flat


def total_loss(loss_dict, weight_dict):
    return sum(loss_dict[k] * weight_dict[k] for k in loss_dict if k in weight_dict)
