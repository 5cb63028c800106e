"""Generate tests/golden/selftrain_golden.npz from the REFERENCE's own code (build container only): the target-domain
half of the mutual-learning step (BASELINE.json configs[4]; engine.py:196-260).

With the small seeded model of tests/model_cases.py in training mode and self_training_flag=True:
  * the student's outputs are split into source / target dicts (self_training_utils.py:99-107), the target dict is
    reduced to the images that have pseudo labels (:110-146) -- here both -- and scored by the reference's
    SetCriterion with target_domain_flag=True against seeded pseudo labels; every loss, the weighted total and the
    gradient digests are recorded;
  * the teacher-side post-processing of engine.py:203-206, PostProcess(..., not_to_xyxy=True) on unit image sizes.
Usage:  python tests/golden/make_selftrain_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.dirname(HERE)]
import model_cases as mcase  # noqa: E402
import ref_loader  # noqa: E402

STRIDE = 101


def main():
    ns = ref_loader.load()
    G = {}
    torch.manual_seed(0)
    with ref_loader.cpu_cuda_shim():
        model, crit, post = ns.dino.build_dino(mcase.small_args())
    model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
    imgs = mcase.images()
    model.train(); crit.train()
    model.global_proto = torch.zeros_like(model.global_proto); model.Amount = torch.zeros_like(model.Amount)
    torch.manual_seed(7)
    with ref_loader.cpu_cuda_shim():
        out = model(ns.misc.nested_tensor_from_tensor_list(imgs), mcase.targets(), self_training_flag=True)
        target_out = mcase.split_target_outputs(out)                 # == spilt_output + get_valid_output with idx [0, 1]
        pseudo = mcase.pseudo_targets()
        losses = crit(target_out, pseudo, target_domain_flag=True)
    for k, v in losses.items():
        G[f"loss.{k}"] = v.detach().numpy().copy()
    total = mcase.total_loss(losses, crit.weight_dict)
    G["total"] = total.detach().numpy()
    model.zero_grad()
    total.backward()
    for k, p in model.named_parameters():
        if p.grad is not None:
            G[f"grad_sub.{k}"] = p.grad.reshape(-1)[::STRIDE].numpy().copy()
    # teacher-side: eval forward of the target images, normalised cxcywh boxes (engine.py:201-206)
    model.eval()
    with torch.no_grad():
        pred = model(ns.misc.nested_tensor_from_tensor_list(imgs[2:]))
        res = post["bbox"](pred, torch.ones(2, 2), not_to_xyxy=True)
    for i, r in enumerate(res):
        for k, v in r.items():
            G[f"teacher[{i}].{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "selftrain_golden.npz"), **G)
    print("wrote selftrain_golden.npz", len(G), "arrays")


if __name__ == "__main__":
    main()
