"""Generate tests/golden/fivescale_golden.npz from the REFERENCE's own model code (build container only): the small
seeded model of tests/model_cases.py at the 5-scale configuration of BASELINE.json configs[3]
(return_interm_indices [0,1,2,3], num_feature_levels 5: C2..C5 + one extra stride-2 level) -- eval-mode outputs, and
the training step's losses and weighted total (CDN with torch.manual_seed(7)).
Usage:  python tests/golden/make_5scale_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.dirname(HERE)]
import model_cases as mcase  # noqa: E402
import ref_loader  # noqa: E402


def main():
    ns = ref_loader.load()
    G = {}
    torch.manual_seed(0)
    with ref_loader.cpu_cuda_shim():
        model, crit, post = ns.dino.build_dino(mcase.small_args(**mcase.FIVE_SCALE))
    model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
    G["keys"] = np.array([f"{k}|{','.join(map(str, v.shape))}" for k, v in model.state_dict().items()])
    imgs = mcase.images()
    model.eval()
    with torch.no_grad():
        out = model(ns.misc.nested_tensor_from_tensor_list(imgs))
    for k, v in mcase.flatten(out).items():
        G["eval." + k] = v.numpy().copy()
    model.train(); crit.train()
    model.global_proto = torch.zeros_like(model.global_proto); model.Amount = torch.zeros_like(model.Amount)
    torch.manual_seed(7)
    with ref_loader.cpu_cuda_shim():
        out = model(ns.misc.nested_tensor_from_tensor_list(imgs), mcase.targets())
        losses = crit(out, mcase.targets())
    for k, v in losses.items():
        G[f"train.loss.{k}"] = v.detach().numpy().copy()
    G["train.total"] = mcase.total_loss(losses, crit.weight_dict).detach().numpy()
    for k in ("pred_logits", "pred_boxes"):
        G["train." + k] = out[k].detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "fivescale_golden.npz"), **G)
    print("wrote fivescale_golden.npz", len(G), "arrays")


if __name__ == "__main__":
    main()
