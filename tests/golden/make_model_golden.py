"""Generate tests/golden/model_golden.npz from the REFERENCE's own model code (build container only).

Builds the reference DINO (models/dino/dino.py:999 build_dino, via tests/ref_loader.py) at the small
configuration of tests/model_cases.py, loads the numpy-seeded weights, and records on CPU:
  * the list of state_dict keys and shapes at the small AND the full DINO-4scale / 5scale configuration,
  * eval-mode outputs, PostProcess results,
  * training-mode outputs (DA branch, CDN with torch.manual_seed(7)), every loss of SetCriterion, and
    the gradient of the weighted total loss w.r.t. every parameter (digest: sum, abs-sum, and a strided
    subsample), with and without self_training_flag,
  * module-level outputs: MSDeformAttn forward (2-d and 4-d reference points), one encoder layer,
    the helper functions (sine embeddings, proposals, CDN, matcher indices, prototypes).
The MSDeformAttn op inside the reference runs its own pure-PyTorch path (func.py:41-61).

Usage:  python tests/golden/make_model_golden.py
"""
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


def digest(t):
    a = t.detach().double().reshape(-1)
    return np.array([a.sum().item(), a.abs().sum().item(), (a * a).sum().item()])


def main():
    ns = ref_loader.load()
    G = {}
    torch.manual_seed(0)
    with ref_loader.cpu_cuda_shim():
        args = mcase.small_args()
        model, crit, post = ns.dino.build_dino(args)
        for tag, a in (("4scale", mcase.dino_args(device="cpu")),
                       ("5scale", mcase.dino_args(device="cpu", return_interm_indices=[0, 1, 2, 3], num_feature_levels=5))):
            with torch.device("meta"):
                full = ns.dino.build_dino(a)[0]
            sd = full.state_dict()
            G[f"keys_{tag}"] = np.array([f"{k}|{','.join(map(str, v.shape))}" for k, v in sd.items()])
            G[f"trainable_{tag}"] = np.array([k for k, p in full.named_parameters() if p.requires_grad])
    sd = model.state_dict()
    G["keys_small"] = np.array([f"{k}|{','.join(map(str, v.shape))}" for k, v in sd.items()])
    G["weight_dict_keys"] = np.array(sorted(crit.weight_dict))
    G["weight_dict_vals"] = np.array([crit.weight_dict[k] for k in sorted(crit.weight_dict)], dtype=np.float64)
    model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
    imgs = mcase.images()

    # ---- eval ----
    model.eval()
    with torch.no_grad():
        out = model(ns.misc.nested_tensor_from_tensor_list(imgs))
        res = post["bbox"](out, torch.tensor([[h, w] for h, w in mcase.IMAGE_SIZES], dtype=torch.float32))
    for k, v in mcase.flatten(out).items():
        G["eval." + k] = v.numpy().copy()
    for i, r in enumerate(res):
        for k, v in r.items():
            G[f"post[{i}].{k}"] = v.numpy()

    # ---- train (+ criterion + grads) ----
    model.train(); crit.train()
    for flag in (False, True):
        tag = "train_st" if flag else "train"
        model.global_proto = torch.zeros_like(model.global_proto); model.Amount = torch.zeros_like(model.Amount)
        torch.manual_seed(7)
        with ref_loader.cpu_cuda_shim():
            out = model(ns.misc.nested_tensor_from_tensor_list(imgs), mcase.targets(), self_training_flag=flag)
            losses = crit(out, mcase.targets())
        for k, v in mcase.flatten(out).items():
            G[f"{tag}.{k}"] = v.detach().numpy().copy()
        for k, v in losses.items():
            G[f"{tag}.loss.{k}"] = v.detach().numpy().copy()
        total = mcase.total_loss(losses, crit.weight_dict)
        G[f"{tag}.total"] = total.detach().numpy()
        model.zero_grad()
        total.backward()
        for k, p in model.named_parameters():
            if p.grad is not None:
                G[f"{tag}.grad_digest.{k}"] = digest(p.grad)
                G[f"{tag}.grad_sub.{k}"] = p.grad.reshape(-1)[::STRIDE].numpy().copy()
        G[f"{tag}.global_proto"] = model.global_proto.numpy().copy()
        G[f"{tag}.Amount"] = model.Amount.numpy().copy()
        # matcher indices of the final layer (bit-exact index work)
        with torch.no_grad():
            idx = crit.matcher({"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"]}, mcase.targets())
        for i, (a, b) in enumerate(idx):
            G[f"{tag}.match[{i}].src"], G[f"{tag}.match[{i}].tgt"] = a.numpy(), b.numpy()

    # ---- module level: MSDeformAttn + one encoder layer of the seeded model ----
    enc0 = model.transformer.encoder.layers[0]
    levels = [(8, 10), (4, 5), (2, 3), (1, 2)]
    S = sum(h * w for h, w in levels)
    rng = np.random.default_rng(11)
    src = torch.from_numpy(rng.standard_normal((2, S, 256)).astype(np.float32))
    pos = torch.from_numpy(rng.standard_normal((2, S, 256)).astype(np.float32))
    shapes = torch.tensor(levels)
    lstart = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    vr = torch.from_numpy(rng.uniform(0.7, 1.0, (2, 4, 2)).astype(np.float32))
    mask = torch.zeros(2, S, dtype=torch.bool); mask[1, -3:] = True
    ref2 = ns.transformer.TransformerEncoder.get_reference_points(shapes, vr, device="cpu")
    with torch.no_grad():
        G["mod.ref2"] = ref2.numpy()
        G["mod.enc_layer"] = enc0(src, pos, ref2, shapes, lstart, mask).numpy()
        G["mod.msda_2d"] = enc0.self_attn(src + pos, ref2, src, shapes, lstart, mask).numpy()
        q = torch.from_numpy(rng.standard_normal((2, 7, 256)).astype(np.float32))
        ref4 = torch.from_numpy(rng.uniform(0.1, 0.9, (2, 7, 4, 4)).astype(np.float32))
        G["mod.msda_4d"] = model.transformer.decoder.layers[0].cross_attn(q, ref4, src, shapes, lstart, mask).numpy()
        G["mod.sine4"] = ns.utils.gen_sineembed_for_position(ref4[:, :, 0, :]).numpy()
        om, op = ns.utils.gen_encoder_output_proposals(src, mask, shapes)
        G["mod.prop_memory"], G["mod.prop_boxes"] = om.numpy(), op.numpy()
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **G)
    print("wrote model_golden.npz", os.path.getsize(os.path.join(HERE, "model_golden.npz")), "bytes", len(G), "arrays")


if __name__ == "__main__":
    main()
