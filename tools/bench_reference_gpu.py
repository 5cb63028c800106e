#!/usr/bin/env python
"""GPU baseline + full-size parity: the UNMODIFIED reference DINO (staged under baseline/_ref, imported by
tests/ref_loader.py) with its own MSDeformAttn CUDA extension (oracle/_ref/*.so, built unmodified for sm_100a)
run on the B200 next to datr_b200 on IDENTICAL weights, images, targets and de-noising noise.

  (a) reference step time, eager, following engine.py:54-111 (forward, SetCriterion, weighted sum, .item(),
      zero_grad, backward, clip_grad_norm_(0.1), AdamW): torch defaults (fp32 matmul, TF32 cuDNN), allow_tf32
      everywhere, and --amp (autocast fp16 + GradScaler);
  (b) parity of OUR step in the benchmarked mode (tf32 + NHWC + CUDA graphs) and in DATR_MATMUL=fp32 against the
      reference's strict-fp32 run: every loss, the two-stage top-k indices, the Hungarian assignments of all 7
      matchings (flips counted), gradient digests.

Usage (GPU box):  python tools/bench_reference_gpu.py [--config 4scale|5scale] [--steps 5] [--out profiles/x.json]
TEST / BASELINE INFRASTRUCTURE: nothing in datr_b200/ imports this."""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import model_cases as mcase  # noqa: E402
import ref_loader  # noqa: E402
from datr_b200.config import dino_5scale_args, dino_args  # noqa: E402


def set_precision(matmul_tf32, cudnn_tf32):
    torch.backends.cuda.matmul.allow_tf32 = matmul_tf32
    torch.backends.cudnn.allow_tf32 = cudnn_tf32


def batch(cfg, device, hw=None, num_classes=91):
    n_src = 2 if cfg == "4scale" else 1
    rng = np.random.default_rng(42)
    from datr_b200.bench_dino import H_IMG, W_IMG, synth_targets
    h, w = hw or (H_IMG, W_IMG)
    images = torch.from_numpy(rng.standard_normal((2 * n_src, 3, h, w)).astype(np.float32)).to(device)
    mask = torch.zeros((2 * n_src, h, w), dtype=torch.bool, device=device)
    targets = synth_targets(rng, n_src, num_classes, device)
    return images, mask, targets


def grads_of(model):
    return {k: p.grad.detach().double().cpu() for k, p in model.named_parameters() if p.grad is not None}


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def index_flips(a, b):
    """a, b: lists of (src, tgt) index pairs per image -> (#differing assignments, #total)."""
    diff = tot = 0
    for (sa, ta), (sb, tb) in zip(a, b):
        ma = dict(zip(ta.tolist(), sa.tolist()))
        mb = dict(zip(tb.tolist(), sb.tolist()))
        tot += len(mb)
        diff += sum(1 for k in mb if ma.get(k) != mb[k])
    return diff, tot


def run_parity_pass(model, crit, samples, targets, seed=7, keep_grad_views=False):
    """One training forward + criterion + backward with recorded index work.  Returns dict of python values.
    `keep_grad_views`: the .grad tensors are views of a FlatGradients buffer the captured backward graphs add into
    (datr_b200.graphs) -- they are zeroed in place instead of being dropped."""
    model.train(); crit.train()
    for p in model.parameters():
        if keep_grad_views and p.grad is not None:
            p.grad.zero_()
        else:
            p.grad = None
    torch.manual_seed(seed)
    out = model(samples, targets)
    # the two-stage top-k (deformable_transformer.py:342) gathers one proposal box per selected token, and every
    # token's proposal is unique (its own pixel centre and level size), so the gathered proposals identify the
    # selected indices and their order without hooking into the (possibly graph-captured) forward
    topk = [out["interm_outputs_for_matching_pre"]["pred_boxes"].detach().cpu()]
    losses = crit(out, targets)
    wd = crit.weight_dict
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    total.backward()
    with torch.no_grad():
        # all 7 matchings (final, 5 auxiliary, intermediate) on the same predictions
        sets = [{"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"]}] + \
               [dict(a) for a in out["aux_outputs"]] + [dict(out["interm_outputs"])]
        match = [[(s.cpu(), t.cpu()) for s, t in crit.matcher(s_, targets)] for s_ in sets]
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return {"losses": {k: float(v) for k, v in losses.items()}, "total": float(total), "topk": topk, "match": match,
            "grads": grads_of(model),
            "pred_logits": out["pred_logits"].detach().cpu(), "pred_boxes": out["pred_boxes"].detach().cpu()}


def compare(tag, ours, ref, report):
    worst, worst_k = 0.0, None
    for k, v in ref["losses"].items():
        e = abs(ours["losses"][k] - v) / max(1.0, abs(v))
        if e > worst:
            worst, worst_k = e, k
    tk_diff = tk_tot = 0
    quant = lambda t: [set(map(tuple, (img * 1e5).round().long().tolist())) for img in t]
    for a, b in zip(ours["topk"], ref["topk"]):
        if a.shape == b.shape:
            tk_tot += b.shape[0] * b.shape[1]
            for sa, sb in zip(quant(a), quant(b)):
                tk_diff += len(sb - sa)
    tk_order_equal = all(a.shape == b.shape and torch.allclose(a, b, atol=2e-6, rtol=0) for a, b in zip(ours["topk"], ref["topk"]))
    tk_pos = sum(int(((a - b).abs().amax(-1) > 2e-6).sum()) for a, b in zip(ours["topk"], ref["topk"]) if a.shape == b.shape)
    m_diff = m_tot = 0
    for a, b in zip(ours["match"], ref["match"]):
        d, t = index_flips(a, b)
        m_diff += d; m_tot += t
    gworst, gk, gn = 0.0, None, 0
    gerrs = []
    for k, g in ref["grads"].items():
        if k in ours["grads"]:
            e = rel(ours["grads"][k], g)
            gerrs.append((e, k))
            gn += 1
            if e > gworst:
                gworst, gk = e, k
    gerrs.sort(reverse=True)
    # all gradients as one vector: relative L2 error and cosine (insensitive to which near-tied query slot a token took)
    num = sum(float(((ours["grads"][k] - g) ** 2).sum()) for k, g in ref["grads"].items() if k in ours["grads"])
    den = sum(float((g ** 2).sum()) for k, g in ref["grads"].items() if k in ours["grads"])
    dot = sum(float((ours["grads"][k] * g).sum()) for k, g in ref["grads"].items() if k in ours["grads"])
    no = sum(float((ours["grads"][k] ** 2).sum()) for k in ref["grads"] if k in ours["grads"])
    report[tag] = {
        "total_loss": {"ours": ours["total"], "reference": ref["total"],
                       "rel": abs(ours["total"] - ref["total"]) / abs(ref["total"])},
        "worst_loss_rel": {"key": worst_k, "rel": worst, "n_losses": len(ref["losses"])},
        "two_stage_topk": {"calls": len(ref["topk"]), "indices": tk_tot, "membership_flips": tk_diff,
                           "positions_with_another_token": tk_pos, "identical_incl_order": bool(tk_order_equal)},
        "hungarian": {"matchings": len(ref["match"]), "assignments": m_tot, "flips": m_diff},
        "pred_logits_rel": rel(ours["pred_logits"], ref["pred_logits"]),
        "pred_boxes_rel": rel(ours["pred_boxes"], ref["pred_boxes"]),
        "worst_grad_rel": {"key": gk, "rel": gworst, "n_params": gn,
                           "missing_in_ours": sorted(set(ref["grads"]) - set(ours["grads"]))[:5]},
        "grads": {"median_rel": gerrs[len(gerrs) // 2][0] if gerrs else None,
                  "n_above_1e-2": sum(1 for e, _ in gerrs if e > 1e-2), "n_above_1e-3": sum(1 for e, _ in gerrs if e > 1e-3),
                  "worst5": [(k, round(e, 5)) for e, k in gerrs[:5]],
                  "global_rel_l2": (num / max(den, 1e-300)) ** 0.5, "global_cosine": dot / max((den * no) ** 0.5, 1e-300)},
    }
    print(f"[parity:{tag}]", json.dumps(report[tag]), flush=True)


def time_reference(model, crit, opt, samples, targets, steps, warmup, amp):
    """engine.py:54-111, one process, eager."""
    scaler = torch.amp.GradScaler("cuda", enabled=amp)
    model.train(); crit.train()

    def one():
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            out = model(samples, targets)
            loss_dict = crit(out, targets)
            wd = crit.weight_dict
            losses = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
        value = losses.item()                                   # engine.py:76 (logging sync)
        opt.zero_grad()
        if amp:
            scaler.scale(losses).backward()
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
            scaler.step(opt); scaler.update()
        else:
            losses.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
            opt.step()
        return value
    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        v = one()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="4scale", choices=["4scale", "5scale"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-timing", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--weights", default="seeded", choices=["seeded", "init"],
                    help="seeded: tests/model_cases.seeded_state_dict (large random weights, chaotic: every near-tie flips); "
                         "init: the reference's own initialisation under torch.manual_seed(0) (what training starts from)")
    ap.add_argument("--reference-only", action="store_true", help="time / check the reference model only")
    ap.add_argument("--dry-run", action="store_true", help="CPU, toy configuration: exercises this script's logic only")
    a = ap.parse_args()
    dry = a.dry_run
    stack = contextlib.ExitStack()
    if dry:
        dev = torch.device("cpu")
        a.skip_timing = True
        small = dict(mcase.SMALL)
        if a.config == "5scale":
            small.update(mcase.FIVE_SCALE)
        mk = lambda device: mcase.dino_args(device="cpu", **small)
        images, mask, targets = batch(a.config, dev, hw=(128, 160), num_classes=9)
        stack.enter_context(ref_loader.cpu_cuda_shim())
    else:
        assert torch.cuda.is_available()
        dev = torch.device("cuda", 0)
        mk = dino_args if a.config == "4scale" else dino_5scale_args
        images, mask, targets = batch(a.config, dev)
    bs = 2 if a.config == "4scale" else 1
    n_images = images.shape[0]
    report = {"config": a.config, "images_per_step": n_images,
              "gpu": torch.cuda.get_device_name(0) if not dry else "cpu dry run", "torch": torch.__version__,
              "reference": "baseline/_ref (unmodified Python tree) + oracle/_ref CUDA MSDeformAttn (unmodified, sm_100a)"}

    # ---------------- reference ----------------
    ns = ref_loader.load(cuda_ext=not dry)
    torch.manual_seed(0)
    ref_model, ref_crit, _ = ns.dino.build_dino(mk(device="cuda"))
    ref_model.to(dev); ref_crit.to(dev)
    if a.weights == "seeded":
        sd = mcase.seeded_state_dict(ref_model)
    else:
        sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    ref_model.load_state_dict(sd, strict=True)
    report["weights"] = a.weights
    ref_samples = ns.misc.NestedTensor(images, mask)
    ref_runs = {}
    if not a.skip_parity:
        set_precision(False, False)                           # strict fp32 reference
        ref_model.global_proto = torch.zeros_like(ref_model.global_proto); ref_model.Amount = torch.zeros_like(ref_model.Amount)
        ref_runs["fp32"] = run_parity_pass(ref_model, ref_crit, ref_samples, targets)
        print("[reference] strict fp32 total loss", ref_runs["fp32"]["total"], flush=True)
        set_precision(False, True)                            # torch defaults: what a user of the reference gets
        ref_model.global_proto = torch.zeros_like(ref_model.global_proto); ref_model.Amount = torch.zeros_like(ref_model.Amount)
        ref_runs["default"] = run_parity_pass(ref_model, ref_crit, ref_samples, targets)
        set_precision(True, True)                             # the reference with TF32 GEMMs too (fair twin of our tf32 mode)
        ref_model.global_proto = torch.zeros_like(ref_model.global_proto); ref_model.Amount = torch.zeros_like(ref_model.Amount)
        ref_runs["tf32"] = run_parity_pass(ref_model, ref_crit, ref_samples, targets)
        rep = {}
        compare("reference_default_vs_reference_fp32", ref_runs["default"], ref_runs["fp32"], rep)
        compare("reference_allow_tf32_vs_reference_fp32", ref_runs["tf32"], ref_runs["fp32"], rep)
        report["reference_self_noise"] = rep
    if not a.skip_timing:
        from datr_b200.parallel import param_groups
        timing = {}
        for name, (mm, cd, amp) in {"torch_default(fp32 matmul, tf32 cudnn)": (False, True, False),
                                    "allow_tf32": (True, True, False), "amp_fp16": (True, True, True)}.items():
            set_precision(mm, cd)
            ref_model.load_state_dict(sd, strict=True)
            opt = torch.optim.AdamW(param_groups(ref_model, 1e-4, 1e-5), lr=1e-4, weight_decay=1e-4)
            try:
                ms, v = time_reference(ref_model, ref_crit, opt, ref_samples, targets, a.steps, a.warmup, amp)
                timing[name] = {"ms_per_step": ms, "images_per_s": n_images / ms * 1e3, "last_loss": v}
            except Exception as e:  # noqa: BLE001
                timing[name] = {"error": repr(e)[:300]}
            print("[reference timing]", name, timing[name], flush=True)
            del opt
        report["reference_step"] = timing
    del ref_model
    stack.close()
    if not dry:
        torch.cuda.empty_cache()

    if a.reference_only:
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            json.dump(report, open(a.out, "w"), indent=1)
        print(json.dumps(report))
        return

    # ---------------- ours ----------------
    from datr_b200 import bench_dino, linear as dl
    from datr_b200.util.misc import NestedTensor
    over = {} if a.config == "4scale" else {"return_interm_indices": [0, 1, 2, 3], "num_feature_levels": 5}
    if dry:
        over = dict(small, height=128, width=160)
        from oracle import msda as om
        from datr_b200.models.dino.ops.modules import ms_deform_attn as mod

        class CpuFn:
            @staticmethod
            def apply(value, shapes, level_start, loc, attn, step):
                return om.core_torch(value, shapes, loc, attn)
        mod.MSDeformAttnFunction = CpuFn
    ours = {}
    for mode in ("tf32", "fp32"):
        os.environ["DATR_MATMUL"] = mode
        wl = bench_dino.DinoStep(dev, batch_size=bs, **over)
        wl.model.load_state_dict(sd, strict=True)
        if mode == "fp32":
            set_precision(False, False)
        from datr_b200.models.dino import dn_components as _dn
        if not a.skip_parity:
            _dn.SYNC_FREE = False          # parity pass: the reference's own sequence of de-noising draws
            wl.model.global_proto = None
            samples = NestedTensor(wl.images.copy_(images), wl.mask.copy_(mask))
            wl.grads.zero()
            r = run_parity_pass(wl.model, wl.criterion, samples, targets, keep_grad_views=True)
            ours[mode] = r
            compare(f"ours_{mode}_vs_reference_fp32", r, ref_runs["fp32"], report)
        _dn.SYNC_FREE = True
        if mode == "tf32" and not a.skip_timing:
            if not wl.grads.check_views():
                wl.grads = __import__("datr_b200.parallel", fromlist=["FlatGradients"]).FlatGradients(wl.model, late=lambda n: n.startswith("backbone"))
            wl.model._on_backbone_output_grad = wl.grads.reduce_early
            _groups = __import__("datr_b200.parallel", fromlist=["param_groups"]).param_groups(wl.model, 1e-4, 1e-5)
            if getattr(wl, "flat_opt", False):
                wl.opt = __import__("datr_b200.optim", fromlist=["FlatAdamW"]).FlatAdamW(_groups, wl.grads, weight_decay=1e-4)
            else:
                wl.opt = torch.optim.AdamW(_groups, lr=1e-4, weight_decay=1e-4, fused=True)
            wl.images.copy_(images)
            for _ in range(4):
                wl.step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                wl.step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            report["ours_step"] = {"mode": "tf32 + NHWC + CUDA graphs", "ms_per_step": ms,
                                   "images_per_s": n_images / ms * 1e3, "last_loss": float(wl.last_loss)}
            print("[ours timing]", report["ours_step"], flush=True)
        wl.set_graphs(False)
        del wl
        if not dry:
            torch.cuda.empty_cache()
    dl.set_mode("fp32")
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(report, f, indent=1)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
