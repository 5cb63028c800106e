"""Full-model workload of bench.py: one DINO-4scale ResNet-50 DA training step (BASELINE.json configs[1]/[2]).

Per GPU and step: 2 source + 2 target synthetic 1333x800 images -> backbone -> input projections -> transformer
(source pass with de-noising queries, target pass) -> heads + DA branch -> SetCriterion -> backward -> one flat
gradient all-reduce -> clip (0.1) -> AdamW.  fp32 parameters and activations; library convolutions run with
cuDNN's default TF32 policy exactly as the reference does on an Ampere+ GPU."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from datr_b200.config import dino_args
from datr_b200.parallel import FlatGradients, broadcast_parameters, param_groups
from datr_b200.util.misc import NestedTensor

H_IMG, W_IMG = 800, 1333
WORKLOAD = ("DINO-4scale ResNet-50 DA training step (forward + losses + backward + gradient all-reduce + clip + AdamW), "
            "synthetic 1333x800, batch_size 2/GPU = 2 source + 2 target images, 900 queries + CDN; fp32 tensors, "
            "MSDeformAttn fp32, dense layers TF32 products with fp32 accumulation, encoder FFN GEMMs bf16 operands with fp32 accumulation")


def synth_targets(rng, n_images, num_classes, device):
    out = []
    for _ in range(n_images):
        k = int(rng.integers(1, 21))
        cxcy = rng.uniform(0.2, 0.8, (k, 2))
        wh = rng.uniform(0.05, 0.5, (k, 2))
        boxes = np.clip(np.concatenate([cxcy, wh], 1), 0.01, 0.99).astype(np.float32)
        out.append({"labels": torch.from_numpy(rng.integers(0, num_classes, k)).long().to(device),
                    "boxes": torch.from_numpy(boxes).to(device)})
    return out


class DinoStep:
    name = "dino"
    workload = WORKLOAD

    def __init__(self, device, rank=0, world=1, batch_size=2, height=H_IMG, width=W_IMG, **over):
        from datr_b200.models.dino.dino import build_dino
        self.device, self.rank, self.world = device, rank, world
        # dense contractions run on the tensor cores with TF32 operands / fp32 accumulation (10-bit mantissa, above the
        # bf16 floor BASELINE.json allows); DATR_MATMUL=fp32 restores SIMT fp32 GEMMs (what the parity tests use)
        self.matmul = os.environ.get("DATR_MATMUL", "tf32")
        from datr_b200 import linear as _dl
        tensor_core = self.matmul == "tf32" and device.type == "cuda"
        self.dtype = ("tf32+bf16" if _dl._FFN == "bf16" else "tf32") if tensor_core else "f32"
        torch.backends.cuda.matmul.allow_tf32 = self.matmul == "tf32"
        torch.backends.cudnn.allow_tf32 = True
        # static shapes: let cuDNN time its algorithms once for the layers that stay on it (7x7 stem, convolution
        # backward); DATR_CUDNN_BENCHMARK=0 keeps its heuristics
        torch.backends.cudnn.benchmark = os.environ.get("DATR_CUDNN_BENCHMARK", "1") != "0"
        from datr_b200 import linear as dl
        dl.set_mode("tf32" if self.matmul == "tf32" and device.type == "cuda" else "fp32")
        # CUDA graphs for the ResNet body, encoder and decoder (datr_b200/graphs.py); DATR_GRAPHS=0 keeps the step eager
        self.graphs = None
        if device.type == "cuda" and os.environ.get("DATR_GRAPHS", "1") != "0":
            from datr_b200 import graphs
            self.graphs = graphs.StepGraphs()
        self.set_graphs(True)
        os.environ.setdefault("DATR_BACKBONE_WEIGHTS", "none")  # synthetic benchmark: random-init weights by design
        torch.manual_seed(42)                                   # identical initial weights on every rank
        args = dino_args(device=str(device), **over)
        self.args = args
        self.model, self.criterion, _ = build_dino(args)
        self.model.to(device).train()
        if device.type == "cuda" and os.environ.get("DATR_CHANNELS_LAST", "1") != "0":
            # NHWC activations/weights: cuDNN's tensor-core convolutions run without NCHW<->NHWC transposes, and
            # flattening a feature map to [N, HW, C] tokens becomes a view
            self.model.to(memory_format=torch.channels_last)
        self.criterion.train()
        self.criterion.fold_weighted_sum = True
        broadcast_parameters(self.model)
        # backbone gradients are produced last: they sit at the end of the flat buffer, and the exchange of everything
        # else starts from a backward hook while the ResNet backward is still running (parallel.py)
        self.grads = FlatGradients(self.model, late=lambda n: n.startswith("backbone"))
        self.model._on_backbone_output_grad = self.grads.reduce_early
        # clip + AdamW as ONE kernel over all parameters (datr_b200.optim, include/datr_adamw.h); DATR_OPTIMIZER=torch keeps
        # torch.nn.utils.clip_grad_norm_-style scaling + torch.optim.AdamW(fused=True)
        self.flat_opt = device.type == "cuda" and os.environ.get("DATR_OPTIMIZER", "flat") == "flat"
        if self.flat_opt:
            from datr_b200.optim import FlatAdamW
            self.opt = FlatAdamW(param_groups(self.model, args.lr, args.lr_backbone), self.grads, weight_decay=args.weight_decay)
        else:
            self.opt = torch.optim.AdamW(param_groups(self.model, args.lr, args.lr_backbone), lr=args.lr,
                                         weight_decay=args.weight_decay, fused=device.type == "cuda")
        rng = np.random.default_rng(42 + rank)                  # main.py:138: seed + rank
        n = 2 * batch_size
        self.host_images = torch.from_numpy(rng.standard_normal((n, 3, height, width)).astype(np.float32))
        self.host_mask = torch.zeros((n, height, width), dtype=torch.bool)
        if device.type == "cuda":
            self.host_images, self.host_mask = self.host_images.pin_memory(), self.host_mask.pin_memory()
        self.host_targets = synth_targets(rng, batch_size, args.num_classes, "cpu")
        self.channels_last = device.type == "cuda" and os.environ.get("DATR_CHANNELS_LAST", "1") != "0"
        self.images = self.host_images.to(device)
        if self.channels_last:
            self.images = self.images.contiguous(memory_format=torch.channels_last)
        self.mask = self.host_mask.to(device)
        self.targets = [{k: v.to(device) for k, v in t.items()} for t in self.host_targets]
        self.n_images = n
        self.last_loss = None
        self._loss_w, self._loss_keys = None, None
        torch.manual_seed(1000 + rank)                          # CDN noise stream

    def set_graphs(self, on: bool):
        from datr_b200 import graphs
        graphs.ACTIVE = self.graphs if on else None

    def _step(self, images, mask, targets):
        if self.graphs is not None:
            self.graphs.begin_step()
        self.grads.zero()
        out = self.model(NestedTensor(images, mask), targets)
        losses = self.criterion(out, targets)
        # engine.py:99 `sum(loss_dict[k] * weight_dict[k] ...)` as one stack + dot (2 kernels instead of ~120)
        if "_weighted_total" in losses:       # the criterion segment already summed the weighted losses (graph replay)
            loss = losses["_weighted_total"]
            loss.backward()
            self.grads.all_reduce()
            self._optimise()
            return loss
        wd = self.criterion.weight_dict
        keys = [k for k in losses if k in wd]
        if self._loss_w is None or self._loss_keys != keys:
            self._loss_keys = keys
            self._loss_w = torch.tensor([wd[k] for k in keys], dtype=torch.float32, device=self.device)
        loss = torch.dot(torch.stack([losses[k].reshape(()) for k in keys]), self._loss_w)
        loss.backward()
        self.grads.all_reduce()
        self._optimise()
        return loss

    def _optimise(self):
        """engine.py:108-111: clip_grad_norm_(max_norm) + optimizer.step() on the all-reduced gradients."""
        if self.flat_opt:
            self.opt.clip_and_step(self.args.clip_max_norm)
        else:
            self.grads.clip_(self.args.clip_max_norm)
            self.opt.step()

    def step(self):
        self.last_loss = self._step(self.images, self.mask, self.targets)

    def e2e_step(self):
        """Host batch in (pinned -> device copies inside the step), loss value out (device -> host read)."""
        images = self.host_images.to(self.device, non_blocking=True)
        if self.channels_last:
            images = images.contiguous(memory_format=torch.channels_last)
        mask = self.host_mask.to(self.device, non_blocking=True)
        targets = [{k: v.to(self.device, non_blocking=True) for k, v in t.items()} for t in self.host_targets]
        loss = self._step(images, mask, targets)
        value = loss.item()
        assert np.isfinite(value), "loss diverged"
        h2d = self.host_images.numel() * 4 + self.host_mask.numel() + sum(v.numel() * v.element_size() for t in self.host_targets for v in t.values())
        return h2d, 4

    def roofline(self, timers, peak, peak_src):
        import bench
        return bench.MsdaStep.roofline(self, timers, peak, peak_src)

    def extra(self):
        return {"loss": float(self.last_loss.detach()) if self.last_loss is not None else None,
                "grad_allreduce_bytes": self.grads.numel * 4,
                "cuda_graphs": None if self.graphs is None else {"segments": self.graphs.captures,
                                                                 "what": "ResNet body, input projections, flatten, encoder (source + target halves in one call), two-stage selection x2, decoder x2, heads, image discriminator, criterion losses (forward + backward each)"}}


WORKLOAD_5SCALE = ("DINO-5scale ResNet-50 DA training step (forward + losses + backward + gradient all-reduce + clip + AdamW), "
                   "synthetic 1333x800, batch_size 1/GPU = 1 source + 1 target image, 5 feature levels (S = 89023 tokens), "
                   "900 queries + CDN; fp32 tensors, MSDeformAttn fp32, dense layers TF32 products with fp32 accumulation")
WORKLOAD_TEACHER = ("Teacher-student mutual-learning step of main_teacher.py / engine.py:146-300 on DINO-4scale ResNet-50: EMA "
                    "teacher eval pass on the 2 target images -> PostProcess -> thresholded + NMS pseudo labels; student "
                    "DA pass on 2 source + 2 target images (self_training_flag) -> source criterion + target-domain criterion on "
                    "the pseudo labels -> backward -> gradient all-reduce -> clip -> AdamW -> teacher EMA update; synthetic "
                    "1333x800, batch_size 2/GPU; fp32 tensors, dense layers TF32 products with fp32 accumulation")


class Dino5Step(DinoStep):
    """BASELINE.json configs[3]: DINO-5scale (return_interm_indices [0,1,2,3], 5 feature levels), batch_size 1 per GPU."""
    name = "dino5"
    workload = WORKLOAD_5SCALE

    def __init__(self, device, rank=0, world=1, **over):
        super().__init__(device, rank, world, batch_size=1, return_interm_indices=[0, 1, 2, 3], num_feature_levels=5, **over)


class TeacherStep(DinoStep):
    """BASELINE.json configs[4]: the self-training step (engine.py:196-300) followed by the teacher's EMA update
    (main_teacher.py:384; the reference updates the teacher once per epoch, here it is part of EVERY step, i.e. the
    step does at least the reference's work).  A randomly initialised teacher scores every query ~0.01, so the
    reference's threshold 0.3 (config/DA/*_self_training.py:124) would leave the target-domain criterion without work;
    the benchmark threshold is 0: all PostProcess detections pass, class-wise NMS then keeps <= 100 pseudo boxes per
    target image -- an upper bound on the pseudo-label work of a real run."""
    name = "teacher"
    workload = WORKLOAD_TEACHER

    def __init__(self, device, rank=0, world=1, batch_size=2, **over):
        super().__init__(device, rank, world, batch_size=batch_size, **over)
        from datr_b200.models.dino import EMA
        from datr_b200.models.dino.dino import PostProcess
        self.teacher = EMA.ModelEMA(self.model, decay=0.9997)            # main_teacher.py:292, ema_decay_teacher
        self.post = PostProcess(num_select=self.args.num_select, nms_iou_threshold=self.args.nms_iou_threshold)
        self.threshold = np.asarray([0.0] * self.args.num_classes)
        h, w = self.host_images.shape[-2:]
        size = torch.tensor([h, w], device=device)
        # target-domain label dicts: only their bookkeeping keys are read (self_training_utils.py:52-67)
        self.target_labels = [{"image_id": torch.tensor([i], device=device), "area": torch.zeros(0, device=device),
                               "iscrowd": torch.zeros(0, dtype=torch.long, device=device), "orig_size": size, "size": size}
                              for i in range(batch_size)]
        self.unit_sizes = torch.ones((batch_size, 2), dtype=torch.long, device=device)
        self.n_pseudo = 0

    def _step(self, images, mask, targets):
        from datr_b200.models.dino import self_training_utils as st
        from datr_b200 import graphs
        if self.graphs is not None:
            self.graphs.begin_step()
        self.grads.zero()
        samples = NestedTensor(images, mask)
        # 1. teacher on the (weakly augmented) target images                               engine.py:196-204
        unlabel = st.get_unlabel_img(samples)
        with torch.no_grad():                                             # forward-only graph segments (graphs.py)
            pred = self.teacher.ema(unlabel)
            results = self.post(pred, self.unit_sizes, not_to_xyxy=True)                   # :205-207
        # 2. pseudo labels                                                                  :210-216
        idx_list, labels_d, boxes_d, scores_d = st.get_pseudo_label_via_threshold(results, threshold=self.threshold)
        pseudo = st.deal_pesudo_label(self.target_labels, idx_list, labels_d, boxes_d, scores_d)
        pseudo = st.rescale_pseudo_targets(unlabel, pseudo)
        # 3. student on the whole batch                                                     :221-225
        out = self.model(samples, targets, self_training_flag=True)
        source_out, target_out = st.spilt_output(out)                                       # :232
        valid_out, pseudo_list = st.get_valid_output(target_out, pseudo, idx_list)          # :235
        self.n_pseudo = sum(int(t["labels"].shape[0]) for t in pseudo_list)
        wd = self.criterion.weight_dict
        loss_src = self.criterion(source_out, targets, target_domain_flag=False)           # :240
        loss_tgt = self.criterion(valid_out, pseudo_list, target_domain_flag=True)         # :243
        total_src = loss_src["_weighted_total"] if "_weighted_total" in loss_src else sum(loss_src[k] * wd[k] for k in loss_src if k in wd)
        total_tgt = sum(loss_tgt[k] * wd[k] for k in loss_tgt if k in wd)
        loss = total_src + total_tgt * wd["loss_self_training"]                             # :257
        loss.backward()
        self.grads.all_reduce()
        self._optimise()
        self.teacher.update(self.model)                                                     # main_teacher.py:384
        return loss

    def extra(self):
        d = super().extra()
        d["pseudo_boxes_last_step"] = self.n_pseudo
        return d


def _reference_model_cpu(batch_size, over):
    """The UNMODIFIED reference model (baseline/_ref, staged by oracle/stage_ref.py; /root/reference in the build
    container) on the host cores, MSDeformAttn served by the reference's own CPU path ms_deform_attn_core_pytorch
    (ops/functions/ms_deform_attn_func.py:41-61).  Returns (model, criterion, make_samples) or None."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tests = os.path.join(root, "tests")
    if tests not in sys.path:
        sys.path.insert(0, tests)
    try:
        import ref_loader
        if not ref_loader.available():
            return None
        ns = ref_loader.load(cuda_ext=False)
    except Exception as e:      # noqa: BLE001
        print(f"[bench] reference model unavailable ({e!r}); falling back to the port", file=sys.stderr)
        return None
    return ns, ref_loader


def reference_arm(args, threads, workload="dino"):
    """CPU arm (`bench.py --impl reference`, and the `cpu_baseline` of the default run): the reference's own CPU
    implementation of the path -- the unmodified reference DINO model + SetCriterion from baseline/_ref with its
    pure-PyTorch MSDeformAttn (kind "reference"); if the staged tree is missing, datr_b200's mirror of the model with the
    oracle's port of that op (kind "port").  Every step is a BOUNDED SAMPLE of the GPU arm's workload: the same training
    step (forward, losses, backward, clip, AdamW; engine.py:54-111) on 1 source + 1 target image instead of the
    2 + 2 (4-scale) of a GPU step, all host threads.  Runs exactly --warmup + --steps steps."""
    torch.set_num_threads(threads)
    five = workload == "dino5"
    over = {"return_interm_indices": [0, 1, 2, 3], "num_feature_levels": 5} if five else {}
    ref = _reference_model_cpu(1, over)
    n_images = 2
    rng = np.random.default_rng(42)
    images = torch.from_numpy(rng.standard_normal((n_images, 3, H_IMG, W_IMG)).astype(np.float32))
    mask = torch.zeros((n_images, H_IMG, W_IMG), dtype=torch.bool)
    targets = synth_targets(rng, 1, 91, "cpu")
    if ref is not None:
        ns, ref_loader = ref
        kind = "reference"
        os.environ.setdefault("DATR_BACKBONE_WEIGHTS", "none")
        torch.manual_seed(42)
        with ref_loader.cpu_cuda_shim():
            model, criterion, _ = ns.dino.build_dino(dino_args(device="cpu", **over))
        model.train(); criterion.train()
        opt = torch.optim.AdamW(param_groups(model, 1e-4, 1e-5), lr=1e-4, weight_decay=1e-4)
        samples = ns.misc.NestedTensor(images, mask)

        def one():
            with ref_loader.cpu_cuda_shim():                     # the reference hard-codes .cuda() in its training path
                out = model(samples, targets)
                loss_dict = criterion(out, targets)
            wd = criterion.weight_dict
            loss = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
            opt.step()
            return float(loss)
        what = "the UNMODIFIED reference DINO + SetCriterion (baseline/_ref) with ms_deform_attn_core_pytorch"
    else:
        kind = "port"
        from oracle import msda as om
        from datr_b200.models.dino.ops.modules import ms_deform_attn as mod

        class CpuFn:
            @staticmethod
            def apply(value, shapes, level_start, loc, attn, step):
                return om.core_torch(value, shapes, loc, attn)
        mod.MSDeformAttnFunction = CpuFn
        wl = DinoStep(torch.device("cpu"), batch_size=1, **over)

        def one():
            wl.step()
            return float(wl.last_loss)
        what = "datr_b200's mirror of the model on torch CPU ops with the oracle's port of ms_deform_attn_core_pytorch"
    for _ in range(max(0, args.warmup)):
        one()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        last = one()
    t = (time.perf_counter() - t0) / max(1, args.steps)
    ips = n_images / t
    sample = (f"{max(1, args.steps)} timed steps (+{max(0, args.warmup)} warm-up) of the same training step on 1 source + 1 target "
              f"1333x800 image (a GPU step has {'1 + 1' if five else '2 + 2'}), {what}, {threads} host threads; last loss {last:.3f}")
    return {"impl": "reference", "metric": "images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
