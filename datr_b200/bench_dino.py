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
            "MSDeformAttn fp32, dense layers TF32 products with fp32 accumulation")


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
        self.dtype = "tf32" if self.matmul == "tf32" and device.type == "cuda" else "f32"
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
        broadcast_parameters(self.model)
        self.grads = FlatGradients(self.model)
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
        wd = self.criterion.weight_dict
        keys = [k for k in losses if k in wd]
        if self._loss_w is None or self._loss_keys != keys:
            self._loss_keys = keys
            self._loss_w = torch.tensor([wd[k] for k in keys], dtype=torch.float32, device=self.device)
        loss = torch.dot(torch.stack([losses[k].reshape(()) for k in keys]), self._loss_w)
        loss.backward()
        self.grads.all_reduce()
        self.grads.clip_(self.args.clip_max_norm)
        self.opt.step()
        return loss

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
                                                                 "what": "ResNet body, encoder x2, two-stage selection x2, decoder x2, heads, image discriminator, criterion losses (forward + backward each)"}}


def reference_arm(args, threads):
    """CPU arm: the same training step run on the host cores with the MSDeformAttn op served by the oracle's port of
    the reference's grid_sample CPU path (func.py:41-61).  Bounded sample: 1 source + 1 target image per step."""
    import json  # noqa: F401
    from oracle import msda as om
    from datr_b200.models.dino.ops.modules import ms_deform_attn as mod

    class CpuFn:
        @staticmethod
        def apply(value, shapes, level_start, loc, attn, step):
            return om.core_torch(value, shapes, loc, attn)
    mod.MSDeformAttnFunction = CpuFn
    torch.set_num_threads(threads)
    wl = DinoStep(torch.device("cpu"), batch_size=1)
    times = []
    for i in range(max(1, min(args.steps, 2)) + (1 if args.warmup > 0 else 0)):
        t0 = time.perf_counter()
        wl.step()
        times.append(time.perf_counter() - t0)
    t = min(times[1:] or times)
    ips = wl.n_images / t
    sample = ("1 source + 1 target 1333x800 image through the same DINO-4scale DA training step on the host cores "
              "(torch CPU ops; MSDeformAttn = port of the reference's grid_sample path, oracle/msda.py)")
    return {"impl": "reference", "metric": "images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
