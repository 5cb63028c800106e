"""CPU (build container only: reads /root/reference): the REAL caller of the hot path -- the reference's own
engine.py::train_one_epoch (engine.py:29-143) and ::train_one_epoch_with_self_training (:146-330), unmodified -- driven
over datr_b200's drop-in model, criterion and post-processor for two iterations each.

What is the reference's: engine.py, util/misc.py (MetricLogger, reduce_dict, NestedTensor of the data loader),
util/utils.py, and every import they make.  What is ours: everything engine.py imports under `models.*`
(models.dino.dino.PostProcess, models.dino.self_training_utils.*, the model built through models.registry, the EMA teacher
of models.dino.EMA).  Packages the image lacks are stubbed at their import names only (pycocotools-based evaluators,
addict / yapf behind util.slconfig); the MSDeformAttn op runs the oracle's port of the reference's CPU path because the
product has no CPU implementation (like the reference's native op)."""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

import model_cases as mcase

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "engine.py")), reason="needs /root/reference")


@pytest.fixture(scope="module")
def engine():
    import datr_b200
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for k in list(sys.modules):                      # a reference `models` / `util` left by another test must not leak in
        if k.split(".")[0] in ("models", "util", "datasets", "engine"):
            del sys.modules[k]
    datr_b200.install_dropin()                       # `models.*` -> datr_b200.models.*
    stub = lambda name, **kw: sys.modules.setdefault(name, _mod(name, **kw))
    stub("addict", Dict=dict)
    stub("yapf"); stub("yapf.yapflib"); stub("yapf.yapflib.yapf_api", FormatCode=lambda s, **k: (s, False))
    stub("datasets"); stub("datasets.coco_eval", CocoEvaluator=object); stub("datasets.panoptic_eval", PanopticEvaluator=object)
    sys.path.insert(0, REF)                          # `util.*` and engine.py itself: the reference's
    try:
        spec = importlib.util.spec_from_file_location("engine", os.path.join(REF, "engine.py"))
        eng = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(eng)
        assert eng.PostProcess.__module__.startswith("datr_b200."), "engine.py must have bound OUR PostProcess"
        assert eng.spilt_output.__module__.startswith("datr_b200.")
        assert eng.utils.__file__.startswith(REF), "util.misc must be the reference's"
        # the op: oracle port of ms_deform_attn_core_pytorch on CPU (test infrastructure)
        from oracle import msda as om
        from datr_b200.models.dino.ops.modules import ms_deform_attn as mod

        class CpuFn:
            @staticmethod
            def apply(value, shapes, level_start, loc, attn, step):
                return om.core_torch(value, shapes, loc, attn)
        old_fn, mod.MSDeformAttnFunction = mod.MSDeformAttnFunction, CpuFn
        yield eng
        mod.MSDeformAttnFunction = old_fn
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k.split(".")[0] in ("models", "util", "datasets", "engine", "addict", "yapf") and k not in saved_mods:
                del sys.modules[k]


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    return m


def _loader(engine, n_iter, self_training=False):
    """What DAcoco + collate_fn_da hand to the loop (util/misc.py:291-300): a NestedTensor of source + target images,
    source label dicts, target label dicts, and (self-training) the strongly augmented batch."""
    nt = engine.utils.nested_tensor_from_tensor_list
    out = []
    for it in range(n_iter):
        imgs = mcase.images(seed=3 + it)
        samples = nt(imgs)
        labels = mcase.targets(seed=5 + it)
        h, w = samples.tensors.shape[-2:]
        tgt_labels = [{"image_id": torch.tensor([i]), "area": torch.zeros(0), "iscrowd": torch.zeros(0, dtype=torch.long),
                       "orig_size": torch.tensor([h, w]), "size": torch.tensor([h, w]),
                       "boxes": torch.zeros(0, 4), "labels": torch.zeros(0, dtype=torch.long)} for i in range(2)]
        out.append((samples, labels, tgt_labels, nt(imgs) if self_training else None))
    return out


def _build():
    from models.registry import MODULE_BUILD_FUNCS            # main.py:79-85
    import models  # noqa: F401
    torch.manual_seed(0)
    args = mcase.small_args()
    model, criterion, post = MODULE_BUILD_FUNCS.get(args.modelname)(args)
    model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
    for k, v in dict(amp=False, use_dn=True, onecyclelr=False, use_ema=False, debug=False, pseudo_label_threshold=0.0).items():
        setattr(args, k, v)
    from datr_b200.parallel import param_groups
    opt = torch.optim.AdamW(param_groups(model, 1e-4, 1e-5), lr=1e-4, weight_decay=1e-4)
    return args, model, criterion, opt


def test_reference_train_one_epoch_runs_over_the_dropin(engine):
    args, model, criterion, opt = _build()
    before = {k: v.detach().clone() for k, v in model.named_parameters() if v.requires_grad}
    stats = engine.train_one_epoch(model, criterion, _loader(engine, 2), opt, torch.device("cpu"), epoch=0, max_norm=0.1,
                                   wo_class_error=False, lr_scheduler=None, args=args, logger=None, ema_m=None)
    assert np.isfinite(stats["loss"]) and stats["loss"] > 0
    for k in ("loss_ce", "loss_bbox", "loss_giou", "loss_backbone_DA", "loss_proto_DA", "loss_global_proto_DA", "class_error", "lr"):
        assert k in stats, k
    moved = sum(int(not torch.equal(before[k], v.detach())) for k, v in model.named_parameters() if k in before)
    assert moved > 0.9 * len(before), "the optimizer step of engine.py must have updated the drop-in's parameters"


def test_reference_self_training_epoch_runs_over_the_dropin(engine, tmp_path):
    """engine.py:146-343 with the EMA teacher of OUR models.dino.EMA (main_teacher.py:292).  Threshold 0 lets every
    teacher detection through, so the pseudo-label path (threshold -> NMS -> rescale -> target-domain criterion) does
    real work."""
    from models.dino import EMA
    args, model, criterion, opt = _build()
    args.output_dir = str(tmp_path)                              # engine.py:331 appends its loss log there
    teacher = EMA.ModelEMA(model, decay=0.9997)
    stats = engine.train_one_epoch_with_self_training(model, teacher, criterion, _loader(engine, 2, True), None, opt,
                                                      torch.device("cpu"), epoch=0, max_norm=0.1, wo_class_error=False,
                                                      lr_scheduler=None, args=args, logger=None, ema_m=None)
    assert np.isfinite(stats["loss"]) and "sup_loss" in open(tmp_path / "loss_txt").read()
    teacher.update(model)                                      # main_teacher.py:384
    assert all(torch.isfinite(v).all() for v in teacher.ema.state_dict().values() if v.dtype.is_floating_point)
