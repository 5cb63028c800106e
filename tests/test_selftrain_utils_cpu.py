"""CPU: the pseudo-label plumbing mirror (datr_b200/models/dino/self_training_utils.py) against the reference's own
models/dino/self_training_utils.py imported live in the build container, on seeded predictions: thresholding (incl.
thresholds whose float32 rounding lies below the float64 value), label formatting, NMS + rescaling, output
splitting.  Bit-exact (same torch ops, same library NMS)."""
import copy

import numpy as np
import pytest
import torch

import ref_loader
from datr_b200.models.dino import self_training_utils as ours
from datr_b200.util.misc import NestedTensor


def predictions(seed=0, n_images=3, n=60, num_classes=9):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_images):
        scores = torch.from_numpy(rng.uniform(0, 1, n).astype(np.float32))
        if i == 0:
            scores[:5] = torch.tensor(0.7)                   # float32(0.7) < 0.7: the float64 comparison rejects these
        if i == 2:
            scores = scores * 0.2                            # nothing survives on the last image
        cxcy = rng.uniform(0.3, 0.7, (n, 2)); wh = rng.uniform(0.05, 0.4, (n, 2))
        out.append({"scores": scores, "labels": torch.from_numpy(rng.integers(0, num_classes, n)).long(),
                    "boxes": torch.from_numpy(np.concatenate([cxcy, wh], 1).astype(np.float32))})
    return out


def unlabeled_targets(n_images=3):
    return [{"image_id": torch.tensor([10 + i]), "area": torch.tensor([1.0]), "iscrowd": torch.tensor([0]),
             "orig_size": torch.tensor([480, 640]), "size": torch.tensor([100 + 4 * i, 140 + 2 * i]),
             "labels": torch.tensor([1]), "boxes": torch.tensor([[0.5, 0.5, 0.2, 0.2]])} for i in range(n_images)]


def tree_equal(a, b):
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(tree_equal(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(tree_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, torch.Tensor):
        return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    return a == b


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    ns = ref_loader.load()
    if ns.selftrain is None:
        pytest.skip("the reference's self_training_utils needs cv2 / torchvision")
    return ns.selftrain


def test_pseudo_label_pipeline_is_bit_identical_to_the_reference(ref):
    threshold = np.asarray([0.7] * 9)
    imgs = torch.zeros(3, 3, 128, 160)
    a = ours.get_pseudo_label_via_threshold(predictions(), threshold=threshold)
    b = ref.get_pseudo_label_via_threshold(predictions(), threshold=threshold)
    assert a[0] == b[0] == [0, 1] and tree_equal(a[1:], b[1:])
    assert not bool((a[3][0] == torch.tensor(0.7)).any())            # the float32(0.7) scores were rejected, as in the reference
    pa = ours.deal_pesudo_label(unlabeled_targets(), *a)
    pb = ref.deal_pesudo_label(unlabeled_targets(), *b)
    assert tree_equal(pa, pb)
    ra = ours.rescale_pseudo_targets(imgs, copy.deepcopy(pa))
    rb = ref.rescale_pseudo_targets(imgs, copy.deepcopy(pb))
    assert tree_equal(ra, rb)
    assert all(len(t["labels"]) <= 100 for t in ra.values())


def test_output_splitting_matches_the_reference(ref):
    g = torch.Generator().manual_seed(0)
    mk = lambda: {"pred_logits": torch.randn(3, 7, 9, generator=g), "pred_boxes": torch.rand(3, 7, 4, generator=g)}
    out = dict(mk(), aux_outputs=[mk()], interm_outputs=mk(), dn_meta=None,
               pred_logits_target=mk()["pred_logits"], pred_boxes_target=mk()["pred_boxes"], aux_outputs_target=[mk(), mk()],
               interm_outputs_target=mk(), interm_outputs_for_matching_pre_target=mk())
    sa, ta = ours.spilt_output(out)
    sb, tb = ref.spilt_output(out)
    assert sa.keys() == sb.keys() and ta.keys() == tb.keys()
    labels = {0: {"labels": torch.tensor([1])}, 2: {"labels": torch.tensor([3, 4])}}
    va, la = ours.get_valid_output(ta, labels, [0, 2])
    vb, lb = ref.get_valid_output(tb, labels, [0, 2])
    assert tree_equal(va, vb) and tree_equal(la, lb)
    import model_cases as mcase
    assert tree_equal(mcase.split_target_outputs(out, idx=(0, 2)), vb)


def test_unlabelled_half_and_dropin_import():
    nt = NestedTensor(torch.arange(4 * 3 * 2 * 2, dtype=torch.float32).view(4, 3, 2, 2), torch.zeros(4, 2, 2, dtype=torch.bool))
    assert torch.equal(ours.get_unlabel_img(nt), nt.tensors[2:])
    import datr_b200
    datr_b200.install_dropin()
    from models.dino.self_training_utils import (deal_pesudo_label, get_pseudo_label_via_threshold, get_unlabel_img,  # noqa: F401
                                                 get_valid_output, rescale_pseudo_targets, show_pesudo_label_with_gt,
                                                 spilt_output)
    from models.dino import EMA  # noqa: F401
    from models.dino.dino import PostProcess  # noqa: F401


def test_scalar_threshold_and_empty_results():
    res = ours.get_pseudo_label_via_threshold(predictions(seed=1, n_images=1), threshold=2.0)
    assert res == ([], {}, {}, {})
    res = ours.get_pseudo_label_via_threshold(predictions(seed=1, n_images=1), threshold=0.5)
    assert res[0] == [0] and bool((res[3][0] >= 0.5).all())
