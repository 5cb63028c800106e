"""CPU: host logic of the models/dino package against golden vectors produced by the REFERENCE's own model
code (tests/golden/make_model_golden.py) and, in the build container, against the reference run live.

The MSDeformAttn op has no CPU implementation in the product (and none in the reference's extension), so for
these host-logic tests -- and only here -- the autograd function is swapped for the oracle's grid_sample port.
The CUDA path is exercised by tests/test_model_gpu.py."""
import os

import numpy as np
import pytest
import torch

import model_cases as mcase
import ref_loader
from oracle import msda as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))


@pytest.fixture(scope="module")
def cpu_op():
    from datr_b200.models.dino.ops.modules import ms_deform_attn as mod

    class OracleFn:
        @staticmethod
        def apply(value, shapes, level_start, loc, attn, step):
            return om.core_torch(value, shapes, loc, attn)
    saved = mod.MSDeformAttnFunction
    mod.MSDeformAttnFunction = OracleFn
    yield
    mod.MSDeformAttnFunction = saved


@pytest.fixture(scope="module")
def small(cpu_op):
    from datr_b200.models.dino.dino import build_dino
    torch.manual_seed(0)
    model, crit, post = build_dino(mcase.small_args())
    model.load_state_dict(mcase.seeded_state_dict(model), strict=True)
    return model, crit, post


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def finite_rel(a, b):
    """like rel() but +/-inf must coincide (proposal logits use +inf as 'invalid')."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isinf(a), np.isinf(b))
    m = np.isfinite(b)
    return rel(a[m], b[m]) if m.any() else 0.0


@pytest.mark.parametrize("tag,builder", [("4scale", mcase.dino_args),
                                          ("5scale", lambda **k: mcase.dino_args(return_interm_indices=[0, 1, 2, 3], num_feature_levels=5, **k))])
def test_state_dict_keys_shapes_and_trainable_set_match_reference(G, tag, builder):
    """Checkpoint compatibility: 640 tensors at DINO-4scale, same names, shapes and frozen set."""
    from datr_b200.models.dino.dino import build_dino
    with torch.device("meta"):
        model = build_dino(builder(device="cpu"))[0]
    mine = [f"{k}|{','.join(map(str, v.shape))}" for k, v in model.state_dict().items()]
    assert mine == list(G[f"keys_{tag}"])
    assert [k for k, p in model.named_parameters() if p.requires_grad] == list(G[f"trainable_{tag}"])
    if tag == "4scale":
        assert len(mine) == 640
        assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 47836480   # 191.3 MB fp32 allreduce


def test_weight_dict_matches_reference(G, small):
    _, crit, _ = small
    assert sorted(crit.weight_dict) == list(G["weight_dict_keys"])
    np.testing.assert_allclose([crit.weight_dict[k] for k in sorted(crit.weight_dict)], G["weight_dict_vals"])


def test_eval_forward_and_postprocess_match_reference(G, small):
    model, _, post = small
    model.eval()
    with torch.no_grad():
        out = model(mcase.images())
        res = post["bbox"](out, torch.tensor([[h, w] for h, w in mcase.IMAGE_SIZES], dtype=torch.float32))
    flat = mcase.flatten(out)
    keys = [k[5:] for k in G.files if k.startswith("eval.")]
    assert sorted(flat) == sorted(keys)
    for k in keys:
        assert rel(flat[k].numpy(), G["eval." + k]) < 2e-5, k
    for i, r in enumerate(res):
        assert np.array_equal(r["labels"].numpy(), G[f"post[{i}].labels"])          # top-k index work: bit-exact
        assert rel(r["scores"].numpy(), G[f"post[{i}].scores"]) < 1e-5
        assert rel(r["boxes"].numpy(), G[f"post[{i}].boxes"]) < 1e-5


@pytest.mark.parametrize("flag", [False, True])
def test_train_forward_losses_and_gradients_match_reference(G, small, flag):
    """DA training step: CDN queries (same RNG stream), both transformer passes, discriminator, prototypes, all
    losses, Hungarian indices (bit-exact) and the gradient of the weighted loss for every parameter."""
    model, crit, _ = small
    tag = "train_st" if flag else "train"
    model.train(); crit.train()
    model.global_proto = None
    torch.manual_seed(7)
    out = model(mcase.images(), mcase.targets(), self_training_flag=flag)
    losses = crit(out, mcase.targets())
    flat = mcase.flatten(out)
    skip = tuple(f"{tag}.{s}" for s in ("loss.", "grad_", "total", "global_proto", "Amount", "match["))
    keys = [k[len(tag) + 1:] for k in G.files if k.startswith(tag + ".") and not k.startswith(skip)]
    assert sorted(flat) == sorted(keys)
    for k in keys:
        assert finite_rel(flat[k].detach().numpy(), G[f"{tag}.{k}"]) < 5e-5, k
    loss_keys = [k.split(".loss.")[1] for k in G.files if k.startswith(tag + ".loss.")]
    assert sorted(losses) == sorted(loss_keys)
    for k in loss_keys:
        assert abs(float(losses[k]) - float(G[f"{tag}.loss.{k}"])) < 1e-4 * max(1.0, abs(float(G[f"{tag}.loss.{k}"]))), k
    total = mcase.total_loss(losses, crit.weight_dict)
    assert abs(float(total) - float(G[f"{tag}.total"])) < 1e-4 * abs(float(G[f"{tag}.total"]))
    with torch.no_grad():
        idx = crit.matcher({"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"]}, mcase.targets())
    for i, (a, b) in enumerate(idx):
        assert np.array_equal(a.numpy(), G[f"{tag}.match[{i}].src"]) and np.array_equal(b.numpy(), G[f"{tag}.match[{i}].tgt"])
    model.zero_grad()
    total.backward()
    with_grad = {k for k, p in model.named_parameters() if p.grad is not None}
    assert with_grad == {k.split("grad_digest.")[1] for k in G.files if k.startswith(tag + ".grad_digest.")}
    for k, p in model.named_parameters():
        if p.grad is not None:
            sub = G[f"{tag}.grad_sub.{k}"]
            got = p.grad.reshape(-1)[::101].numpy()
            scale = max(np.abs(sub).max(), 1e-6)
            assert np.abs(got - sub).max() / scale < 2e-3, k
    assert rel(model.global_proto.numpy(), G[f"{tag}.global_proto"]) < 1e-5
    assert np.array_equal(model.Amount.numpy(), G[f"{tag}.Amount"])


def test_module_level_pieces_match_reference(G, small):
    from datr_b200.models.dino.deformable_transformer import TransformerEncoder
    from datr_b200.models.dino.utils import gen_encoder_output_proposals, gen_sineembed_for_position
    model, _, _ = small
    model.eval()
    levels = [(8, 10), (4, 5), (2, 3), (1, 2)]
    S = sum(h * w for h, w in levels)
    rng = np.random.default_rng(11)
    src = torch.from_numpy(rng.standard_normal((2, S, 256)).astype(np.float32))
    pos = torch.from_numpy(rng.standard_normal((2, S, 256)).astype(np.float32))
    shapes = torch.tensor(levels)
    lstart = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    vr = torch.from_numpy(rng.uniform(0.7, 1.0, (2, 4, 2)).astype(np.float32))
    mask = torch.zeros(2, S, dtype=torch.bool); mask[1, -3:] = True
    with torch.no_grad():
        ref2 = TransformerEncoder.get_reference_points(levels, vr, device="cpu")
        assert rel(ref2.numpy(), G["mod.ref2"]) < 1e-6
        enc0 = model.transformer.encoder.layers[0]
        assert rel(enc0(src, pos, ref2, shapes, lstart, mask).numpy(), G["mod.enc_layer"]) < 2e-5
        assert rel(enc0.self_attn(src + pos, ref2, src, shapes, lstart, mask).numpy(), G["mod.msda_2d"]) < 2e-5
        q = torch.from_numpy(rng.standard_normal((2, 7, 256)).astype(np.float32))
        ref4 = torch.from_numpy(rng.uniform(0.1, 0.9, (2, 7, 4, 4)).astype(np.float32))
        got = model.transformer.decoder.layers[0].cross_attn(q, ref4, src, shapes, lstart, mask)
        assert rel(got.numpy(), G["mod.msda_4d"]) < 2e-5
        assert rel(gen_sineembed_for_position(ref4[:, :, 0, :]).numpy(), G["mod.sine4"]) < 1e-5
        om_, op_ = gen_encoder_output_proposals(src, mask, shapes)
        assert rel(om_.numpy(), G["mod.prop_memory"]) < 1e-6
        assert finite_rel(op_.numpy(), G["mod.prop_boxes"]) < 1e-5


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")
def test_live_against_reference_modules(cpu_op):
    """Runs the reference's own classes next to ours on the same seeded inputs (build container only)."""
    ns = ref_loader.load()
    from datr_b200.models.dino import DA_utils, dn_components, matcher, position_encoding
    from datr_b200.util import box_ops, misc
    g = torch.Generator().manual_seed(3)
    # CDN: identical RNG stream, identical queries and mask
    label_enc = torch.nn.Embedding(10, 256)
    tg = mcase.targets(counts=(4, 0, 2))
    torch.manual_seed(11)
    with ref_loader.cpu_cuda_shim():
        want = ns.dn.prepare_for_cdn((tg, 100, 0.5, 0.4), True, 30, 9, 256, label_enc)
    torch.manual_seed(11)
    got = dn_components.prepare_for_cdn((tg, 100, 0.5, 0.4), True, 30, 9, 256, label_enc)
    for a, b in zip(got[:3], want[:3]):
        assert torch.equal(a, b)
    assert got[3] == want[3]
    # matcher: bit-exact assignment
    outputs = {"pred_logits": torch.randn(3, 30, 9, generator=g), "pred_boxes": torch.rand(3, 30, 4, generator=g) * 0.5 + 0.2}
    args = mcase.small_args()
    mi, ri = matcher.build_matcher(args)(outputs, tg), ns.matcher.build_matcher(args)(outputs, tg)
    for (a, b), (c, d) in zip(mi, ri):
        assert torch.equal(a, c) and torch.equal(b, d)
    # box ops
    b1 = box_ops.box_cxcywh_to_xyxy(torch.rand(7, 4, generator=g) * 0.4 + 0.3)
    b2 = box_ops.box_cxcywh_to_xyxy(torch.rand(5, 4, generator=g) * 0.4 + 0.3)
    assert torch.allclose(box_ops.generalized_box_iou(b1, b2), ns.box_ops.generalized_box_iou(b1, b2), atol=1e-6)
    assert torch.allclose(box_ops.paired_giou(b1[:5], b2), torch.diag(ns.box_ops.generalized_box_iou(b1[:5], b2)), atol=1e-6)
    # position encoding on a ragged mask
    nt = misc.nested_tensor_from_tensor_list([torch.randn(3, 20, 31, generator=g), torch.randn(3, 17, 25, generator=g)])
    rnt = ns.misc.nested_tensor_from_tensor_list([nt.tensors[0, :, :20, :31], nt.tensors[1, :, :17, :25]])
    assert torch.equal(nt.mask, rnt.mask) and torch.equal(nt.tensors, rnt.tensors)
    pe, rpe = position_encoding.build_position_encoding(args), ns.posenc.build_position_encoding(args)
    assert torch.allclose(pe(nt), rpe(rnt), atol=1e-6)
    # prototypes
    feats, logits = torch.randn(2, 30, 256, generator=g), torch.randn(2, 30, 9, generator=g)
    gp, ga = torch.randn(9, 256, generator=g), torch.randint(0, 5, (9,), generator=g).float()
    for a, b in zip(DA_utils.get_prototype_class_wise(feats, logits, 9, gp.clone(), ga.clone()),
                    ns.da.get_prototype_class_wise(feats, logits, 9, gp.clone(), ga.clone())):
        assert torch.allclose(a.float(), b.float(), atol=1e-5)
    assert misc.inverse_sigmoid(torch.tensor([0.0, 0.3, 1.0])).equal(ns.misc.inverse_sigmoid(torch.tensor([0.0, 0.3, 1.0])))
