"""DINO detector with DATR's domain-adaptation branch, its criterion and post-processor.

Mirrors the reference's models/dino/dino.py: DINO (:43-483), SetCriterion (:486-941), PostProcess (:944-996)
and the registry entry build_dino (:999-1143) -- same constructor arguments, sub-module names (state_dict
keys), forward signature `model(samples, targets=None, self_training_flag=False)` and output dict keys, so
engine.py / main.py / main_teacher.py of the reference drive it unchanged.

Hot-path differences that keep the numbers:
  * nothing is hard-wired to .cuda(): helper tensors are created on the device of the parameters
    (the reference calls .cuda()/.to('cuda') at :106-107, :790-818);
  * the de-noising queries are prepared before the backbone is launched, so their one data-dependent host
    read happens while the GPU queue is empty;
  * box losses use the paired GIoU instead of the diagonal of an N x N matrix (:563-565);
  * segmentation heads (DETRsegm) are outside the hot path (`masks=False` in every DINO/DATR config).
"""
import copy
import os
import math
from typing import List

import torch
import torch.nn.functional as F
from torch import nn

from datr_b200 import graphs
from datr_b200.util import box_ops
from datr_b200.util.misc import (NestedTensor, accuracy, get_world_size, inverse_sigmoid, upload,
                                 is_dist_avail_and_initialized, nested_tensor_from_tensor_list)
from ..registry import MODULE_BUILD_FUNCS
from .backbone import build_backbone
from .DA_utils import FCDiscriminator_img, decompose_features, get_prototype_class_wise, grad_reverse
from .deformable_transformer import build_deformable_transformer
from .dn_components import dn_post_process, prepare_for_cdn
from .matcher import BatchedMatch, batchable, build_matcher, match_many, matching_sets, prefetch, take_prefetched
from .utils import MLP, sigmoid_focal_loss

_JOINT_ENCODER = os.environ.get("DATR_JOINT_ENCODER", "1") != "0"
_JOINT_DECODER = os.environ.get("DATR_JOINT_DECODER", "1") != "0"


_OWN_GROUPNORM = os.environ.get("DATR_OWN_GROUPNORM", "1") != "0"


class DINO(nn.Module):
    """Backbone -> input projections -> (CDN queries) -> deformable transformer -> class / box heads,
    plus in training mode the image-level discriminator, class prototypes and a second transformer pass
    over the target-domain half of the batch."""

    def __init__(self, backbone, transformer, num_classes, num_queries, aux_loss=False, iter_update=False,
                 query_dim=2, random_refpoints_xy=False, fix_refpoints_hw=-1, num_feature_levels=1, nheads=8,
                 two_stage_type="no", two_stage_add_query_num=0, dec_pred_class_embed_share=True,
                 dec_pred_bbox_embed_share=True, two_stage_class_embed_share=True, two_stage_bbox_embed_share=True,
                 decoder_sa_type="sa", num_patterns=0, dn_number=100, dn_box_noise_scale=0.4,
                 dn_label_noise_ratio=0.5, dn_labelbook_size=100):
        super().__init__()
        assert query_dim == 4
        assert iter_update, "Why not iter_update?"
        assert two_stage_type in ("no", "standard"), f"unknown param {two_stage_type} of two_stage_type"
        assert decoder_sa_type in ("sa", "ca_label", "ca_content")
        self.num_queries, self.transformer, self.num_classes = num_queries, transformer, num_classes
        self.hidden_dim = hidden_dim = transformer.d_model
        self.num_feature_levels, self.nheads = num_feature_levels, nheads
        self.label_enc = nn.Embedding(dn_labelbook_size + 1, hidden_dim)
        self.query_dim, self.random_refpoints_xy, self.fix_refpoints_hw = query_dim, random_refpoints_xy, fix_refpoints_hw
        self.num_patterns, self.dn_number = num_patterns, dn_number
        self.dn_box_noise_scale, self.dn_label_noise_ratio = dn_box_noise_scale, dn_label_noise_ratio
        self.dn_labelbook_size = dn_labelbook_size

        # domain-adaptation heads (a fresh module is in training mode, so the reference always builds them, :102-108)
        self.D_img = FCDiscriminator_img(256)
        self.global_proto = None          # [num_classes, 256] running class prototypes, created on first use;
        self.Amount = None                # plain attributes like the reference: not saved, not rank-synchronised
        self.Proto_D = MLP(hidden_dim, hidden_dim, 1, 3)

        if num_feature_levels > 1:
            n_backbone = len(backbone.num_channels)
            proj = [nn.Sequential(nn.Conv2d(c, hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim))
                    for c in backbone.num_channels]
            c_in = backbone.num_channels[-1]
            for _ in range(num_feature_levels - n_backbone):
                proj.append(nn.Sequential(nn.Conv2d(c_in, hidden_dim, kernel_size=3, stride=2, padding=1),
                                          nn.GroupNorm(32, hidden_dim)))
                c_in = hidden_dim
            self.input_proj = nn.ModuleList(proj)
        else:
            assert two_stage_type == "no", "two_stage_type should be no if num_feature_levels=1 !!!"
            self.input_proj = nn.ModuleList([nn.Sequential(nn.Conv2d(backbone.num_channels[-1], hidden_dim, kernel_size=1),
                                                           nn.GroupNorm(32, hidden_dim))])
        self.backbone = backbone
        self.aux_loss = aux_loss
        self.box_pred_damping = None
        self.iter_update = iter_update

        self.dec_pred_class_embed_share, self.dec_pred_bbox_embed_share = dec_pred_class_embed_share, dec_pred_bbox_embed_share
        cls_head = nn.Linear(hidden_dim, num_classes)
        box_head = MLP(hidden_dim, hidden_dim, 4, 3)
        prior = 0.01
        cls_head.bias.data = torch.ones(num_classes) * (-math.log((1 - prior) / prior))
        nn.init.constant_(box_head.layers[-1].weight.data, 0)
        nn.init.constant_(box_head.layers[-1].bias.data, 0)
        n_dec = transformer.num_decoder_layers
        self.bbox_embed = nn.ModuleList([box_head if dec_pred_bbox_embed_share else copy.deepcopy(box_head) for _ in range(n_dec)])
        self.class_embed = nn.ModuleList([cls_head if dec_pred_class_embed_share else copy.deepcopy(cls_head) for _ in range(n_dec)])
        self.transformer.decoder.bbox_embed = self.bbox_embed
        self.transformer.decoder.class_embed = self.class_embed

        self.two_stage_type, self.two_stage_add_query_num = two_stage_type, two_stage_add_query_num
        if two_stage_type != "no":
            if two_stage_bbox_embed_share:
                assert dec_pred_class_embed_share and dec_pred_bbox_embed_share
                self.transformer.enc_out_bbox_embed = box_head
            else:
                self.transformer.enc_out_bbox_embed = copy.deepcopy(box_head)
            if two_stage_class_embed_share:
                assert dec_pred_class_embed_share and dec_pred_bbox_embed_share
                self.transformer.enc_out_class_embed = cls_head
            else:
                self.transformer.enc_out_class_embed = copy.deepcopy(cls_head)
            self.refpoint_embed = None
            if two_stage_add_query_num > 0:
                self.init_ref_points(two_stage_add_query_num)

        self.decoder_sa_type = decoder_sa_type
        self.label_embedding = nn.Embedding(num_classes, hidden_dim) if decoder_sa_type == "ca_label" else None
        for layer in self.transformer.decoder.layers:
            layer.label_embedding = self.label_embedding
        self._reset_parameters()

    def _reset_parameters(self):
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)

    def init_ref_points(self, use_num_queries):
        self.refpoint_embed = nn.Embedding(use_num_queries, self.query_dim)
        if self.random_refpoints_xy:
            self.refpoint_embed.weight.data[:, :2].uniform_(0, 1)
            self.refpoint_embed.weight.data[:, :2] = inverse_sigmoid(self.refpoint_embed.weight.data[:, :2])
            self.refpoint_embed.weight.data[:, :2].requires_grad = False
        if self.fix_refpoints_hw > 0:
            assert self.random_refpoints_xy
            self.refpoint_embed.weight.data[:, 2:] = self.fix_refpoints_hw
            self.refpoint_embed.weight.data[:, 2:] = inverse_sigmoid(self.refpoint_embed.weight.data[:, 2:])
            self.refpoint_embed.weight.data[:, 2:].requires_grad = False
        elif int(self.fix_refpoints_hw) == -2:
            assert self.random_refpoints_xy
            self.refpoint_embed = nn.Embedding(use_num_queries, 2)
            self.refpoint_embed.weight.data[:, :2].uniform_(0, 1)
            self.refpoint_embed.weight.data[:, :2] = inverse_sigmoid(self.refpoint_embed.weight.data[:, :2])
            self.refpoint_embed.weight.data[:, :2].requires_grad = False
            self.hw_embed = nn.Embedding(1, 1)
        elif int(self.fix_refpoints_hw) != -1:
            raise NotImplementedError(f"Unknown fix_refpoints_hw {self.fix_refpoints_hw}")

    # ------------------------------------------------------------------------------------------
    def _features(self, samples: NestedTensor):
        """Backbone + input projections (+ the extra stride-2 levels): per level (src, mask, pos)."""
        if graphs.ACTIVE is not None and samples.tensors.is_cuda:
            base = self.backbone[0]
            feats = graphs.ACTIVE.run("body", lambda: graphs.BodySegment(base.body), (samples.tensors,), owner=base.body)
            self._watch_backbone_output_grads(feats)
            srcs, masks, poss = graphs.ACTIVE.call("project", self.input_proj, self._project, tuple(feats), samples.mask)
            return list(srcs), list(masks), list(poss)
        features, poss = self.backbone(samples)
        self._watch_backbone_output_grads([f.tensors for f in features])
        return self._project_levels(features, poss, samples.mask)

    def _watch_backbone_output_grads(self, feats):
        """Data-parallel hook (datr_b200.parallel.FlatGradients.reduce_early): `self._on_backbone_output_grad`, if set, is
        called during the backward pass at the moment the gradients of ALL backbone feature maps exist -- every
        parameter outside the backbone has its final gradient then, and only the ResNet backward is still to run, so
        the caller can start exchanging the first part of the gradients underneath it."""
        cb = getattr(self, "_on_backbone_output_grad", None)
        feats = [f for f in feats if f.requires_grad]
        if cb is None or not feats or not torch.is_grad_enabled():
            return
        pending = [len(feats)]

        def fire(_grad):
            pending[0] -= 1
            if pending[0] == 0:
                cb()
        for f in feats:
            f.register_hook(fire)

    def _project(self, feats, mask):
        """Everything between the ResNet body and the transformer as a pure tensor function: mask down-sampling
        (BackboneBase.forward), sine position encoding (Joiner) and the input projections."""
        features = [NestedTensor(x, F.interpolate(mask[None].float(), size=x.shape[-2:]).to(torch.bool)[0]) for x in feats]
        poss = [self.backbone[1](f).to(f.tensors.dtype) for f in features]
        srcs, masks, poss = self._project_levels(features, poss, mask)
        return tuple(srcs), tuple(masks), tuple(poss)

    @staticmethod
    def _input_proj(proj, x):
        """One input projection (reference dino.py:111-126: Conv2d 1x1, or 3x3 stride 2 for the extra levels, + GroupNorm).
        On NHWC CUDA tensors in tensor-core mode the 1x1 convolution is a GEMM over pixels on the tcgen05 linear kernel
        and the 3x3 one runs on the implicit-GEMM kernel (bias in the epilogue); GroupNorm stays ATen."""
        from datr_b200 import conv as dconv, linear as dl
        conv, norm = proj[0], proj[1]
        if (dl.get_mode() == "tf32" and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
                and x.is_contiguous(memory_format=torch.channels_last)):
            n, cin, h, w = x.shape
            if conv.kernel_size == (1, 1) and conv.stride == (1, 1) and cin % 32 == 0:
                y = dl.linear(x.permute(0, 2, 3, 1).reshape(-1, cin), conv.weight.reshape(conv.out_channels, cin), conv.bias)
                return DINO._group_norm(norm, y.view(n, h, w, conv.out_channels).permute(0, 3, 1, 2))
            if dconv.use_kernel(x, conv):
                return DINO._group_norm(norm, dconv.conv3x3_bias_act(x, conv.weight, conv.bias, conv.stride[0], 0))
        return proj(x)

    @staticmethod
    def _group_norm(norm, y):
        """GroupNorm of an NHWC map on the NHWC kernel (datr_b200.groupnorm); ATen's CUDA GroupNorm is NCHW and hands back
        an NCHW-contiguous tensor, which then returns to NHWC once so that everything downstream works on views."""
        from datr_b200 import groupnorm as gn
        if isinstance(norm, nn.GroupNorm) and gn.applicable(norm, y) and _OWN_GROUPNORM:
            return gn.group_norm_nhwc(norm, y)
        return norm(y).contiguous(memory_format=torch.channels_last)

    def _project_levels(self, features, poss, full_mask):
        srcs, masks = [], []
        for l, feat in enumerate(features):
            src, mask = feat.decompose()
            assert mask is not None
            srcs.append(self._input_proj(self.input_proj[l], src))
            masks.append(mask)
        for l in range(len(srcs), self.num_feature_levels):
            src = self._input_proj(self.input_proj[l], features[-1].tensors if l == len(features) else srcs[-1])
            mask = F.interpolate(full_mask[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
            poss.append(self.backbone[1](NestedTensor(src, mask)).to(src.dtype))
            srcs.append(src)
            masks.append(mask)
        return srcs, masks, poss

    def _class_logits(self, hs, hs_enc):
        """Per-decoder-layer class logits [n_layers, N, nq, classes] and the logits of the selected encoder proposals.
        (Kept outside the captured head segment: the 91-wide library GEMMs of these layers do not survive CUDA-graph
        capture of their backward on this stack; they are 7 small launches.)"""
        from datr_b200 import linear as dl
        with dl.fp32_products():        # these logits feed the Hungarian cost matrices: fp32 products
            classes = torch.stack([head(h) for head, h in zip(self.class_embed, hs)])
            interm = self.transformer.enc_out_class_embed(hs_enc[-1]) if hs_enc is not None else None
        return classes, interm

    def _boxes(self, hs, reference):
        """Per-decoder-layer boxes: sigmoid(delta + logit(reference)), stacked over layers."""
        return torch.stack([(head(h) + inverse_sigmoid(ref)).sigmoid()
                            for ref, head, h in zip(reference[:-1], self.bbox_embed, hs)])

    def _heads(self, hs, reference):
        """Per-decoder-layer class logits and boxes, stacked over layers."""
        return self._class_logits(hs, None)[0], self._boxes(hs, reference)

    def _interm(self, out, interm_class, ref_enc, init_box_proposal, suffix=""):
        out["interm_outputs" + suffix] = {"pred_logits": interm_class, "pred_boxes": ref_enc[-1]}
        out["interm_outputs_for_matching_pre" + suffix] = {"pred_logits": interm_class, "pred_boxes": init_box_proposal}
        if ref_enc.shape[0] > 1:     # per-encoder-layer heads: unreachable with two_stage_type 'standard' (one entry)
            raise NotImplementedError("per-encoder-layer outputs (enc_outputs) are outside the DINO hot path")

    def _prototypes(self, feats, logits):
        if self.global_proto is None or self.global_proto.device != feats.device:
            self.global_proto = torch.zeros(self.num_classes, 256, device=feats.device)
            self.Amount = torch.zeros(self.num_classes, device=feats.device)
        proto, present, self.global_proto, self.Amount, _ = get_prototype_class_wise(
            feats, logits, self.num_classes, global_proto=self.global_proto.detach(), global_amount=self.Amount)
        return proto, present

    def _segment_owner(self, name, modules):
        """The sub-modules whose parameters a graph segment touches, as one (unregistered) container: a segment's
        graphed backward returns gradients for exactly these parameters."""
        cache = self.__dict__.setdefault("_segment_owners", {})
        if name not in cache:
            cache[name] = nn.ModuleList([m for m in modules if m is not None])
        return cache[name]

    def _image_discriminator(self, srcs_all):
        """Image-level domain discriminator behind a gradient reversal, on every level of both domains: [2B, S, 1]."""
        d_img = [self.D_img(grad_reverse(s)) for s in srcs_all]
        return torch.cat([d.flatten(2).transpose(1, 2) for d in d_img], dim=1)

    def _outputs_from(self, hs, reference, outputs_class, interm_class, ref_enc, init_box_proposal, dn_meta):
        """Decoder outputs + class logits -> the output dict (box heads, de-noising split, auxiliary and intermediate
        sets): pure device work."""
        outputs_coord = self._boxes(list(hs), reference)
        if self.dn_number > 0 and dn_meta is not None:
            dn_meta = dict(dn_meta)
            outputs_class, outputs_coord = dn_post_process(outputs_class, outputs_coord, dn_meta, self.aux_loss, self._set_aux_loss)
        out = {"pred_logits": outputs_class[-1], "pred_boxes": outputs_coord[-1]}
        if self.aux_loss:
            out["aux_outputs"] = self._set_aux_loss(outputs_class, outputs_coord)
        if interm_class is not None:
            self._interm(out, interm_class, ref_enc, init_box_proposal)
        out["dn_meta"] = dn_meta
        return out

    def forward(self, samples: NestedTensor, targets: List = None, self_training_flag=False):
        """samples: NestedTensor (tensors [B,3,H,W], mask [B,H,W] True on padding), a tensor or a list of images.
        In training mode the first half of the batch is the source domain (with `targets`), the second half the
        target domain.  Returns the reference's output dict (pred_logits, pred_boxes, aux_outputs, interm_outputs,
        interm_outputs_for_matching_pre, dn_meta, and in training da_output [+ *_target keys])."""
        if isinstance(samples, (list, torch.Tensor)):
            samples = nested_tensor_from_tensor_list(samples)

        if self.dn_number > 0 or targets is not None:
            dn_label, dn_bbox, attn_mask, dn_meta = prepare_for_cdn(
                dn_args=(targets, self.dn_number, self.dn_label_noise_ratio, self.dn_box_noise_scale),
                training=self.training, num_queries=self.num_queries, num_classes=self.num_classes,
                hidden_dim=self.hidden_dim, label_enc=self.label_enc)
        else:
            dn_bbox = dn_label = attn_mask = dn_meta = None

        srcs, masks, poss = self._features(samples)
        enc_t = joint_t = None
        if self.training:
            srcs, masks, poss, srcs_all, masks_all, poss_all, srcs_t, masks_t, poss_t = decompose_features(srcs, masks, poss)
            # The reference runs the transformer on the source half here (:291) and on the target half below (:380-382).
            # Its encoder treats every image independently, so both halves go through it in ONE call (DATR_JOINT_ENCODER=0
            # keeps two calls); the query selection and the decoder then run per half as in the reference.
            if _JOINT_ENCODER:
                half = srcs[0].shape[0]
                enc_all = self.transformer.encode(srcs_all, masks_all, poss_all)
                if _JOINT_DECODER and srcs_all[0].shape[0] == 2 * half:
                    # Query selection and decoder of both halves in ONE call as well.  The target half gets all-zero
                    # de-noising queries in the slots the source half fills: the de-noising attention mask
                    # (dn_components.py:105-121) hides those slots from the matching queries, and everything else in the
                    # decoder acts on one query at a time, so the matching queries of the target half come out exactly as
                    # from a separate pass without de-noising queries; the dummy slots are dropped below.  The decoder is
                    # bound by kernel count, not by arithmetic: one call costs little more than each of the two did.
                    dn_b = dn_l = None
                    if dn_bbox is not None:
                        dn_b = torch.cat([dn_bbox, torch.zeros_like(dn_bbox)], 0)
                        dn_l = torch.cat([dn_label, torch.zeros_like(dn_label)], 0)
                    hs_a, ref_a, hs_enc_a, ref_enc_a, ibp_a = self.transformer.decode(enc_all, dn_b, dn_l, attn_mask)
                    npad = dn_bbox.shape[1] if dn_bbox is not None else 0
                    hs, reference = [h[:half] for h in hs_a], [r[:half] for r in ref_a]
                    hs_enc = hs_enc_a[:, :half] if hs_enc_a is not None else None
                    ref_enc = ref_enc_a[:, :half] if ref_enc_a is not None else None
                    init_box_proposal = ibp_a[:half]
                    joint_t = ([h[half:, npad:] for h in hs_a], [r[half:, npad:] for r in ref_a],
                               hs_enc_a[:, half:] if hs_enc_a is not None else None,
                               ref_enc_a[:, half:] if ref_enc_a is not None else None, ibp_a[half:])
                    enc_t = enc_all
                else:
                    enc_s, enc_t = self.transformer.split_encoded(enc_all, [half, srcs_all[0].shape[0] - half])
                    hs, reference, hs_enc, ref_enc, init_box_proposal = self.transformer.decode(enc_s, dn_bbox, dn_label, attn_mask)
        if enc_t is None:
            hs, reference, hs_enc, ref_enc, init_box_proposal = self.transformer(srcs, masks, dn_bbox, poss, dn_label, attn_mask)
        # keeps label_enc in the autograd graph when there are no objects.  (Eager on purpose: label_enc already feeds
        # the eager de-noising query construction, and a parameter must not be shared between a live eager graph and a
        # segment being captured.)
        hs[0] = hs[0] + self.label_enc.weight[0, 0] * 0.0
        da = {}
        if self.training and graphs.ACTIVE is not None and srcs_all[0].is_cuda:
            # image-level discriminator (dense 3x3 convolutions that depend on the projected feature maps only) as a
            # segment on a second stream: forward beside the prediction heads / the matcher, and -- autograd runs a
            # node's backward on the stream of its forward, latest nodes first -- its backward beside the heads' and the
            # decoder's backward, whose short kernels leave most SMs idle.  Joined before forward() returns.
            da["backbone_DA"] = graphs.ACTIVE.call("d_img", self.D_img, self._image_discriminator, tuple(srcs_all), side=True)
        outputs_class, interm_class = self._class_logits(hs, hs_enc)
        if graphs.ACTIVE is not None and hs[0].is_cuda:
            out = graphs.ACTIVE.call("heads", self.bbox_embed, self._outputs_from, tuple(hs), tuple(reference),
                                     outputs_class, interm_class, ref_enc, init_box_proposal, dn_meta)
        else:
            out = self._outputs_from(tuple(hs), tuple(reference), outputs_class, interm_class, ref_enc, init_box_proposal, dn_meta)
        if not self.training:
            return out
        if targets is not None and torch.is_grad_enabled():
            # the criterion's Hungarian matchings only need the source-domain predictions: start them now so that the
            # host part overlaps the target-domain pass below (matcher.prefetch)
            prefetch(getattr(self, "_prefetch_matcher", None), out, targets)

        # ---- domain adaptation -----------------------------------------------------------------
        if "backbone_DA" not in da:
            da["backbone_DA"] = self._image_discriminator(tuple(srcs_all))

        pad = dn_meta["pad_size"] if dn_meta is not None else 0
        proto_s, present_s = self._prototypes(hs[-1][:, pad:, :], out["pred_logits"])

        if joint_t is not None:
            hs_t, reference_t, hs_enc_t, ref_enc_t, init_box_proposal_t = joint_t
        elif enc_t is not None:
            hs_t, reference_t, hs_enc_t, ref_enc_t, init_box_proposal_t = self.transformer.decode(enc_t, None, None, None)
        else:
            hs_t, reference_t, hs_enc_t, ref_enc_t, init_box_proposal_t = self.transformer(srcs_t, masks_t, None, poss_t, None, None)
        proto_t, present_t = self._prototypes(hs_t[-1], self.class_embed[-1](hs_t[-1]))

        da["proto_DA"] = {"da_protos": self.Proto_D(grad_reverse(torch.cat([proto_s, proto_t], dim=0))),
                          "class_map_source": present_s, "class_map_target": present_t}
        da["global_proto_DA"] = {"output_source": proto_s, "outputs_target": proto_t, "query_mask_source": present_s,
                                 "query_mask_target": present_t, "global_proto": self.global_proto}
        out["da_output"] = da

        if self_training_flag:          # expose the target-domain predictions for the pseudo-label losses
            hs_t[0] = hs_t[0] + self.label_enc.weight[0, 0] * 0.0
            class_t, coord_t = self._heads(hs_t, reference_t)
            out["pred_logits_target"], out["pred_boxes_target"] = class_t[-1], coord_t[-1]
            if self.aux_loss:
                out["aux_outputs_target"] = self._set_aux_loss(class_t, coord_t)
            if hs_enc_t is not None:
                self._interm(out, self.transformer.enc_out_class_embed(hs_enc_t[-1]), ref_enc_t, init_box_proposal_t,
                             suffix="_target")
        if graphs.ACTIVE is not None:
            graphs.ACTIVE.join_side()       # the discriminator's outputs are valid on the caller's stream from here on
        return out

    @torch.jit.unused
    def _set_aux_loss(self, outputs_class, outputs_coord):
        return [{"pred_logits": a, "pred_boxes": b} for a, b in zip(outputs_class[:-1], outputs_coord[:-1])]


class SetCriterion(nn.Module):
    """Hungarian matching + focal / L1 / GIoU losses for the final, auxiliary, intermediate and de-noising
    outputs, plus DATR's three domain-adaptation losses.  Loss names match the reference's weight_dict keys."""

    def __init__(self, num_classes, matcher, weight_dict, focal_alpha, losses):
        super().__init__()
        self.num_classes, self.matcher, self.weight_dict = num_classes, matcher, weight_dict
        self.losses, self.focal_alpha = losses, focal_alpha
        self.batched = os.environ.get("DATR_BATCHED_LOSSES", "1") != "0"    # False: one pass per prediction set, like the reference

    # ---- individual losses -------------------------------------------------------------------------
    def loss_labels(self, outputs, targets, indices, num_boxes, log=True):
        logits = outputs["pred_logits"]
        idx = self._get_src_permutation_idx(indices)
        matched = torch.cat([t["labels"][J] for t, (_, J) in zip(targets, indices)])
        onehot = torch.zeros_like(logits)
        onehot.index_put_((idx[0], idx[1], matched), onehot.new_ones(()))      # (a python scalar would be a host->device copy)
        loss_ce = sigmoid_focal_loss(logits, onehot, num_boxes, alpha=self.focal_alpha, gamma=2) * logits.shape[1]
        losses = {"loss_ce": loss_ce}
        if log:
            losses["class_error"] = 100 - accuracy(logits[idx], matched)[0]
        return losses

    @torch.no_grad()
    def loss_cardinality(self, outputs, targets, indices, num_boxes):
        logits = outputs["pred_logits"]
        counts = [len(v["labels"]) for v in targets]
        cached = getattr(self, "_n_tgt", None)          # one host->device copy per criterion call, not one per group
        if cached is not None and cached[0] == counts and cached[1].device == logits.device:
            n_tgt = cached[1]
        else:
            n_tgt = upload(counts, device=logits.device)
            self._n_tgt = (counts, n_tgt)
        n_pred = (logits.argmax(-1) != logits.shape[-1] - 1).sum(1)
        return {"cardinality_error": F.l1_loss(n_pred.float(), n_tgt.float())}

    def loss_boxes(self, outputs, targets, indices, num_boxes):
        idx = self._get_src_permutation_idx(indices)
        src = outputs["pred_boxes"][idx]
        tgt = torch.cat([t["boxes"][i] for t, (_, i) in zip(targets, indices)], dim=0)
        l1 = F.l1_loss(src, tgt, reduction="none")
        giou = box_ops.paired_giou(box_ops.box_cxcywh_to_xyxy(src), box_ops.box_cxcywh_to_xyxy(tgt))
        losses = {"loss_bbox": l1.sum() / num_boxes, "loss_giou": (1 - giou).sum() / num_boxes}
        with torch.no_grad():
            losses["loss_xy"] = l1[..., :2].sum() / num_boxes
            losses["loss_hw"] = l1[..., 2:].sum() / num_boxes
        return losses

    def _get_src_permutation_idx(self, indices):
        batch_idx = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        return batch_idx, torch.cat([src for (src, _) in indices])

    def _get_tgt_permutation_idx(self, indices):
        batch_idx = torch.cat([torch.full_like(tgt, i) for i, (_, tgt) in enumerate(indices)])
        return batch_idx, torch.cat([tgt for (_, tgt) in indices])

    def get_loss(self, loss, outputs, targets, indices, num_boxes, **kwargs):
        table = {"labels": self.loss_labels, "cardinality": self.loss_cardinality, "boxes": self.loss_boxes}
        assert loss in table, f"do you really want to compute {loss} loss?"
        return table[loss](outputs, targets, indices, num_boxes, **kwargs)

    # ---- domain-adaptation losses ------------------------------------------------------------------
    def loss_da(self, outputs):
        """Image-level discriminator logits [2B', S, 1]: source half -> 0, target half -> 1."""
        B = outputs.shape[0]
        assert B % 2 == 0
        src, tgt = outputs[:B // 2], outputs[B // 2:]
        return F.binary_cross_entropy_with_logits(src, torch.zeros_like(src)) \
            + F.binary_cross_entropy_with_logits(tgt, torch.ones_like(tgt))

    def loss_proto_da(self, outputs):
        protos = outputs["da_protos"]
        assert protos.shape[0] % 2 == 0
        k = outputs["class_map_source"].shape[0]
        target = torch.cat([torch.zeros_like(protos[:k]), torch.ones_like(protos[k:])], 0)
        loss = F.binary_cross_entropy_with_logits(protos, target, reduction="none")
        present = torch.cat([outputs["class_map_source"], outputs["class_map_target"]], dim=0).unsqueeze(1)
        return (loss * present).mean()

    def loss_contrast_da(self, outputs):
        proto = outputs["global_proto"]
        mask_s, mask_t = outputs["query_mask_source"], outputs["query_mask_target"]
        assert not proto.requires_grad and not mask_s.requires_grad and not mask_t.requires_grad
        q_s, q_t = outputs["output_source"], outputs["outputs_target"]
        assert q_s.requires_grad and q_t.requires_grad
        k = q_s.shape[0]
        proto = F.normalize(proto, dim=1).t().contiguous()
        eye = torch.eye(k, device=proto.device)
        return F.cross_entropy(F.normalize(q_s, dim=1) @ proto, eye * mask_s) \
            + F.cross_entropy(F.normalize(q_t, dim=1) @ proto, eye * mask_t)

    # ---- batched loss groups ------------------------------------------------------------
    # The step scores 13 prediction sets of identical shape against the same targets (final + 5 auxiliary decoder
    # layers + intermediate, and the 6 de-noising sets).  The reference runs loss_labels / loss_boxes /
    # loss_cardinality once per set (dino.py:723-933), ~40 tiny kernels each and as many again in the backward; here a
    # FAMILY of sets is stacked along a leading group axis and scored in one pass, per-group sums taken at the end.
    def _family_helpers(self, fam, counts, lens, groups, device):
        """Constant index vectors of a family, cached by signature: for every matched pair (ordered group-major, then
        image) the flattened (group, image) row and the image's offset into the concatenated targets."""
        key = (fam, tuple(counts), tuple(lens), groups, str(device))
        cache = self.__dict__.setdefault("_family_cache", {})
        if key not in cache:
            import numpy as np
            bs = len(counts)
            offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
            gb = np.concatenate([np.full(lens[i], g * bs + i, dtype=np.int64) for g in range(groups) for i in range(bs)])
            toff = np.concatenate([np.full(lens[i], offs[i], dtype=np.int64) for g in range(groups) for i in range(bs)])
            if len(cache) > 64:
                cache.clear()
            cache[key] = (upload(gb, device=device), upload(toff, device=device))
        return cache[key]

    def _batched_helpers(self, outputs, targets, pre):
        """Everything the batched loss path needs that is a function of the target counts only (host-known), built
        outside the captured segment: index helpers of both families and the per-image target counts."""
        if list(self.losses) != ["labels", "boxes", "cardinality"] or "enc_outputs" in outputs:
            return None
        sets = [outputs] + list(outputs.get("aux_outputs", [])) + ([outputs["interm_outputs"]] if "interm_outputs" in outputs else [])
        shape = outputs["pred_logits"].shape
        if len(sets) != len(pre) or any(o["pred_logits"].shape != shape for o in sets):
            return None
        device = outputs["pred_logits"].device
        counts = [len(t["labels"]) for t in targets]
        lens = [int(src.numel()) for src, _ in pre[0]]
        if sum(lens) == 0 or any([int(src.numel()) for src, _ in ind] != lens for ind in pre):
            return None
        helpers = {"m": self._family_helpers("m", counts, lens, len(sets), device)}
        cached = getattr(self, "_n_tgt", None)
        if cached is None or cached[0] != counts or cached[1].device != device:
            self._n_tgt = cached = (counts, upload(counts, device=device))
        helpers["n_tgt"] = cached[1]
        dn_meta = outputs.get("dn_meta")
        if self.training and dn_meta and "output_known_lbs_bboxes" in dn_meta:
            known = dn_meta["output_known_lbs_bboxes"]
            scalar, pad = dn_meta["num_dn_group"], dn_meta["pad_size"]
            dsets = [known] + list(known.get("aux_outputs", []))
            if any(o["pred_logits"].shape != known["pred_logits"].shape for o in dsets):
                return None
            key = ("dn_idx", tuple(counts), scalar, pad, len(dsets), str(device))
            cache = self.__dict__.setdefault("_family_cache", {})
            if key not in cache:
                pos = self._dn_indices(targets, pad // scalar, scalar, device)
                cache[key] = (torch.cat([o for o, _ in pos]).repeat(len(dsets)), torch.cat([t for _, t in pos]).repeat(len(dsets)))
            dlens = [scalar * n for n in counts]
            helpers["d"] = self._family_helpers("d", counts, dlens, len(dsets), device) + cache[key]
        return helpers

    def _family_losses(self, sets, targets, src, tgt, gb, toff, n_tgt, num_boxes, suffixes, log_group=None):
        """loss_labels + loss_boxes + loss_cardinality of `sets` (same shapes) in one pass.  src / tgt: query and
        target index of every matched pair, group-major; gb / toff: see _family_helpers."""
        G = len(sets)
        logits = torch.stack([o["pred_logits"] for o in sets]).flatten(0, 1)        # [G*bs, nq, nc]
        boxes = torch.stack([o["pred_boxes"] for o in sets]).flatten(0, 1)          # [G*bs, nq, 4]
        labels_all = torch.cat([t["labels"] for t in targets])
        boxes_all = torch.cat([t["boxes"] for t in targets])
        tsel = tgt + toff
        matched = labels_all[tsel]
        onehot = torch.zeros_like(logits)
        onehot.index_put_((gb, src, matched), onehot.new_ones(()))
        p = logits.sigmoid()
        ce = F.binary_cross_entropy_with_logits(logits, onehot, reduction="none")
        p_t = p * onehot + (1 - p) * (1 - onehot)
        focal = ce * (1 - p_t) ** 2
        if self.focal_alpha >= 0:
            focal = (self.focal_alpha * onehot + (1 - self.focal_alpha) * (1 - onehot)) * focal
        # sigmoid_focal_loss(...) * num_queries of the reference (= sum over the set / num_boxes)
        loss_ce = focal.view(G, -1).sum(1) / num_boxes
        sb = boxes[gb, src]
        tb = boxes_all[tsel]
        l1 = F.l1_loss(sb, tb, reduction="none")
        giou = box_ops.paired_giou(box_ops.box_cxcywh_to_xyxy(sb), box_ops.box_cxcywh_to_xyxy(tb))
        K = src.numel() // G
        loss_bbox = l1.view(G, -1).sum(1) / num_boxes
        loss_giou = (1 - giou).view(G, K).sum(1) / num_boxes
        with torch.no_grad():
            l1d = l1.detach().view(G, K, 4)
            loss_xy = l1d[..., :2].sum((1, 2)) / num_boxes
            loss_hw = l1d[..., 2:].sum((1, 2)) / num_boxes
            n_pred = (logits.argmax(-1) != logits.shape[-1] - 1).sum(1).view(G, -1)
            card = (n_pred.float() - n_tgt.float()[None]).abs().mean(1)
        out = {}
        for g, sfx in enumerate(suffixes):
            out["loss_ce" + sfx] = loss_ce[g]
            out["loss_bbox" + sfx], out["loss_giou" + sfx] = loss_bbox[g], loss_giou[g]
            out["loss_xy" + sfx], out["loss_hw" + sfx] = loss_xy[g], loss_hw[g]
            out["cardinality_error" + sfx] = card[g]
        if log_group is not None:
            with torch.no_grad():
                sel = slice(log_group * K, (log_group + 1) * K)
                out["class_error"] = 100 - accuracy(logits[gb[sel], src[sel]], matched[sel])[0]
        return out

    def _losses_batched(self, outputs, targets, pre, num_boxes, helpers):
        aux = list(outputs.get("aux_outputs", []))
        sets = [outputs] + aux + ([outputs["interm_outputs"]] if "interm_outputs" in outputs else [])
        suffixes = [""] + [f"_{i}" for i in range(len(aux))] + (["_interm"] if "interm_outputs" in outputs else [])
        src = torch.cat([s_ for ind in pre for s_, _ in ind])
        tgt = torch.cat([t_ for ind in pre for _, t_ in ind])
        losses = {}
        dn_zero = ("loss_bbox_dn", "loss_giou_dn", "loss_ce_dn", "loss_xy_dn", "loss_hw_dn", "cardinality_error_dn")
        if "d" in helpers:
            dn_meta = outputs["dn_meta"]
            known = dn_meta["output_known_lbs_bboxes"]
            dsets = [known] + list(known.get("aux_outputs", []))
            dsfx = ["_dn"] + [f"_dn_{i}" for i in range(len(dsets) - 1)]
            gb, toff, dsrc, dtgt = helpers["d"]
            losses.update(self._family_losses(dsets, targets, dsrc, dtgt, gb, toff, helpers["n_tgt"],
                                              num_boxes * dn_meta["num_dn_group"], dsfx))
        else:
            device = outputs["pred_logits"].device
            losses.update({k: torch.zeros((), device=device) for k in dn_zero})
            losses.update({f"{k}_{i}": torch.zeros((), device=device) for k in dn_zero for i in range(len(aux))})
        gb, toff = helpers["m"]
        losses.update(self._family_losses(sets, targets, src, tgt, gb, toff, helpers["n_tgt"], num_boxes, suffixes, log_group=0))
        return losses

    # ---- driver ------------------------------------------------------------------------------------
    def _dn_indices(self, targets, single_pad, scalar, device):
        pos = []
        for t in targets:
            n = len(t["labels"])
            if n > 0:
                tgt_idx = torch.arange(n, device=device).repeat(scalar)
                out_idx = (torch.arange(scalar, device=device)[:, None] * single_pad + torch.arange(n, device=device)[None]).flatten()
            else:
                out_idx = tgt_idx = torch.zeros(0, dtype=torch.long, device=device)
            pos.append((out_idx, tgt_idx))
        return pos

    def _group(self, outputs, targets, indices, num_boxes, suffix, log_labels=False):
        out = {}
        for loss in self.losses:
            kwargs = {"log": log_labels} if loss == "labels" else {}
            out.update({k + suffix: v for k, v in self.get_loss(loss, outputs, targets, indices, num_boxes, **kwargs).items()})
        return out

    def _losses_from(self, outputs, targets, pre, num_boxes, target_domain_flag, training, return_indices, helpers=None):
        """All losses given the matchings `pre` (device index tensors): pure device work, no host synchronisation.
        With `helpers` (see _batched_helpers) the 13 set losses run as two batched families."""
        if helpers is not None:
            losses = self._losses_batched(outputs, targets, pre, num_boxes, helpers)
            if "da_output" in outputs:
                da = outputs["da_output"]
                losses["loss_backbone_DA"] = self.loss_da(da["backbone_DA"])
                losses["loss_proto_DA"] = self.loss_proto_da(da["proto_DA"])
                losses["loss_global_proto_DA"] = self.loss_contrast_da(da["global_proto_DA"])
            return losses
        device = outputs["pred_logits"].device
        key_aux = "aux_outputs_target" if target_domain_flag else "aux_outputs"
        key_interm = "interm_outputs_target" if target_domain_flag else "interm_outputs"
        indices = indices0 = pre[0]
        indices_list = []
        losses = {}
        dn_zero = ("loss_bbox_dn", "loss_giou_dn", "loss_ce_dn", "loss_xy_dn", "loss_hw_dn", "cardinality_error_dn")
        use_dn = False
        if not target_domain_flag:
            dn_meta = outputs["dn_meta"]
            use_dn = bool(training and dn_meta and "output_known_lbs_bboxes" in dn_meta)
            if use_dn:
                known = dn_meta["output_known_lbs_bboxes"]
                scalar, pad_size = dn_meta["num_dn_group"], dn_meta["pad_size"]
                assert pad_size % scalar == 0
                single_pad = pad_size // scalar
                dn_pos_idx = self._dn_indices(targets, single_pad, scalar, device)
                losses.update(self._group(known, targets, dn_pos_idx, num_boxes * scalar, "_dn"))
            else:
                losses.update({k: torch.zeros((), device=device) for k in dn_zero})
            losses.update(self._group(outputs, targets, indices, num_boxes, "", log_labels=True))

        if key_aux in outputs:
            for i, aux in enumerate(outputs[key_aux]):
                indices = pre[1 + i]
                if return_indices:
                    indices_list.append(indices)
                losses.update(self._group(aux, targets, indices, num_boxes, f"_{i}"))
                if not target_domain_flag:
                    if use_dn:
                        losses.update(self._group(known["aux_outputs"][i], targets, dn_pos_idx, num_boxes * scalar, f"_dn_{i}"))
                    else:
                        losses.update({f"{k}_{i}": torch.zeros((), device=device) for k in dn_zero})

        if key_interm in outputs:
            interm = outputs[key_interm]
            indices = pre[1 + len(outputs.get(key_aux, []))]
            if return_indices:
                indices_list.append(indices)
            losses.update(self._group(interm, targets, indices, num_boxes, "_interm"))

        key_enc = "enc_outputs_target" if target_domain_flag else "enc_outputs"
        if key_enc in outputs:
            for i, enc in enumerate(outputs[key_enc]):
                indices = self.matcher(enc, targets)
                if return_indices:
                    indices_list.append(indices)
                losses.update(self._group(enc, targets, indices, num_boxes, f"_enc_{i}"))

        if "da_output" in outputs:
            da = outputs["da_output"]
            losses["loss_backbone_DA"] = self.loss_da(da["backbone_DA"])
            losses["loss_proto_DA"] = self.loss_proto_da(da["proto_DA"])
            losses["loss_global_proto_DA"] = self.loss_contrast_da(da["global_proto_DA"])

        if return_indices:
            indices_list.append(indices0)
            return losses, indices_list
        return losses

    def _weighted_losses_from(self, outputs, targets, pre, num_boxes, target_domain_flag, training, helpers):
        """(loss names, their values packed into one detached tensor, sum_k weight_dict[k] * loss_k)."""
        losses = self._losses_from(outputs, targets, pre, num_boxes, target_domain_flag, training, False, helpers)
        keys = tuple(losses)
        vals = torch.stack([losses[k].reshape(()) for k in keys])
        cache = self.__dict__.setdefault("_weight_vectors", {})
        ck = (keys, str(vals.device))
        if ck not in cache:       # first built during the (uncaptured) warm-up passes of the segment
            cache[ck] = torch.tensor([float(self.weight_dict.get(k, 0.0)) for k in keys], dtype=vals.dtype, device=vals.device)
        return keys, vals.detach(), torch.dot(vals, cache[ck])

    def forward(self, outputs, targets, return_indices=False, target_domain_flag=False):
        """outputs: the model's dict; targets: list of {'labels','boxes'} per (source or pseudo-labelled) image.
        With target_domain_flag the *_target keys are scored instead (self-training)."""
        if target_domain_flag:
            outputs_without_aux = {k.replace("_target", ""): v for k, v in outputs.items() if k != "aux_outputs_target"}
            outputs.update({"pred_boxes": outputs.pop("pred_boxes_target")})
            outputs.update({"pred_logits": outputs.pop("pred_logits_target")})
            device = outputs["pred_logits"].device
        else:
            outputs_without_aux = {k: v for k, v in outputs.items() if k != "aux_outputs"}
            device = next(iter(outputs.values())).device

        key_aux = "aux_outputs_target" if target_domain_flag else "aux_outputs"
        key_interm = "interm_outputs_target" if target_domain_flag else "interm_outputs"
        pre = nb = None
        if len(targets) > 0:
            # all matchings of the step (final, auxiliary decoder layers, intermediate) in one batched pass; the
            # num_boxes all-reduce (reference :767-770) rides along with it
            sets = [outputs_without_aux] + list(outputs.get(key_aux, [])) + ([outputs[key_interm]] if key_interm in outputs else [])
            handle = None if target_domain_flag else take_prefetched(outputs, sets, targets)
            if handle is None and batchable(self.matcher, sets):
                handle = BatchedMatch(self.matcher, sets, targets)
            if handle is not None:
                pre, nb = handle.result()
            else:
                pre = match_many(self.matcher, sets, targets)
            indices = pre[0]
            indices0, indices_list = indices, []
        else:       # no pseudo labels on this rank: still take part in the collective below
            indices = None
        if nb is not None:
            num_boxes = nb      # counted (and all-reduced) by BatchedMatch without a host->device copy or a sync here
        else:
            n_local = sum(len(t["labels"]) for t in targets) if indices is not None else 1
            num_boxes = upload([n_local], dtype=torch.float, device=outputs["pred_logits"].device)
            if is_dist_avail_and_initialized():
                torch.distributed.all_reduce(num_boxes)
            if indices is None:
                num_boxes = num_boxes - 1
            num_boxes = torch.clamp(num_boxes / get_world_size(), min=1).item()
        if indices is None:
            return {}

        helpers = None
        if self.batched and not return_indices and not target_domain_flag and isinstance(pre, list):
            helpers = self._batched_helpers(outputs, targets, pre)
        if (graphs.ACTIVE is not None and device.type == "cuda" and torch.is_grad_enabled() and not return_indices
                and not target_domain_flag):
            # every loss of the step as one captured segment: the matched indices, targets and predictions are its
            # tensor inputs, everything host-side (matching, num_boxes, index helpers) happened above
            if getattr(self, "fold_weighted_sum", False):
                # engine.py:99's weighted sum inside the segment: ONE differentiable output (the total) and one packed,
                # detached tensor of the individual losses, instead of ~40 scalar outputs whose stack / gradient copies
                # would be launched one by one between the criterion's forward and backward graphs
                keys, packed, total = graphs.ACTIVE.call("criterion", None, self._weighted_losses_from, outputs, targets, pre,
                                                         num_boxes, target_domain_flag, self.training, helpers)
                losses = {k: packed[i] for i, k in enumerate(keys)}
                losses["_weighted_total"] = total
                return losses
            return graphs.ACTIVE.call("criterion", None, self._losses_from, outputs, targets, pre, num_boxes,
                                      target_domain_flag, self.training, False, helpers)
        return self._losses_from(outputs, targets, pre, num_boxes, target_domain_flag, self.training, return_indices, helpers)

    def prep_for_dn(self, dn_meta):
        groups, pad = dn_meta["num_dn_group"], dn_meta["pad_size"]
        assert pad % groups == 0
        return dn_meta["output_known_lbs_bboxes"], pad // groups, groups


class PostProcess(nn.Module):
    """Top-k over (query, class) scores -> per-image {'scores','labels','boxes'} in absolute xyxy pixels."""

    def __init__(self, num_select=100, nms_iou_threshold=-1) -> None:
        super().__init__()
        self.num_select, self.nms_iou_threshold = num_select, nms_iou_threshold

    @torch.no_grad()
    def forward(self, outputs, target_sizes, not_to_xyxy=False, test=False):
        logits, bbox = outputs["pred_logits"], outputs["pred_boxes"]
        assert len(logits) == len(target_sizes) and target_sizes.shape[1] == 2
        n_cls = logits.shape[2]
        scores, flat = torch.topk(logits.sigmoid().view(logits.shape[0], -1), self.num_select, dim=1)
        query = torch.div(flat, n_cls, rounding_mode="floor")
        labels = flat % n_cls
        boxes = bbox if not_to_xyxy else box_ops.box_cxcywh_to_xyxy(bbox)
        if test:
            assert not not_to_xyxy
            boxes[:, :, 2:] = boxes[:, :, 2:] - boxes[:, :, :2]
        boxes = torch.gather(boxes, 1, query.unsqueeze(-1).repeat(1, 1, 4))
        img_h, img_w = target_sizes.unbind(1)
        boxes = boxes * torch.stack([img_w, img_h, img_w, img_h], dim=1)[:, None, :]
        if self.nms_iou_threshold > 0:
            from torchvision.ops.boxes import nms
            keep = [nms(b, s, iou_threshold=self.nms_iou_threshold) for b, s in zip(boxes, scores)]
            return [{"scores": s[i], "labels": l[i], "boxes": b[i]} for s, l, b, i in zip(scores, labels, boxes, keep)]
        return [{"scores": s, "labels": l, "boxes": b} for s, l, b in zip(scores, labels, boxes)]


@MODULE_BUILD_FUNCS.registe_with_name(module_name="dino")
def build_dino(args):
    """args -> (model, criterion, {'bbox': PostProcess}); reads the same fields as the reference (:1018-1136)."""
    num_classes = args.num_classes
    device = torch.device(args.device)
    if getattr(args, "masks", False):
        raise NotImplementedError("segmentation heads are outside the DINO hot path (masks=False in all DATR configs)")
    backbone = build_backbone(args)
    transformer = build_deformable_transformer(args)
    model = DINO(
        backbone, transformer, num_classes=num_classes, num_queries=args.num_queries, aux_loss=True, iter_update=True,
        query_dim=4, random_refpoints_xy=args.random_refpoints_xy, fix_refpoints_hw=args.fix_refpoints_hw,
        num_feature_levels=args.num_feature_levels, nheads=args.nheads,
        dec_pred_class_embed_share=getattr(args, "dec_pred_class_embed_share", True),
        dec_pred_bbox_embed_share=getattr(args, "dec_pred_bbox_embed_share", True),
        two_stage_type=args.two_stage_type, two_stage_bbox_embed_share=args.two_stage_bbox_embed_share,
        two_stage_class_embed_share=args.two_stage_class_embed_share, decoder_sa_type=args.decoder_sa_type,
        num_patterns=args.num_patterns, dn_number=args.dn_number if args.use_dn else 0,
        dn_box_noise_scale=args.dn_box_noise_scale, dn_label_noise_ratio=args.dn_label_noise_ratio,
        dn_labelbook_size=getattr(args, "dn_labelbook_size", num_classes))
    matcher = build_matcher(args)
    # lets DINO.forward start the criterion's matchings early (matcher.prefetch); a plain attribute, not a sub-module:
    # module list and state_dict stay the reference's
    object.__setattr__(model, "_prefetch_matcher", matcher)

    weight_dict = {"loss_ce": args.cls_loss_coef, "loss_bbox": args.bbox_loss_coef, "loss_giou": args.giou_loss_coef}
    plain = copy.deepcopy(weight_dict)
    weight_dict["loss_backbone_DA"] = args.da_backbone_loss_coef
    weight_dict["loss_proto_DA"] = args.da_proto_loss_coef
    weight_dict["loss_global_proto_DA"] = args.da_global_proto_coef
    weight_dict["loss_self_training"] = args.self_training_loss_coef
    if args.use_dn:
        weight_dict.update({"loss_ce_dn": args.cls_loss_coef, "loss_bbox_dn": args.bbox_loss_coef,
                            "loss_giou_dn": args.giou_loss_coef})
    per_layer = copy.deepcopy(weight_dict)
    if args.aux_loss:
        for i in range(args.dec_layers - 1):
            weight_dict.update({f"{k}_{i}": v for k, v in per_layer.items()})
    if args.two_stage_type != "no":
        no_box = getattr(args, "no_interm_box_loss", False)
        coef = getattr(args, "interm_loss_coef", 1.0)
        scale = {"loss_ce": 1.0, "loss_bbox": 0.0 if no_box else 1.0, "loss_giou": 0.0 if no_box else 1.0}
        weight_dict.update({k + "_interm": v * coef * scale[k] for k, v in plain.items()})

    criterion = SetCriterion(num_classes, matcher=matcher, weight_dict=weight_dict, focal_alpha=args.focal_alpha,
                             losses=["labels", "boxes", "cardinality"])
    criterion.to(device)
    postprocessors = {"bbox": PostProcess(num_select=args.num_select, nms_iou_threshold=args.nms_iou_threshold)}
    return model, criterion, postprocessors
