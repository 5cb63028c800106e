"""Deformable transformer of DINO: multi-scale deformable encoder, two-stage query selection, decoder.

Mirrors the reference's models/dino/deformable_transformer.py for the configuration DINO uses
(deformable encoder + decoder, two_stage_type 'standard' or 'no', decoder_sa_type 'sa', post-norm):
DeformableTransformer (:25-431), TransformerEncoder (:434-577), TransformerDecoder (:579-763),
DeformableTransformerEncoderLayer (:765-820), DeformableTransformerDecoderLayer (:822-994),
build_deformable_transformer (:1004-1066).  Module / parameter names are the reference's, so its
checkpoints load with strict=True.

Differences that do not change results:
  * tensors stay batch-first ([N, tokens, C]) end to end; the reference flips the decoder to
    sequence-first and back around every MSDeformAttn call (:392-399, :950-952);
  * level geometry is carried as Python ints next to the int64 device tensor the op needs, so building
    reference points / proposals never synchronises with the device (the reference iterates over a
    CUDA tensor, :479-484);
  * in the two-stage block the box head runs on the 900 selected tokens only (it is row-wise, so
    gathering first gives the same numbers as :340-345 for a third of the FLOPs);
  * decoder self-attention runs on the packed in_proj weights of nn.MultiheadAttention (same parameter names and
    math): batched GEMMs around the masked-softmax kernel of datr_b200.attention on CUDA fp32,
    F.scaled_dot_product_attention otherwise (DATR_OWN_ATTENTION=0 forces the latter).
Options the DINO/DATR configs never enable (box attention, layer sharing, dec_layer_number, patterns,
'ca_label' / 'ca_content' self-attention, key-aware cross-attention) raise NotImplementedError.
"""
import copy
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

import os

from datr_b200.util.misc import inverse_sigmoid
from datr_b200 import attention, graphs
from datr_b200 import linear as dl
from datr_b200.layernorm import layer_norm as ln
from .ops.modules import MSDeformAttn
from .utils import MLP, _get_activation_fn, gen_encoder_output_proposals, gen_sineembed_for_position, level_sizes


_OWN_ATTENTION = os.environ.get("DATR_OWN_ATTENTION", "1") != "0"
_FUSED_ATTENTION = os.environ.get("DATR_FUSED_ATTENTION", "1") != "0"   # 0: batched GEMMs around the softmax kernel
_MEMORY_GRAD_CHAIN = os.environ.get("DATR_MEMORY_GRAD_CHAIN", "1") != "0"  # 0: autograd sums the decoder's memory gradients


def _fused_attention_on():
    """The tensor-core attention kernels multiply in TF32: they belong to the 'tf32' mode of datr_b200.linear (the
    benchmarked mode); the strict-fp32 mode keeps fp32 products (batched GEMMs around the softmax kernel)."""
    return _OWN_ATTENTION and _FUSED_ATTENTION and dl.get_mode() == "tf32"


def _get_clones(module, N, layer_share=False):
    if layer_share:
        return nn.ModuleList([module for _ in range(N)])
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class PackedSelfAttention(nn.Module):
    """Multi-head attention with nn.MultiheadAttention's parameter layout (in_proj_weight [3C,C],
    in_proj_bias [3C], out_proj.{weight,bias}); query and key share one input, value has its own."""

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, qk_in, v_in, attn_mask=None, mask_bits=None, residual=None):
        """qk_in, v_in: [N, T, C]; attn_mask [T, T] bool with True = blocked (nn.MultiheadAttention's
        convention) or an additive float mask; mask_bits (optional): attention.pack_mask(attn_mask, T), shared by the
        layers of a decoder pass.  Returns [N, T, C]."""
        N, T, C = qk_in.shape
        H = self.num_heads
        qk = dl.linear(qk_in, self.in_proj_weight[:2 * C], self.in_proj_bias[:2 * C])
        v = dl.linear(v_in, self.in_proj_weight[2 * C:], self.in_proj_bias[2 * C:])
        drop = self.dropout if self.training else 0.0
        if _fused_attention_on() and attention.fused_applicable(qk, v, H, attn_mask, drop):
            # one tcgen05 kernel on the packed projections: no head-split copies, no score matrix in HBM
            o = attention.fused_self_attention(qk, v, H, attn_mask, bits=mask_bits)
            return dl.linear(o, self.out_proj.weight, self.out_proj.bias, residual=residual)
        q, k = qk.view(N, T, 2, H, C // H).permute(2, 0, 3, 1, 4)
        v = v.view(N, T, H, C // H).transpose(1, 2)
        if _OWN_ATTENTION and attention.applicable(q, attn_mask, drop):
            # score matrix in HBM: two batched GEMMs around the in-place masked-softmax kernel (datr_b200.attention)
            o = attention.self_attention(q, k, v, attn_mask)
        else:
            if attn_mask is not None and attn_mask.dtype == torch.bool:
                attn_mask = ~attn_mask                      # SDPA: True = may attend
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask, dropout_p=drop)
        return dl.linear(o.transpose(1, 2).reshape(N, T, C), self.out_proj.weight, self.out_proj.bias, residual=residual)


def _fusable(layer, *dropouts):
    """Bias / ReLU / residual epilogue fusion is exact only when the dropouts in between are inactive (p = 0, the
    DINO configuration, or eval mode) and the activation is ReLU."""
    act = getattr(layer, "activation", F.relu)
    return act is F.relu and all(d is None or d.p == 0.0 or not layer.training for d in dropouts)


class DeformableTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 add_channel_attention=False, use_deformable_box_attn=False, box_attn_type="roi_align"):
        super().__init__()
        if use_deformable_box_attn or add_channel_attention:
            raise NotImplementedError("box attention / channel attention are outside the DINO hot path")
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation, d_model=d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, src):
        if _fusable(self, self.dropout2, self.dropout3):
            # linear1 + bias + ReLU and linear2 + bias + residual are one kernel each (datr_b200.linear)
            return ln(self.norm2, dl.ffn(src, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias))
        return ln(self.norm2, src + self.dropout3(self.linear2(self.dropout2(self.activation(self.linear1(src))))))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask=None):
        if _fusable(self, self.dropout1):
            return self.forward_ffn(ln(self.norm1, self.self_attn(self.with_pos_embed(src, pos), reference_points, src,
                                                              spatial_shapes, level_start_index, key_padding_mask,
                                                              residual=src)))
        attn = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes,
                              level_start_index, key_padding_mask)
        return self.forward_ffn(ln(self.norm1, src + self.dropout1(attn)))


class TransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, norm=None, d_model=256, num_queries=300, deformable_encoder=False,
                 enc_layer_share=False, enc_layer_dropout_prob=None, two_stage_type="no"):
        super().__init__()
        if enc_layer_dropout_prob is not None or two_stage_type not in ("no", "standard"):
            raise NotImplementedError("encoder layer dropout / per-layer two-stage are outside the DINO hot path")
        self.layers = _get_clones(encoder_layer, num_layers, layer_share=enc_layer_share) if num_layers > 0 else []
        self.query_scale = None
        self.num_queries, self.num_layers, self.norm, self.d_model = num_queries, num_layers, norm, d_model
        self.deformable_encoder = deformable_encoder
        self.two_stage_type = two_stage_type

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """Token centres in valid-area units, broadcast to every level: [N, S, L, 2] (x, y)."""
        pts = []
        for lvl, (H, W) in enumerate(level_sizes(spatial_shapes)):
            ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, dtype=torch.float32, device=device),
                                    torch.linspace(0.5, W - 0.5, W, dtype=torch.float32, device=device), indexing="ij")
            y = ys.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
            x = xs.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
            pts.append(torch.stack((x, y), -1))
        return torch.cat(pts, 1)[:, :, None] * valid_ratios[:, None]

    def forward(self, src: Tensor, pos: Tensor, spatial_shapes: Tensor, level_start_index: Tensor,
                valid_ratios: Tensor, key_padding_mask: Tensor, ref_token_index: Optional[Tensor] = None,
                ref_token_coord: Optional[Tensor] = None, shapes_list=None):
        """src, pos [N,S,C]; returns (memory [N,S,C], None, None) like the reference for 'no'/'standard'."""
        assert ref_token_index is None
        out = src
        if self.num_layers > 0:
            ref = self.get_reference_points(shapes_list if shapes_list is not None else spatial_shapes,
                                            valid_ratios, device=src.device)
        for layer in self.layers:
            out = layer(src=out, pos=pos, reference_points=ref, spatial_shapes=spatial_shapes,
                        level_start_index=level_start_index, key_padding_mask=key_padding_mask)
        if self.norm is not None:
            out = ln(self.norm, out)
        return out, None, None


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 use_deformable_box_attn=False, box_attn_type="roi_align", key_aware_type=None,
                 decoder_sa_type="ca", module_seq=("sa", "ca", "ffn")):
        super().__init__()
        if use_deformable_box_attn or key_aware_type is not None or decoder_sa_type != "sa":
            raise NotImplementedError("only decoder_sa_type='sa' with plain MSDeformAttn is on the DINO hot path")
        self.module_seq = list(module_seq)
        assert sorted(self.module_seq) == ["ca", "ffn", "sa"]
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = PackedSelfAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation, d_model=d_ffn, batch_dim=1)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.key_aware_type, self.key_aware_proj, self.decoder_sa_type = key_aware_type, None, decoder_sa_type

    def rm_self_attn_modules(self):
        self.self_attn = self.dropout2 = self.norm2 = None

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, tgt):
        if _fusable(self, self.dropout3, self.dropout4):
            return ln(self.norm3, dl.ffn(tgt, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias))
        return ln(self.norm3, tgt + self.dropout4(self.linear2(self.dropout3(self.activation(self.linear1(tgt))))))

    def forward_sa(self, tgt, tgt_query_pos=None, self_attn_mask=None, self_attn_mask_bits=None):
        if self.self_attn is None:
            return tgt
        qk = self.with_pos_embed(tgt, tgt_query_pos)
        if _fusable(self, self.dropout2):     # residual add in the output projection's epilogue
            return ln(self.norm2, self.self_attn(qk, tgt, attn_mask=self_attn_mask, mask_bits=self_attn_mask_bits,
                                                 residual=tgt))
        return ln(self.norm2, tgt + self.dropout2(self.self_attn(qk, tgt, attn_mask=self_attn_mask,
                                                                  mask_bits=self_attn_mask_bits)))

    def forward_ca(self, tgt, tgt_query_pos, tgt_reference_points, memory, memory_key_padding_mask,
                   memory_level_start_index, memory_spatial_shapes, memory_grad_chain=None):
        if _fusable(self, self.dropout1):
            return ln(self.norm1, self.cross_attn(self.with_pos_embed(tgt, tgt_query_pos), tgt_reference_points, memory,
                                              memory_spatial_shapes, memory_level_start_index,
                                              memory_key_padding_mask, residual=tgt, value_grad_chain=memory_grad_chain))
        attn = self.cross_attn(self.with_pos_embed(tgt, tgt_query_pos), tgt_reference_points, memory,
                               memory_spatial_shapes, memory_level_start_index, memory_key_padding_mask,
                               value_grad_chain=memory_grad_chain)
        return ln(self.norm1, tgt + self.dropout1(attn))

    def forward(self, tgt, tgt_query_pos=None, tgt_query_sine_embed=None, tgt_key_padding_mask=None,
                tgt_reference_points=None, memory=None, memory_key_padding_mask=None, memory_level_start_index=None,
                memory_spatial_shapes=None, memory_pos=None, self_attn_mask=None, cross_attn_mask=None,
                self_attn_mask_bits=None, memory_grad_chain=None):
        """Batch-first: tgt/query_pos [N,nq,C], reference points [N,nq,L,4], memory [N,S,C]."""
        for step in self.module_seq:
            if step == "sa":
                tgt = self.forward_sa(tgt, tgt_query_pos, self_attn_mask, self_attn_mask_bits)
            elif step == "ca":
                tgt = self.forward_ca(tgt, tgt_query_pos, tgt_reference_points, memory, memory_key_padding_mask,
                                      memory_level_start_index, memory_spatial_shapes, memory_grad_chain)
            elif step == "ffn":
                tgt = self.forward_ffn(tgt)
            else:
                raise ValueError(f"unknown funcname {step}")
        return tgt


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False, d_model=256, query_dim=4,
                 modulate_hw_attn=False, num_feature_levels=1, deformable_decoder=False, decoder_query_perturber=None,
                 dec_layer_number=None, rm_dec_query_scale=False, dec_layer_share=False, dec_layer_dropout_prob=None,
                 use_detached_boxes_dec_out=False):
        super().__init__()
        assert return_intermediate, "support return_intermediate only"
        assert query_dim in (2, 4), f"query_dim should be 2/4 but {query_dim}"
        if not deformable_decoder or not rm_dec_query_scale or dec_layer_number is not None \
                or dec_layer_dropout_prob is not None:
            raise NotImplementedError("only the deformable decoder without query scaling / query pruning is supported")
        self.layers = _get_clones(decoder_layer, num_layers, layer_share=dec_layer_share) if num_layers > 0 else []
        self.num_layers, self.norm, self.return_intermediate = num_layers, norm, return_intermediate
        self.query_dim, self.num_feature_levels = query_dim, num_feature_levels
        self.use_detached_boxes_dec_out = use_detached_boxes_dec_out
        self.ref_point_head = MLP(query_dim // 2 * d_model, d_model, d_model, 2)
        self.query_pos_sine_scale = None
        self.query_scale = None
        self.bbox_embed = None          # set by DINO (shared prediction heads)
        self.class_embed = None
        self.d_model, self.modulate_hw_attn, self.deformable_decoder = d_model, modulate_hw_attn, deformable_decoder
        self.ref_anchor_head = None
        self.decoder_query_perturber = decoder_query_perturber
        self.box_pred_damping = None
        self.dec_layer_number = dec_layer_number
        self.dec_layer_dropout_prob = dec_layer_dropout_prob
        self.rm_detach = None

    def forward(self, tgt, memory, tgt_mask: Optional[Tensor] = None, memory_mask: Optional[Tensor] = None,
                tgt_key_padding_mask: Optional[Tensor] = None, memory_key_padding_mask: Optional[Tensor] = None,
                pos: Optional[Tensor] = None, refpoints_unsigmoid: Optional[Tensor] = None,
                level_start_index: Optional[Tensor] = None, spatial_shapes: Optional[Tensor] = None,
                valid_ratios: Optional[Tensor] = None):
        """Batch-first: tgt [N,nq,C], memory [N,S,C], refpoints_unsigmoid [N,nq,4], valid_ratios [N,L,2].
        Returns ([per-layer normed output [N,nq,C]], [reference boxes [N,nq,4], one more than layers])."""
        out = tgt
        ref = refpoints_unsigmoid.sigmoid()
        refs, inter = [ref], []
        vr = torch.cat([valid_ratios, valid_ratios], -1)[:, None] if ref.shape[-1] == 4 else valid_ratios[:, None]
        # the attention mask is the same for every layer: pack it once for the fused self-attention kernel
        mask_bits = None
        if (_fused_attention_on() and tgt.is_cuda and tgt.dtype == torch.float32
                and (tgt_mask is None or tgt_mask.dtype == torch.bool)):
            mask_bits = attention.pack_mask(tgt_mask, tgt.shape[1], tgt.device, transposed=torch.is_grad_enabled())
        # every layer projects the same memory (cross-attention values): their input gradients are summed along a chain
        # inside the input-gradient GEMMs instead of five accumulation passes over [N, S, C] (linear.GradChain)
        chain = None
        if (_MEMORY_GRAD_CHAIN and len(self.layers) > 1 and memory.is_cuda and memory.requires_grad and torch.is_grad_enabled()
                and dl.get_mode() == "tf32" and memory.dtype == torch.float32
                and all(getattr(l, "module_seq", ()).count("ca") == 1 and dl.eligible(memory, l.cross_attn.value_proj.weight)
                        for l in self.layers)):
            chain = dl.GradChain(len(self.layers))
        for lid, layer in enumerate(self.layers):
            if self.training and self.decoder_query_perturber is not None and lid != 0:
                ref = self.decoder_query_perturber(ref)
            ref_in = ref[:, :, None] * vr                                        # [N,nq,L,4]
            query_pos = self.ref_point_head(gen_sineembed_for_position(ref_in[:, :, 0, :]))
            out = layer(tgt=out, tgt_query_pos=query_pos, tgt_reference_points=ref_in, memory=memory,
                        memory_key_padding_mask=memory_key_padding_mask, memory_level_start_index=level_start_index,
                        memory_spatial_shapes=spatial_shapes, memory_pos=pos, self_attn_mask=tgt_mask,
                        self_attn_mask_bits=mask_bits, memory_grad_chain=chain)
            if self.bbox_embed is not None:
                new_ref = (self.bbox_embed[lid](out) + inverse_sigmoid(ref)).sigmoid()
                ref = new_ref if (self.rm_detach and "dec" in self.rm_detach) else new_ref.detach()
                refs.append(ref if self.use_detached_boxes_dec_out else new_ref)
            inter.append(ln(self.norm, out))
        return [inter, refs]


class DeformableTransformer(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_queries=300, num_encoder_layers=6, num_unicoder_layers=0,
                 num_decoder_layers=6, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False,
                 return_intermediate_dec=False, query_dim=4, num_patterns=0, modulate_hw_attn=False,
                 deformable_encoder=False, deformable_decoder=False, num_feature_levels=1, enc_n_points=4,
                 dec_n_points=4, use_deformable_box_attn=False, box_attn_type="roi_align", learnable_tgt_init=False,
                 decoder_query_perturber=None, add_channel_attention=False, add_pos_value=False,
                 random_refpoints_xy=False, two_stage_type="no", two_stage_pat_embed=0, two_stage_add_query_num=0,
                 two_stage_learn_wh=False, two_stage_keep_all_tokens=False, dec_layer_number=None,
                 rm_enc_query_scale=True, rm_dec_query_scale=True, rm_self_attn_layers=None, key_aware_type=None,
                 layer_share_type=None, rm_detach=None, decoder_sa_type="ca", module_seq=("sa", "ca", "ffn"),
                 embed_init_tgt=False, use_detached_boxes_dec_out=False):
        super().__init__()
        assert query_dim == 4
        assert layer_share_type is None
        assert learnable_tgt_init, "why not learnable_tgt_init"
        assert two_stage_type in ("no", "standard"), f"unknown param {two_stage_type} of two_stage_type"
        if not (deformable_encoder and deformable_decoder):
            raise NotImplementedError("only the deformable encoder/decoder are supported")
        if num_patterns or two_stage_pat_embed or two_stage_add_query_num:
            raise NotImplementedError("pattern embeddings / extra two-stage queries are outside the DINO hot path")
        self.num_feature_levels = num_feature_levels
        self.num_encoder_layers, self.num_unicoder_layers = num_encoder_layers, num_unicoder_layers
        self.num_decoder_layers = self.dec_layers = num_decoder_layers
        self.deformable_encoder, self.deformable_decoder = deformable_encoder, deformable_decoder
        self.two_stage_keep_all_tokens = two_stage_keep_all_tokens
        self.num_queries, self.random_refpoints_xy = num_queries, random_refpoints_xy
        self.use_detached_boxes_dec_out = use_detached_boxes_dec_out
        self.decoder_sa_type = decoder_sa_type
        self.d_model, self.nhead, self.num_patterns = d_model, nhead, 0

        enc_layer = DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                      nhead, enc_n_points, add_channel_attention=add_channel_attention,
                                                      use_deformable_box_attn=use_deformable_box_attn,
                                                      box_attn_type=box_attn_type)
        self.encoder = TransformerEncoder(enc_layer, num_encoder_layers, nn.LayerNorm(d_model) if normalize_before else None,
                                          d_model=d_model, num_queries=num_queries, deformable_encoder=True,
                                          two_stage_type=two_stage_type)
        dec_layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                      nhead, dec_n_points, use_deformable_box_attn=use_deformable_box_attn,
                                                      box_attn_type=box_attn_type, key_aware_type=key_aware_type,
                                                      decoder_sa_type=decoder_sa_type, module_seq=module_seq)
        self.decoder = TransformerDecoder(dec_layer, num_decoder_layers, nn.LayerNorm(d_model),
                                          return_intermediate=return_intermediate_dec, d_model=d_model,
                                          query_dim=query_dim, modulate_hw_attn=modulate_hw_attn,
                                          num_feature_levels=num_feature_levels, deformable_decoder=True,
                                          decoder_query_perturber=decoder_query_perturber,
                                          dec_layer_number=dec_layer_number, rm_dec_query_scale=rm_dec_query_scale,
                                          use_detached_boxes_dec_out=use_detached_boxes_dec_out)

        self.level_embed = None
        if num_feature_levels > 1 and num_encoder_layers > 0:
            self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self.learnable_tgt_init, self.embed_init_tgt = learnable_tgt_init, embed_init_tgt
        self.tgt_embed = None
        if (two_stage_type != "no" and embed_init_tgt) or two_stage_type == "no":
            self.tgt_embed = nn.Embedding(num_queries, d_model)
            nn.init.normal_(self.tgt_embed.weight.data)

        self.two_stage_type, self.two_stage_pat_embed = two_stage_type, two_stage_pat_embed
        self.two_stage_add_query_num, self.two_stage_learn_wh = two_stage_add_query_num, two_stage_learn_wh
        if two_stage_type == "standard":
            self.enc_output = nn.Linear(d_model, d_model)
            self.enc_output_norm = nn.LayerNorm(d_model)
            self.two_stage_wh_embedding = nn.Embedding(1, 2) if two_stage_learn_wh else None
        if two_stage_type == "no":
            self.init_ref_points(num_queries)
        self.enc_out_class_embed = None     # set by DINO
        self.enc_out_bbox_embed = None
        self.dec_layer_number = dec_layer_number
        self._reset_parameters()

        self.rm_self_attn_layers = rm_self_attn_layers
        if rm_self_attn_layers is not None:
            for lid, layer in enumerate(self.decoder.layers):
                if lid in rm_self_attn_layers:
                    layer.rm_self_attn_modules()
        self.rm_detach = rm_detach
        if rm_detach:
            assert isinstance(rm_detach, list) and any(i in ("enc_ref", "enc_tgt", "dec") for i in rm_detach)
        self.decoder.rm_detach = rm_detach

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        if self.level_embed is not None:
            nn.init.normal_(self.level_embed)
        if self.two_stage_learn_wh:
            nn.init.constant_(self.two_stage_wh_embedding.weight, math.log(0.05 / (1 - 0.05)))

    def get_valid_ratio(self, mask):
        """[N,H,W] padding mask -> [N,2] (valid width / W, valid height / H)."""
        _, H, W = mask.shape
        keep = ~mask
        return torch.stack([keep[:, 0, :].sum(1).float() / W, keep[:, :, 0].sum(1).float() / H], -1)

    def init_ref_points(self, use_num_queries):
        self.refpoint_embed = nn.Embedding(use_num_queries, 4)
        if self.random_refpoints_xy:
            self.refpoint_embed.weight.data[:, :2].uniform_(0, 1)
            self.refpoint_embed.weight.data[:, :2] = inverse_sigmoid(self.refpoint_embed.weight.data[:, :2])
            self.refpoint_embed.weight.data[:, :2].requires_grad = False

    def _select_queries(self, memory, mask_flat, refpoint_embed, tgt, shapes_list):
        """Two-stage query selection (pure device work): encoder proposals -> top-k -> decoder queries / references,
        concatenated behind the de-noising queries.  Returns (refpoints_unsigmoid, tgt, tgt_undetach, refpoint_undetach,
        init_box_proposal, output_memory, coord_all, output_proposals); unused entries are None."""
        bs = memory.shape[0]
        tgt_undetach = refpoint_undetach = output_memory = coord_all = output_proposals = None
        if self.two_stage_type == "standard":
            input_hw = self.two_stage_wh_embedding.weight[0] if self.two_stage_learn_wh else None
            output_memory, output_proposals = gen_encoder_output_proposals(memory, mask_flat, shapes_list, input_hw)
            output_memory = ln(self.enc_output_norm, dl.linear(output_memory, self.enc_output.weight, self.enc_output.bias))
            with dl.fp32_products():    # feeds the top-k below: fp32 products, so equal inputs give equal indices
                class_all = self.enc_out_class_embed(output_memory)
            topk = torch.topk(class_all.max(-1)[0], self.num_queries, dim=1)[1]               # [N,nq] int64
            tgt_undetach = torch.gather(output_memory, 1, topk.unsqueeze(-1).expand(-1, -1, self.d_model))
            prop_sel = torch.gather(output_proposals, 1, topk.unsqueeze(-1).expand(-1, -1, 4))
            refpoint_undetach = self.enc_out_bbox_embed(tgt_undetach) + prop_sel              # logits
            refpoint_sel = refpoint_undetach.detach()
            init_box_proposal = prop_sel.sigmoid()
            tgt_sel = self.tgt_embed.weight[None].expand(bs, -1, -1) if self.embed_init_tgt else tgt_undetach.detach()
            if self.two_stage_keep_all_tokens:
                coord_all = self.enc_out_bbox_embed(output_memory) + output_proposals
        else:
            tgt_sel = self.tgt_embed.weight[None].expand(bs, -1, -1)
            refpoint_sel = self.refpoint_embed.weight[None].expand(bs, -1, -1)
            init_box_proposal = refpoint_sel.sigmoid()
        if refpoint_embed is not None:
            refpoint_embed = torch.cat([refpoint_embed, refpoint_sel], dim=1)
            tgt = torch.cat([tgt, tgt_sel], dim=1)
        else:
            refpoint_embed, tgt = refpoint_sel, tgt_sel.contiguous()   # materialise the broadcast of the embedding

        return refpoint_embed, tgt, tgt_undetach, refpoint_undetach, init_box_proposal, output_memory, coord_all, output_proposals

    def _flatten_levels(self, srcs, masks, pos_embeds):
        """Per-level [N,C,H,W] maps -> token sequences [N,S,C] (+ level embedding on the positions), the flat
        padding mask [N,S] and the valid ratios [N,L,2]: pure device work."""
        feats, poses = [], []
        for lvl, (src, pos) in enumerate(zip(srcs, pos_embeds)):
            feats.append(src.flatten(2).transpose(1, 2))
            pos = pos.flatten(2).transpose(1, 2)
            if self.num_feature_levels > 1 and self.level_embed is not None:
                pos = pos + self.level_embed[lvl].view(1, 1, -1)
            poses.append(pos)
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        return torch.cat(feats, 1), torch.cat(poses, 1), torch.cat([m.flatten(1) for m in masks], 1), valid_ratios

    def _shape_tensors(self, shapes_list, device):
        """(spatial_shapes [L,2], level_start_index [L]) int64 on the device, cached per shape list (the reference
        rebuilds them from python lists -- two pageable host->device copies -- at every call)."""
        cache = self.__dict__.setdefault("_shape_cache", {})
        key = (tuple(shapes_list), str(device))
        if key not in cache:
            sizes = [h * w for h, w in shapes_list]
            cache[key] = (torch.as_tensor(shapes_list, dtype=torch.long, device=device),
                          torch.as_tensor([sum(sizes[:i]) for i in range(len(sizes))], dtype=torch.long, device=device))
        return cache[key]

    def forward(self, srcs, masks, refpoint_embed, pos_embeds, tgt, attn_mask=None):
        """srcs / pos_embeds: per level [N,C,H,W]; masks: per level [N,H,W] (True = padding);
        refpoint_embed [N,n_dn,4] / tgt [N,n_dn,C]: de-noising queries (None at inference).
        Returns (hs: list of [N,nq,C] per decoder layer, references: list of [N,nq,4] (layers + 1),
        hs_enc [1,N,nq,C] | None, ref_enc [1,N,nq,4] | None, init_box_proposal [N,nq,4])."""
        return self.decode(self.encode(srcs, masks, pos_embeds), refpoint_embed, tgt, attn_mask)

    def encode(self, srcs, masks, pos_embeds):
        """Level flattening + the deformable encoder (reference forward :267-316).  Every operation of this half works on
        one image at a time (token-wise GEMMs / LayerNorm, per-image MSDeformAttn), so DINO.forward runs it ONCE on the
        source and target halves of a domain-adaptation batch together and hands each half to decode(): half the kernel
        launches of two separate passes, twice the rows per GEMM, one gradient arrival per encoder parameter."""
        shapes_list = [tuple(s.shape[-2:]) for s in srcs]
        if graphs.ACTIVE is not None and srcs[0].is_cuda:
            owner = self.__dict__.get("_level_embed_owner")
            if owner is None:                                   # exposes the bare level_embed parameter to the segment
                owner = nn.Module()
                if self.level_embed is not None:
                    owner.register_parameter("level_embed", self.level_embed)
                self.__dict__["_level_embed_owner"] = owner
            src_flat, pos_flat, mask_flat, valid_ratios = graphs.ACTIVE.call(
                "flatten", owner, self._flatten_levels, tuple(srcs), tuple(masks), tuple(pos_embeds))
        else:
            src_flat, pos_flat, mask_flat, valid_ratios = self._flatten_levels(tuple(srcs), tuple(masks), tuple(pos_embeds))
        spatial_shapes, level_start_index = self._shape_tensors(shapes_list, src_flat.device)
        # outputs of a graph segment all come back as differentiable (one autograd node per segment); the geometry is
        # not -- and reference points that "require grad" would push MSDeformAttn off its fused kernels
        valid_ratios = valid_ratios.detach()

        if graphs.ACTIVE is not None and src_flat.is_cuda:
            memory = graphs.ACTIVE.run("encoder", lambda: graphs.EncoderSegment(self.encoder, shapes_list),
                                       (src_flat, pos_flat, spatial_shapes, level_start_index, valid_ratios, mask_flat),
                                       owner=self.encoder)
        else:
            memory, _, _ = self.encoder(src_flat, pos=pos_flat, level_start_index=level_start_index,
                                        spatial_shapes=spatial_shapes, valid_ratios=valid_ratios,
                                        key_padding_mask=mask_flat, shapes_list=shapes_list)
        return memory, pos_flat, mask_flat, valid_ratios, spatial_shapes, level_start_index, shapes_list

    @staticmethod
    def split_encoded(enc, sizes):
        """Cut the batch axis of encode()'s result into consecutive parts of `sizes` images (torch.split: the backward
        of all parts is ONE concatenation into the encoder's output gradient)."""
        memory, pos_flat, mask_flat, valid_ratios, spatial_shapes, level_start_index, shapes_list = enc
        parts = zip(memory.split(sizes), pos_flat.split(sizes), mask_flat.split(sizes), valid_ratios.split(sizes))
        return [(m, p, k, v, spatial_shapes, level_start_index, shapes_list) for m, p, k, v in parts]

    def decode(self, enc, refpoint_embed, tgt, attn_mask=None):
        """Two-stage query selection + decoder on an encode() result (reference forward :318-431)."""
        memory, pos_flat, mask_flat, valid_ratios, spatial_shapes, level_start_index, shapes_list = enc
        if graphs.ACTIVE is not None and memory.is_cuda:
            owners = self.__dict__.setdefault("_two_stage_owner", nn.ModuleList(
                [m for m in (getattr(self, "enc_output", None), getattr(self, "enc_output_norm", None),
                             self.enc_out_class_embed, self.enc_out_bbox_embed, self.tgt_embed,
                             getattr(self, "refpoint_embed", None), getattr(self, "two_stage_wh_embedding", None))
                 if isinstance(m, nn.Module)]))
            sel = graphs.ACTIVE.call("two_stage", owners, self._select_queries, memory, mask_flat, refpoint_embed, tgt,
                                     tuple(shapes_list))
        else:
            sel = self._select_queries(memory, mask_flat, refpoint_embed, tgt, tuple(shapes_list))
        refpoint_embed, tgt, tgt_undetach, refpoint_undetach, init_box_proposal, output_memory, coord_all, output_proposals = sel
        # the decoder's reference boxes carry no gradient (selected proposals are detached, :345 of the reference; the
        # de-noising boxes are inputs); as the output of a graph segment they would look differentiable
        if self.two_stage_type == "standard":      # 'no': the learnable refpoint_embed receives gradients through the decoder
            refpoint_embed = refpoint_embed.detach()

        if graphs.ACTIVE is not None and tgt.is_cuda:
            dec_args = (tgt, memory, mask_flat, pos_flat, refpoint_embed, level_start_index, spatial_shapes, valid_ratios)
            flat = graphs.ACTIVE.run("decoder", lambda: graphs.DecoderSegment(self.decoder, attn_mask is not None),
                                     dec_args + ((attn_mask,) if attn_mask is not None else ()), owner=self.decoder)
            n_layers = len(self.decoder.layers)
            hs, references = list(flat[:n_layers]), list(flat[n_layers:])
        else:
            hs, references = self.decoder(tgt=tgt, memory=memory, memory_key_padding_mask=mask_flat, pos=pos_flat,
                                          refpoints_unsigmoid=refpoint_embed, level_start_index=level_start_index,
                                          spatial_shapes=spatial_shapes, valid_ratios=valid_ratios, tgt_mask=attn_mask)

        hs_enc = ref_enc = None
        if self.two_stage_type == "standard":
            if self.two_stage_keep_all_tokens:
                hs_enc, ref_enc, init_box_proposal = output_memory.unsqueeze(0), coord_all.unsqueeze(0), output_proposals
            else:
                hs_enc, ref_enc = tgt_undetach.unsqueeze(0), refpoint_undetach.sigmoid().unsqueeze(0)
        return hs, references, hs_enc, ref_enc, init_box_proposal


def build_deformable_transformer(args):
    perturber = None
    if args.decoder_layer_noise:
        from .utils import RandomBoxPerturber
        perturber = RandomBoxPerturber(x_noise_scale=args.dln_xy_noise, y_noise_scale=args.dln_xy_noise,
                                       w_noise_scale=args.dln_hw_noise, h_noise_scale=args.dln_hw_noise)
    return DeformableTransformer(
        d_model=args.hidden_dim, dropout=args.dropout, nhead=args.nheads, num_queries=args.num_queries,
        dim_feedforward=args.dim_feedforward, num_encoder_layers=args.enc_layers,
        num_unicoder_layers=args.unic_layers, num_decoder_layers=args.dec_layers, normalize_before=args.pre_norm,
        return_intermediate_dec=True, query_dim=args.query_dim, activation=args.transformer_activation,
        num_patterns=args.num_patterns, modulate_hw_attn=True, deformable_encoder=True, deformable_decoder=True,
        num_feature_levels=args.num_feature_levels, enc_n_points=args.enc_n_points, dec_n_points=args.dec_n_points,
        use_deformable_box_attn=args.use_deformable_box_attn, box_attn_type=args.box_attn_type,
        learnable_tgt_init=True, decoder_query_perturber=perturber, add_channel_attention=args.add_channel_attention,
        add_pos_value=args.add_pos_value, random_refpoints_xy=args.random_refpoints_xy,
        two_stage_type=args.two_stage_type, two_stage_pat_embed=args.two_stage_pat_embed,
        two_stage_add_query_num=args.two_stage_add_query_num, two_stage_learn_wh=args.two_stage_learn_wh,
        two_stage_keep_all_tokens=args.two_stage_keep_all_tokens, dec_layer_number=args.dec_layer_number,
        rm_self_attn_layers=None, key_aware_type=None, layer_share_type=None, rm_detach=None,
        decoder_sa_type=args.decoder_sa_type, module_seq=args.decoder_module_seq,
        embed_init_tgt=args.embed_init_tgt, use_detached_boxes_dec_out=getattr(args, "use_detached_boxes_dec_out", False))
