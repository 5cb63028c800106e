"""Default hyper-parameters of the DINO-4scale / 5scale configs the benchmark uses
(reference config/DINO/DINO_4scale.py:3-112, DINO_5scale.py:12,32,55, plus the DA keys of
config/DA/*/DINO_4scale_*.py:90-126 that build_dino reads).  `dino_args(**overrides)` returns an
argparse-style namespace accepted by models.dino.build_dino."""
from types import SimpleNamespace

_DEFAULTS = dict(
    num_classes=91, device="cuda", modelname="dino", backbone="resnet50", use_checkpoint=False, dilation=False,
    position_embedding="sine", pe_temperatureH=20, pe_temperatureW=20, return_interm_indices=[1, 2, 3],
    backbone_freeze_keywords=None, lr_backbone=1e-5, enc_layers=6, dec_layers=6, unic_layers=0, pre_norm=False,
    dim_feedforward=2048, hidden_dim=256, dropout=0.0, nheads=8, num_queries=900, query_dim=4, num_patterns=0,
    random_refpoints_xy=False, fix_refpoints_hw=-1, use_deformable_box_attn=False, box_attn_type="roi_align",
    dec_layer_number=None, num_feature_levels=4, enc_n_points=4, dec_n_points=4, decoder_layer_noise=False,
    dln_xy_noise=0.2, dln_hw_noise=0.2, add_channel_attention=False, add_pos_value=False, two_stage_type="standard",
    two_stage_pat_embed=0, two_stage_add_query_num=0, two_stage_bbox_embed_share=False,
    two_stage_class_embed_share=False, two_stage_learn_wh=False, two_stage_default_hw=0.05,
    two_stage_keep_all_tokens=False, num_select=300, transformer_activation="relu", masks=False, aux_loss=True,
    set_cost_class=2.0, set_cost_bbox=5.0, set_cost_giou=2.0, cls_loss_coef=1.0, bbox_loss_coef=5.0,
    giou_loss_coef=2.0, enc_loss_coef=1.0, interm_loss_coef=1.0, no_interm_box_loss=False, focal_alpha=0.25,
    decoder_sa_type="sa", matcher_type="HungarianMatcher", decoder_module_seq=["sa", "ca", "ffn"],
    nms_iou_threshold=-1, dec_pred_bbox_embed_share=True, dec_pred_class_embed_share=True, use_dn=True,
    dn_number=100, dn_box_noise_scale=0.4, dn_label_noise_ratio=0.5, embed_init_tgt=True, dn_labelbook_size=91,
    match_unstable_error=True, use_detached_boxes_dec_out=False, frozen_weights=None, dataset_file="coco",
    da_backbone_loss_coef=0.1, da_proto_loss_coef=0.1, da_global_proto_coef=0.1, self_training_loss_coef=1.0,
    lr=1e-4, weight_decay=1e-4, clip_max_norm=0.1, batch_size=2,
)


def dino_args(**overrides):
    d = dict(_DEFAULTS)
    d.update(overrides)
    return SimpleNamespace(**d)


def dino_5scale_args(**overrides):
    return dino_args(**{"return_interm_indices": [0, 1, 2, 3], "num_feature_levels": 5, "batch_size": 1, **overrides})
