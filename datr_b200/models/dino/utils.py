"""Small pieces called inside the transformer forward.

Mirrors models/dino/utils.py of the reference: gen_encoder_output_proposals (:15-61), MLP (:107-119),
sigmoid_focal_loss (:79-104), gen_sineembed_for_position (:138-163), RandomBoxPerturber (:64-76),
_get_activation_fn (:122-135).  Written batch-first and without Python-side host syncs.
"""
import math

import torch
import torch.nn.functional as F

from datr_b200 import linear as dl
from torch import nn


def level_sizes(spatial_shapes):
    """[(H, W), ...] as Python ints.  Accepts the int64 tensor the op takes or a list (no sync for a list)."""
    if isinstance(spatial_shapes, torch.Tensor):
        return [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    return [(int(h), int(w)) for h, w in spatial_shapes]


def gen_encoder_output_proposals(memory, memory_padding_mask, spatial_shapes, learnedwh=None):
    """Per-token anchor boxes for the two-stage query selection.

    memory [N,S,C], memory_padding_mask [N,S] (True = padding), spatial_shapes [L,2].
    Token (y,x) of level l proposes the box centre ((x+.5)/validW, (y+.5)/validH) with side
    0.05 * 2^l (or sigmoid(learnedwh) * 2^l), in logit space.  Proposals with any coordinate outside
    (0.01, 0.99) or on padding become +inf and their memory rows are zeroed.
    Returns (output_memory [N,S,C], output_proposals [N,S,4])."""
    N = memory.shape[0]
    dev = memory.device
    boxes, start = [], 0
    for lvl, (H, W) in enumerate(level_sizes(spatial_shapes)):
        keep = ~memory_padding_mask[:, start:start + H * W].view(N, H, W)
        valid_h = keep[:, :, 0].sum(1)
        valid_w = keep[:, 0, :].sum(1)
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                                torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
        centre = torch.stack([xs, ys], -1)[None].expand(N, -1, -1, -1) + 0.5
        centre = centre / torch.stack([valid_w, valid_h], 1).view(N, 1, 1, 2)
        side = (learnedwh.sigmoid() if learnedwh is not None else 0.05) * (2.0 ** lvl)
        wh = torch.ones_like(centre) * side
        boxes.append(torch.cat([centre, wh], -1).view(N, H * W, 4))
        start += H * W
    prop = torch.cat(boxes, 1)
    inside = ((prop > 0.01) & (prop < 0.99)).all(-1, keepdim=True)
    dead = memory_padding_mask.unsqueeze(-1) | ~inside
    prop = torch.log(prop / (1 - prop)).masked_fill(dead, float("inf"))
    return memory.masked_fill(dead, 0.0), prop


class RandomBoxPerturber:
    def __init__(self, x_noise_scale=0.2, y_noise_scale=0.2, w_noise_scale=0.2, h_noise_scale=0.2):
        self.noise_scale = torch.tensor([x_noise_scale, y_noise_scale, w_noise_scale, h_noise_scale])

    def __call__(self, refanchors):
        scale = self.noise_scale.to(refanchors.device)[:refanchors.shape[-1]]
        return (refanchors * (1 + (torch.rand_like(refanchors) - 0.5) * scale)).clamp_(0, 1)


def sigmoid_focal_loss(inputs, targets, num_boxes, alpha: float = 0.25, gamma: float = 2):
    """RetinaNet focal loss on logits; mean over dim 1 (queries), summed, divided by num_boxes."""
    p = inputs.sigmoid()
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.mean(1).sum() / num_boxes


class MLP(nn.Module):
    """Linear -> ReLU -> ... -> Linear; parameters live in `layers.N` like the reference."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):      # bias + ReLU ride in the GEMM epilogue (datr_b200.linear)
            x = dl.linear(x, layer.weight, layer.bias, relu=i + 1 < self.num_layers)
        return x


def _get_activation_fn(activation, d_model=256, batch_dim=0):
    table = {"relu": F.relu, "gelu": F.gelu, "glu": F.glu, "selu": F.selu}
    if activation == "prelu":
        return nn.PReLU()
    if activation not in table:
        raise RuntimeError(f"activation should be relu/gelu, not {activation}.")
    return table[activation]


_DIM_T = {}


def _sine_dim_t(device):
    """10000 ** (2 * (i // 2) / 128), i = 0..127, computed by torch (once per device, outside of graph capture)."""
    t = _DIM_T.get(str(device))
    if t is None:
        idx = torch.arange(128, dtype=torch.float32, device=device)
        t = _DIM_T[str(device)] = 10000 ** (2 * torch.div(idx, 2, rounding_mode="floor") / 128)
    return t


def gen_sineembed_for_position(pos_tensor):
    """[..., 2|4] normalised (x, y[, w, h]) -> [..., 128 * k] sine embedding ordered (y, x[, w, h]);
    128 features per coordinate, temperature 10000, sin on even / cos on odd feature indices.
    CUDA fp32 input without gradient (the decoder's detached reference boxes): one kernel (csrc/decoder_ops.cu)."""
    k = pos_tensor.size(-1)
    if k not in (2, 4):
        raise ValueError(f"Unknown pos_tensor shape(-1):{k}")
    if (pos_tensor.is_cuda and pos_tensor.dtype == torch.float32 and not pos_tensor.requires_grad
            and (str(pos_tensor.device) in _DIM_T or not torch.cuda.is_current_stream_capturing())):
        from datr_b200 import native
        lib = native.lib()
        src = pos_tensor.contiguous()
        out = torch.empty(pos_tensor.shape[:-1] + (128 * k,), dtype=torch.float32, device=pos_tensor.device)
        with torch.cuda.device(pos_tensor.device):
            rc = lib.datr_sine_embed(src.data_ptr(), _sine_dim_t(pos_tensor.device).data_ptr(), src.numel() // k, k,
                                     out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"datr_sine_embed failed (code {rc}): {lib.datr_decoder_ops_last_error().decode()}")
        return out
    idx = torch.arange(128, dtype=torch.float32, device=pos_tensor.device)
    dim_t = 10000 ** (2 * torch.div(idx, 2, rounding_mode="floor") / 128)
    ang = pos_tensor.unsqueeze(-1) * (2 * math.pi) / dim_t                       # [..., k, 128]
    emb = torch.stack((ang[..., 0::2].sin(), ang[..., 1::2].cos()), dim=-1).flatten(-2)
    # (y, x[, w, h]) order by slicing: an index list would be uploaded from the host (not CUDA-graph capturable)
    return torch.cat((emb[..., 1:2, :], emb[..., 0:1, :], emb[..., 2:, :]), dim=-2).flatten(-2)
