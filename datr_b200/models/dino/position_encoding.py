"""Sine position embedding of the feature maps.

Mirrors PositionEmbeddingSineHW (reference models/dino/position_encoding.py:62-107; temperature 20,
normalize=True in the DINO configs) and build_position_encoding (:138-153).
"""
import math

import torch
from torch import nn


class PositionEmbeddingSineHW(nn.Module):
    def __init__(self, num_pos_feats=64, temperatureH=10000, temperatureW=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and not normalize:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats = num_pos_feats
        self.temperatureH, self.temperatureW = temperatureH, temperatureW
        self.normalize = normalize
        self.scale = 2 * math.pi if scale is None else scale

    def _axis(self, coord, temperature):
        i = torch.arange(self.num_pos_feats, dtype=torch.float32, device=coord.device)
        dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / self.num_pos_feats)
        ang = coord[..., None] / dim_t
        return torch.stack((ang[..., 0::2].sin(), ang[..., 1::2].cos()), dim=-1).flatten(-2)

    def forward(self, tensor_list):
        mask = tensor_list.mask
        assert mask is not None
        keep = ~mask
        y = keep.cumsum(1, dtype=torch.float32)
        x = keep.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y = y / (y[:, -1:, :] + 1e-6) * self.scale
            x = x / (x[:, :, -1:] + 1e-6) * self.scale
        if y.is_cuda and (str(y.device) in self.__dict__.get("_dim_t", {}) or not torch.cuda.is_current_stream_capturing()):
            # one kernel for both axes (csrc/decoder_ops.cu) instead of ~20 ATen launches per level; no gradient involved
            from datr_b200 import native
            tabs = self.__dict__.setdefault("_dim_t", {})
            if str(y.device) not in tabs:
                i = torch.arange(self.num_pos_feats, dtype=torch.float32, device=y.device)
                e = 2 * torch.div(i, 2, rounding_mode="floor") / self.num_pos_feats
                tabs[str(y.device)] = (self.temperatureH ** e, self.temperatureW ** e)
            th, tw = tabs[str(y.device)]
            y, x = y.contiguous(), x.contiguous()
            pos = torch.empty(y.shape + (2 * self.num_pos_feats,), dtype=torch.float32, device=y.device)
            lib = native.lib()
            with torch.cuda.device(y.device):
                rc = lib.datr_pos_embed_hw(y.data_ptr(), x.data_ptr(), th.data_ptr(), tw.data_ptr(), y.numel(),
                                           self.num_pos_feats, pos.data_ptr(), torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise RuntimeError(f"datr_pos_embed_hw failed (code {rc}): {lib.datr_decoder_ops_last_error().decode()}")
            return pos.permute(0, 3, 1, 2)
        pos = torch.cat((self._axis(y, self.temperatureH), self._axis(x, self.temperatureW)), dim=3)
        return pos.permute(0, 3, 1, 2)


class PositionEmbeddingLearned(nn.Module):
    def __init__(self, num_pos_feats=256):
        super().__init__()
        self.row_embed = nn.Embedding(50, num_pos_feats)
        self.col_embed = nn.Embedding(50, num_pos_feats)
        nn.init.uniform_(self.row_embed.weight)
        nn.init.uniform_(self.col_embed.weight)

    def forward(self, tensor_list):
        x = tensor_list.tensors
        h, w = x.shape[-2:]
        col = self.col_embed(torch.arange(w, device=x.device))[None].expand(h, -1, -1)
        row = self.row_embed(torch.arange(h, device=x.device))[:, None].expand(-1, w, -1)
        return torch.cat([col, row], -1).permute(2, 0, 1)[None].expand(x.shape[0], -1, -1, -1)


def build_position_encoding(args):
    n = args.hidden_dim // 2
    if args.position_embedding in ("v2", "sine"):
        return PositionEmbeddingSineHW(n, temperatureH=args.pe_temperatureH, temperatureW=args.pe_temperatureW,
                                       normalize=True)
    if args.position_embedding in ("v3", "learned"):
        return PositionEmbeddingLearned(n)
    raise ValueError(f"not supported {args.position_embedding}")
