"""MSDeformAttn module -- host-side mirror of the reference's
models/dino/ops/modules/ms_deform_attn.py:30-126.

Kept identical on purpose (drop-in / checkpoint compatibility): constructor signature, the
attributes `im2col_step, d_model, n_levels, n_heads, n_points`, the four sub-modules named
`sampling_offsets, attention_weights, value_proj, output_proj` (state_dict keys), the
initialisation of `_reset_parameters` (:62-76) and the forward argument order (:78).
"""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from datr_b200 import MultiScaleDeformableAttention as MSDA
from datr_b200 import linear as dl

from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnMergedFunction

# Module-level fusion: softmax + sampling-location arithmetic inside the MSDeformAttn kernels.
# DATR_MSDA_FUSED=0 (or set_fused(False)) composes the reference's op with the torch prologue instead.
import os
_FUSED = os.environ.get("DATR_MSDA_FUSED", "1") != "0"


_SKIP_CARRIER = os.environ.get("DATR_MSDA_SKIP_CARRIER", "1") != "0"


def set_fused(on: bool) -> None:
    global _FUSED
    _FUSED = bool(on)


# Storage of the value map the fused forward gathers from: "fp32" (the reference's rows, default), or 16-bit pair rows
# "bf16" / "fp16" (datr_msda_pack_value_pairs: half the line gathers per sample; precision class 1e-2 / 1e-3).
_VALUE_STORAGE = {"fp32": None, "bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("DATR_MSDA_VALUE", "fp32")]


def set_value_storage(kind: str) -> None:
    global _VALUE_STORAGE
    _VALUE_STORAGE = {"fp32": None, "bf16": torch.bfloat16, "fp16": torch.float16}[kind]


def _is_power_of_2(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


_CHECKED = set()


def _check_levels(spatial_shapes, S):
    """The reference's `assert (H*W).sum() == S` (ms_deform_attn.py:92).  The kernels clamp every sample inside its
    level but trust level_start + H*W <= S, so inconsistent shapes would read -- and, in the backward, atomically add --
    out of bounds.  The reference pays a device->host sync for this at every call; here a given shapes tensor (the
    transformer caches them per feature-map geometry) is validated once, outside of CUDA-graph capture."""
    key = (spatial_shapes.data_ptr(), spatial_shapes._version, int(S), str(spatial_shapes.device))
    if key in _CHECKED:
        return
    if spatial_shapes.is_cuda and torch.cuda.is_current_stream_capturing():
        return
    total = int((spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum())
    assert total == S, f"spatial_shapes cover {total} positions but input_flatten has {S}"
    if len(_CHECKED) > 1024:
        _CHECKED.clear()
    _CHECKED.add(key)


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("MSDeformAttn: a power-of-two head dimension (32 in DINO) takes the vectorised "
                          "CUDA path; other sizes run the generic kernels.")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        taps = n_heads * n_levels * n_points
        self.sampling_offsets = nn.Linear(d_model, taps * 2)
        self.attention_weights = nn.Linear(d_model, taps)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        """Offsets start as a per-head compass direction scaled by the point index (1..P); attention
        logits start at zero (uniform softmax); projections are Xavier-uniform with zero bias."""
        M, L, P = self.n_heads, self.n_levels, self.n_points
        angle = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        direction = torch.stack([angle.cos(), angle.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]          # on the unit square
        steps = torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, P, 1)
        bias = direction.view(M, 1, 1, 2).repeat(1, L, P, 1) * steps
        nn.init.constant_(self.sampling_offsets.weight.data, 0.0)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(bias.reshape(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.0)
        nn.init.constant_(self.attention_weights.bias.data, 0.0)
        for proj in (self.value_proj, self.output_proj):
            nn.init.xavier_uniform_(proj.weight.data)
            nn.init.constant_(proj.bias.data, 0.0)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                input_level_start_index, input_padding_mask=None, residual=None, value_grad_chain=None):
        """query [N,Lq,C]; reference_points [N,Lq,L,2] (centres) or [N,Lq,L,4] (cx,cy,w,h boxes), in [0,1];
        input_flatten [N,S,C]; input_spatial_shapes [L,2]=(H,W); input_level_start_index [L];
        input_padding_mask [N,S] bool, True on padding.  Returns [N,Lq,C].
        `residual` (extension, optional [N,Lq,C]) is added to the result inside the output projection's epilogue.
        `value_grad_chain` (extension, optional linear.GradChain): the gradient of input_flatten through the value
        projection is summed across the modules sharing the chain (the decoder layers' cross-attention over one memory)."""
        N, Lq, _ = query.shape
        S = input_flatten.shape[1]
        M, L, P = self.n_heads, self.n_levels, self.n_points
        _check_levels(input_spatial_shapes, S)      # reference :92

        # value.masked_fill(mask[..., None], 0) of the reference (:96-97) is applied in place on the fresh projection
        # `residual is input_flatten` (encoder self-attention: the skip connection around the module starts at the tensor
        # the value projection reads): the output projection hands the skip gradient to the value projection's
        # input-gradient GEMM (linear.GradCarrier) instead of leaving autograd a separate accumulation pass
        carrier = None
        if (residual is not None and residual is input_flatten and input_flatten.requires_grad and torch.is_grad_enabled()
                and dl.get_mode() == "tf32" and dl.eligible(input_flatten, self.value_proj.weight)
                and dl.eligible(input_flatten, self.output_proj.weight) and _SKIP_CARRIER):
            carrier = dl.GradCarrier()
        vkw = dict(skip_in=carrier) if carrier is not None else {}
        if value_grad_chain is not None:        # extension: linear.GradChain shared by the modules reading input_flatten
            vkw["chain"] = value_grad_chain
        okw = dict(skip_out=carrier) if carrier is not None else {}
        value = dl.linear(input_flatten, self.value_proj.weight, self.value_proj.bias, zero_rows=input_padding_mask, **vkw)
        value = value.view(N, S, M, self.d_model // M)

        ref_dim = reference_points.shape[-1]
        if ref_dim not in (2, 4):
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {ref_dim} instead.")
        if _FUSED and query.dtype == torch.float32 and MSDA.fused_config_supported(value, reference_points, L, P):
            # the two projections of the query (sampling_offsets :99, attention_weights :100) as ONE GEMM over the
            # stacked weights; the kernels read offsets and logits as column slices of its output
            merged = dl.linear(query, torch.cat((self.sampling_offsets.weight, self.attention_weights.weight), 0),
                               torch.cat((self.sampling_offsets.bias, self.attention_weights.bias), 0))
            sampled = MSDeformAttnMergedFunction.apply(value, input_spatial_shapes, input_level_start_index, merged,
                                                       reference_points.contiguous(), M, L, P, _VALUE_STORAGE)
            return dl.linear(sampled, self.output_proj.weight, self.output_proj.bias, residual=residual, **okw)

        offsets = dl.linear(query, self.sampling_offsets.weight, self.sampling_offsets.bias).view(N, Lq, M, L, P, 2)
        logits = dl.linear(query, self.attention_weights.weight, self.attention_weights.bias).view(N, Lq, M, L * P)
        if _FUSED and MSDA.fused_supported(value, offsets, reference_points):
            sampled = MSDeformAttnFusedFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                                      offsets, logits, reference_points.contiguous())
            return dl.linear(sampled, self.output_proj.weight, self.output_proj.bias, residual=residual, **okw)

        weights = F.softmax(logits, -1).view(N, Lq, M, L, P)
        if ref_dim == 2:      # offsets are in pixels of each level: normalise by (W_l, H_l)
            wh = input_spatial_shapes.flip(-1)
            locations = reference_points[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
        elif ref_dim == 4:    # offsets are fractions of half the reference box, split over P points
            locations = reference_points[:, :, None, :, None, :2] \
                + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {ref_dim} instead.")

        if value.dtype == torch.float16:   # AMP: the op itself runs in fp32 (reference :114-121)
            sampled = MSDeformAttnFunction.apply(value.float(), input_spatial_shapes, input_level_start_index,
                                                 locations.float(), weights.float(), self.im2col_step)
            out = self.output_proj(sampled.to(torch.float16))
            return out if residual is None else out + residual
        sampled = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                             locations, weights, self.im2col_step)
        return dl.linear(sampled, self.output_proj.weight, self.output_proj.bias, residual=residual, **okw)
