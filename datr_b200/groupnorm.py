"""GroupNorm on NHWC activations (host side of include/datr_groupnorm.h).

`group_norm_nhwc(module, x)` == module(x) for an nn.GroupNorm `module` and a CUDA fp32 tensor x [N, C, H, W] that is
contiguous in channels_last; the result is channels_last too.  Replaces ATen's NCHW GroupNorm (+ two layout copies per
direction) behind the reference's input projections (models/dino/dino.py:111-126)."""
from __future__ import annotations

import torch

from . import native


def applicable(module, x) -> bool:
    C, G = module.num_channels, module.num_groups
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
            and module.affine and C % G == 0 and (C // G) % 4 == 0 and 256 % (C // 4) == 0 and x.shape[1] == C)


class _GroupNormNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps):
        N, C, H, W = x.shape
        lib = native.lib()
        with torch.cuda.device(x.device):
            y = torch.empty_like(x)                       # channels_last like x
            mean = torch.empty((N, groups), dtype=torch.float32, device=x.device)
            rstd = torch.empty_like(mean)
            scratch = torch.empty(2 * N * groups, dtype=torch.float64, device=x.device)
            rc = lib.datr_groupnorm_nhwc_forward(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), N, H * W, C, groups, float(eps),
                                                 y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), scratch.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"datr_groupnorm_nhwc_forward failed (code {rc}): {lib.datr_groupnorm_last_error().decode()}")
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.groups = groups
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, gamma, mean, rstd = ctx.saved_tensors
        N, C, H, W = x.shape
        gy = gy.contiguous(memory_format=torch.channels_last)
        lib = native.lib()
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x)
            dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
            dbeta = torch.empty_like(dgamma)
            scratch = torch.empty(2 * N * ctx.groups, dtype=torch.float64, device=x.device)
            rc = lib.datr_groupnorm_nhwc_backward(gy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                                  N, H * W, C, ctx.groups, gx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                                  scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"datr_groupnorm_nhwc_backward failed (code {rc}): {lib.datr_groupnorm_last_error().decode()}")
        return gx, dgamma, dbeta, None, None


def group_norm_nhwc(module, x):
    return _GroupNormNHWC.apply(x, module.weight, module.bias, module.num_groups, module.eps)
