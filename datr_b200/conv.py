"""3x3 convolution + FrozenBN + ReLU on NHWC activations through the implicit-GEMM tcgen05 kernel
(host side of include/datr_conv.h; csrc/conv3x3_tf32.cu).

`conv3x3_bias_relu(x, weight, bias, stride)`: x [N,Cin,H,W] and weight [Cout,Cin,3,3] in channels_last memory
format (i.e. NHWC / [Cout,3,3,Cin] in memory), bias [Cout]; returns relu(conv(x, weight, padding=1) + bias) as a
channels_last tensor.  Forward = our kernel (TF32 products, fp32 accumulation); backward = ReLU mask + ATen's
convolution_backward (cuDNN dgrad / wgrad) on the same operands.  Used by the ResNet bottleneck's conv2 in "tf32" mode
(reference models/dino/backbone.py:97); FrozenBN is folded into weight / bias by the caller."""
from __future__ import annotations

import torch

from . import fallbacks, native


class _Conv3x3BiasReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        n, cin, h, w = x.shape
        cout = weight.shape[0]
        xc = x.contiguous(memory_format=torch.channels_last)
        wc = weight.contiguous(memory_format=torch.channels_last)
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        y = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        lib = native.lib()
        with torch.cuda.device(x.device):
            rc = lib.datr_conv3x3_nhwc_tf32(xc.data_ptr(), wc.data_ptr(), bias.data_ptr() if bias is not None else None,
                                            y.data_ptr(), n, h, w, cin, cout, stride, 1,
                                            torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"datr_conv3x3_nhwc_tf32 failed (code {rc}): {lib.datr_conv_last_error().decode()}")
        ctx.stride = stride
        ctx.save_for_backward(xc, wc, y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        gz = torch.ops.aten.threshold_backward(gy.contiguous(memory_format=torch.channels_last), y, 0.0)
        fallbacks.note("aten.convolution_backward (cuDNN dgrad + wgrad) of a 3x3 ResNet convolution")
        need = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False]
        gx, gw, _ = torch.ops.aten.convolution_backward(gz, x, w, None, [ctx.stride, ctx.stride], [1, 1], [1, 1], False,
                                                        [0, 0], 1, need)
        gb = gz.sum((0, 2, 3)) if ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


def eligible(x: torch.Tensor, conv: torch.nn.Conv2d) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and conv.kernel_size == (3, 3) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None and conv.stride in ((1, 1), (2, 2))
            and conv.in_channels % 32 == 0 and conv.out_channels % 4 == 0)


def conv3x3_bias_relu(x, weight, bias, stride: int):
    return _Conv3x3BiasReLU.apply(x, weight, bias, stride)
