"""3x3 convolution + FrozenBN + ReLU on NHWC activations through the implicit-GEMM tcgen05 kernel
(host side of include/datr_conv.h; csrc/conv3x3_tf32.cu).

`conv3x3_bias_act(x, weight, bias, stride, act)`: x [N,Cin,H,W] and weight [Cout,Cin,3,3] in channels_last memory
format (i.e. NHWC / [Cout,3,3,Cin] in memory), bias [Cout]; returns act(conv(x, weight, padding=1) + bias) as a
channels_last tensor, act = identity / ReLU / LeakyReLU(0.2).  Forward = our kernel (TF32 products, fp32 accumulation).
Backward: activation mask; the input gradient of stride-1 layers is the SAME kernel on the rotated, channel-swapped
filter; weight / bias gradients run on the tensor-core weight-gradient kernel (include/datr_conv.h,
datr_conv3x3_wgrad_nhwc_tf32; Cin % 128 == 0); only the input gradient of stride-2 layers (and weight gradients of
layers with fewer input channels) still come from ATen's convolution_backward (cuDNN).
Used in "tf32" mode by the ResNet bottleneck's conv2 (reference models/dino/backbone.py:97; FrozenBN folded into weight /
bias by the caller), the image-level domain discriminator (DA_utils.py:50-79) and the extra stride-2 input projection
(dino.py:118-123)."""
from __future__ import annotations

import os

import torch

from . import fallbacks, native


# Which engine runs what -- decided by measurement on B200 (tools/bench_conv_backward.py,
# profiles/r02n_bench_conv_backward.txt, 4-image 1333x800 batch, cold L2):
#   forward  : this kernel wins on large maps (>= 40 000 output pixels: ResNet layer2 66 vs 80 us, 58 vs 79 us; the first
#              and third discriminator layers at level 0, 137 vs 181 us, 58 vs 72 us) and loses to cuDNN's 2-SM
#              Blackwell kernels on the small ones (layer3 / layer4 72-80 vs 54-64 us; the 2048 -> 256 extra level 256 vs
#              45 us), so `use_kernel` routes by output size;
#   backward : cuDNN's dgrad / wgrad are 1.3-2.5x faster than the own input-gradient (forward kernel on the rotated
#              filter: 69 vs 40 us at layer2) and weight-gradient (4-D TMA patches on the linear wgrad kernel: 144 vs 64 us)
#              paths at every shape of the step, so the library is the default; DATR_CONV_BACKWARD=own selects the own
#              kernels (parity-tested in tests/test_conv_gpu.py).
MIN_OUTPUT_PIXELS = 40000
OWN_BACKWARD = os.environ.get("DATR_CONV_BACKWARD", "lib") == "own"


def _launch(xc, wc, bias, stride, act):
    n, cin, h, w = xc.shape
    cout = wc.shape[0]
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    y = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=xc.device, memory_format=torch.channels_last)
    lib = native.lib()
    with torch.cuda.device(xc.device):
        rc = lib.datr_conv3x3_nhwc_tf32(xc.data_ptr(), wc.data_ptr(), bias.data_ptr() if bias is not None else None,
                                        y.data_ptr(), n, h, w, cin, cout, stride, act,
                                        torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"datr_conv3x3_nhwc_tf32 failed (code {rc}): {lib.datr_conv_last_error().decode()}")
    return y


class _Conv3x3(torch.autograd.Function):
    """act(conv3x3(x, weight, padding=1, stride) + bias), act: 0 identity, 1 ReLU, 2 LeakyReLU(0.2)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, act):
        xc = x.contiguous(memory_format=torch.channels_last)
        wc = weight.contiguous(memory_format=torch.channels_last)
        y = _launch(xc, wc, bias, stride, act)
        ctx.stride, ctx.act = stride, act
        ctx.save_for_backward(xc, wc, y if act else None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        gz = gy.contiguous(memory_format=torch.channels_last)
        if ctx.act == 1:
            gz = torch.ops.aten.threshold_backward(gz, y, 0.0)
        elif ctx.act == 2:      # LeakyReLU(0.2) keeps the sign: the mask can be read from the output
            gz = torch.where(y > 0, gz, gz * 0.2)
        cout, cin = w.shape[0], w.shape[1]
        gx = gw = None
        own_dgrad = OWN_BACKWARD and ctx.needs_input_grad[0] and ctx.stride == 1 and cout % 32 == 0 and cin % 4 == 0
        if own_dgrad:
            # input gradient of a stride-1 convolution = the same convolution of gz with the filter rotated by 180 degrees
            # and its channel axes swapped: the forward kernel on a [Cin, 3, 3, Cout] copy of the (small) weight
            w_rot = w.flip(2, 3).transpose(0, 1).contiguous(memory_format=torch.channels_last)
            gx = _launch(gz, w_rot, None, 1, 0)
        gb = None
        own_wgrad = OWN_BACKWARD and ctx.needs_input_grad[1] and cin % 128 == 0 and cout % 4 == 0
        if own_wgrad:
            # weight (+ bias) gradient on the tensor-core weight-gradient kernel, shifted input patches by 4-D TMA boxes
            gw = torch.empty_like(w)                                  # channels_last: [Cout, 3, 3, Cin] in memory
            gb = torch.empty(cout, dtype=torch.float32, device=w.device) if ctx.needs_input_grad[2] else None
            n, _, h, wd = x.shape
            lib = native.lib()
            with torch.cuda.device(x.device):
                rc = lib.datr_conv3x3_wgrad_nhwc_tf32(gz.data_ptr(), x.data_ptr(), gw.data_ptr(), gb.data_ptr() if gb is not None else None,
                                                      n, h, wd, cin, cout, ctx.stride, torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise RuntimeError(f"datr_conv3x3_wgrad_nhwc_tf32 failed (code {rc}): {lib.datr_linear_wgrad_last_error().decode()}")
        need = [ctx.needs_input_grad[0] and not own_dgrad, ctx.needs_input_grad[1] and not own_wgrad, False]
        if need[0] or need[1]:
            fallbacks.note("aten.convolution_backward (cuDNN " + " + ".join(n for n, k in (("dgrad", need[0]), ("wgrad", need[1])) if k)
                           + ") of a 3x3 convolution")
            gx2, gw2, _ = torch.ops.aten.convolution_backward(gz, x, w, None, [ctx.stride, ctx.stride], [1, 1], [1, 1], False,
                                                              [0, 0], 1, need)
            gx = gx2 if need[0] else gx
            gw = gw2 if need[1] else gw
        if gb is None and ctx.needs_input_grad[2]:
            gb = gz.sum((0, 2, 3))
        return gx, gw, gb, None, None


def eligible(x: torch.Tensor, conv: torch.nn.Conv2d) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and conv.kernel_size == (3, 3) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.stride in ((1, 1), (2, 2))
            and conv.in_channels % 32 == 0 and conv.out_channels % 4 == 0)


def use_kernel(x: torch.Tensor, conv: torch.nn.Conv2d) -> bool:
    """eligible() and large enough for the own forward kernel to beat the library (see the table above)."""
    if not eligible(x, conv):
        return False
    s = conv.stride[0]
    return x.shape[0] * ((x.shape[2] - 1) // s + 1) * ((x.shape[3] - 1) // s + 1) >= MIN_OUTPUT_PIXELS


def conv3x3_bias_relu(x, weight, bias, stride: int):
    return _Conv3x3.apply(x, weight, bias, stride, 1)


def conv3x3_bias_act(x, weight, bias, stride: int, act: int):
    """act: 0 identity, 1 ReLU, 2 LeakyReLU(0.2)."""
    return _Conv3x3.apply(x, weight, bias, stride, act)
