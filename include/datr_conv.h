/*
 * datr_conv.h -- C ABI of the implicit-GEMM 3x3 convolution of libdatr_b200.so (sm_100a: 4-D TMA + tcgen05 + TMEM).
 *
 *   datr_conv3x3_nhwc_tf32  <-  conv2 (3x3, padding 1, stride 1 or 2, no bias) of the ResNet-50 bottlenecks the
 *                               reference's backbone runs through cuDNN (models/dino/backbone.py:97 ->
 *                               torchvision.models.resnet Bottleneck.conv2), followed by FrozenBatchNorm2d
 *                               (backbone.py:62-72; folded into `w` and `bias` by the caller) and ReLU;
 *                               the 3x3 convolutions + LeakyReLU(0.2) of the image-level domain discriminator
 *                               (models/dino/DA_utils.py:50-79, FCDiscriminator_img); and, on rotated / transposed
 *                               weights, the input gradient (dgrad) of any stride-1 3x3 convolution.
 *
 *   y[n, oy, ox, co] = act( bias[co] + sum_{ky,kx,ci} x[n, s*oy + ky - 1, s*ox + kx - 1, ci] * w[co, ky, kx, ci] )
 *
 * x [N,H,W,Cin], w [Cout,3,3,Cin], y [N,Ho,Wo,Cout] (Ho = (H-1)/s + 1, Wo likewise): fp32, contiguous (NHWC /
 * channels_last), 16-byte aligned, caller-owned device memory; bias [Cout] may be NULL; Cin % 32 == 0, Cout % 4 == 0.
 * `relu`: 0 = identity, 1 = ReLU, 2 = LeakyReLU(0.2).
 * TF32 products, fp32 accumulation.  Work is enqueued on `stream`; returns 0 or a negative code.
 * The weight gradient of these layers (and the input gradient of the stride-2 ones) is taken with the library
 * convolution-backward routines.
 */
#ifndef DATR_CONV_H_
#define DATR_CONV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DATR_CONV_OK = 0, DATR_CONV_ERR_BAD_ARGUMENT = -1, DATR_CONV_ERR_ALIGNMENT = -2, DATR_CONV_ERR_CUDA = -3 };

int datr_conv3x3_nhwc_tf32(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int Cin,
                           int Cout, int stride, int relu, void* stream);
/*
 * Weight (and bias) gradient of the same convolution -- the wgrad half of aten::convolution_backward (cuDNN) behind the
 * layers above:
 *   dw[co, ky, kx, ci] = sum_{n, oy, ox} gz[n, oy, ox, co] * x[n, s*oy + ky - 1, s*ox + kx - 1, ci]     db[co] = sum gz[..., co]
 * gz [N,Ho,Wo,Cout] (gradient of the pre-activation output), x [N,H,W,Cin], dw [Cout,3,3,Cin], db [Cout] or NULL; fp32,
 * NHWC, 16-byte aligned; Cin % 128 == 0, Cout % 4 == 0.  Runs on the tensor-core weight-gradient kernel of
 * datr_linear.h (csrc/wgrad_tf32.cu) with 4-D TMA boxes for the shifted input patches; dw / db are zero-filled by the
 * library on `stream`, partial tiles are reduced with vector atomics.  Returns 0 or a negative DATR_CONV_* code;
 * message: datr_linear_wgrad_last_error(); launches are counted by datr_linear_wgrad_launch_count().
 */
int datr_conv3x3_wgrad_nhwc_tf32(const float* gz, const float* x, float* dw, float* db, int N, int H, int W, int Cin,
                                 int Cout, int stride, void* stream);
const char* datr_conv_last_error(void);
uint64_t datr_conv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_CONV_H_ */
