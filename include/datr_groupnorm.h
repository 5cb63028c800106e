/*
 * datr_groupnorm.h -- C ABI of GroupNorm on NHWC activations (libdatr_b200.so, sm_100a).
 *
 * Replaces nn.GroupNorm(32, hidden_dim) of the reference's input projections (models/dino/dino.py:111-126: Conv2d + GroupNorm
 * per feature level) and its autograd on the channels_last feature maps of this package -- ATen's CUDA GroupNorm is an NCHW
 * kernel, so it cost a layout copy in and out in both directions.
 *
 *   x, y, dy, dx  [N, HW, C] fp32 (NHWC flattened), gamma / beta / dgamma / dbeta [C], mean / rstd [N, G] (outputs of the
 *   forward, inputs of the backward), scratch: 2 * N * G doubles (device memory, used inside the call).
 *   forward : y = (x - mean[n, g]) * rstd[n, g] * gamma[c] + beta[c],  statistics over the HW * C/G elements of a group
 *             (biased variance, rstd = 1 / sqrt(var + eps)), sums accumulated in fp64.
 *   backward: dx, dgamma, dbeta of that expression (dgamma / dbeta zero-filled by the library).
 * Shapes: C % G == 0, (C / G) % 4 == 0, 256 % (C / 4) == 0 (C = 256, G = 32 in DINO).  Algorithmic bytes: forward 12 per element,
 * backward 20.  Returns 0, -1 (bad argument) or -3 (CUDA error).
 */
#ifndef DATR_GROUPNORM_H_
#define DATR_GROUPNORM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int datr_groupnorm_nhwc_forward(const float* x, const float* gamma, const float* beta, int N, long long HW, int C, int G, float eps,
                                float* y, float* mean, float* rstd, double* scratch, void* stream);
int datr_groupnorm_nhwc_backward(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int N,
                                 long long HW, int C, int G, float* dx, float* dgamma, float* dbeta, double* scratch, void* stream);
const char* datr_groupnorm_last_error(void);
uint64_t datr_groupnorm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_GROUPNORM_H_ */
