/*
 * datr_layernorm.h -- C ABI of the LayerNorm(256) kernels of libdatr_b200.so (sm_100a).
 *
 *   datr_layernorm256_forward   <-  torch.nn.LayerNorm(256).forward at the same call sites (ATen native_layer_norm):
 *                                   y = (x - mean) * rstd * gamma + beta, rstd = 1/sqrt(var + eps) with the biased
 *                                   variance; also writes the row statistics mean, rstd [rows] the backward reads.
 *   datr_layernorm256_backward  <-  the backward of torch.nn.LayerNorm(256) as used by the reference's transformer
 *                                   (models/dino/deformable_transformer.py:801-820 norm1/norm2, :941-994 norm1-3,
 *                                   :339 enc_output_norm, decoder norm): ATen's native_layer_norm_backward.
 *
 * Given dy, x [rows, 256], gamma [256] and the forward's statistics mean, rstd [rows] (fp32, contiguous, device
 * memory, caller-owned, 16-byte aligned) it writes
 *   dx [rows, 256], dgamma [256] = sum_rows dy * xhat, dbeta [256] = sum_rows dy,
 *   and, if dx_colsum != NULL, dx_colsum [256] = sum_rows dx (the bias gradient of the layer feeding the norm).
 * dgamma / dbeta / dx_colsum are zero-filled by the library on `stream` and accumulated with atomics (summation
 * order, hence the last bits, vary from run to run).  Returns 0 or a negative code.
 */
#ifndef DATR_LAYERNORM_H_
#define DATR_LAYERNORM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DATR_LN_OK = 0, DATR_LN_ERR_BAD_ARGUMENT = -1, DATR_LN_ERR_ALIGNMENT = -2, DATR_LN_ERR_CUDA = -3 };

int datr_layernorm256_forward(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean,
                              float* rstd, int rows, void* stream);

int datr_layernorm256_backward(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                               float* dx, float* dgamma, float* dbeta, float* dx_colsum, int rows, void* stream);

const char* datr_layernorm_last_error(void);
uint64_t datr_layernorm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_LAYERNORM_H_ */
