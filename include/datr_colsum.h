/*
 * datr_colsum.h -- C ABI of the bias-gradient kernels of libdatr_b200.so (sm_100a).
 *
 *   datr_colsum           <- grad_bias = grad_output.sum(0) of every nn.Linear on the DINO transformer path
 *                            (autograd of torch.nn.functional.linear at models/dino/ops/modules/ms_deform_attn.py:94-125,
 *                            models/dino/deformable_transformer.py:784-805, :941-947)
 *   datr_relu_bwd_colsum  <- the ReLU backward that precedes it for the FFN's linear1
 *                            (deformable_transformer.py:803: linear2(dropout(activation(linear1(src)))))
 *
 * x / dy / y / dz are [rows, cols] fp32 row-major, 16-byte aligned, caller-owned device memory, cols % 4 == 0;
 * out / db are [cols], zero-filled by the library on `stream` and accumulated with atomics.  Returns 0 or a
 * negative code; datr_colsum_last_error() gives the calling thread's message.
 */
#ifndef DATR_COLSUM_H_
#define DATR_COLSUM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DATR_CS_OK = 0, DATR_CS_ERR_BAD_ARGUMENT = -1, DATR_CS_ERR_ALIGNMENT = -2, DATR_CS_ERR_CUDA = -3 };

int datr_colsum(const float* x, float* out, int rows, int cols, void* stream);
int datr_relu_bwd_colsum(const float* dy, const float* y, float* dz, float* db, int rows, int cols, void* stream);
const char* datr_colsum_last_error(void);
uint64_t datr_colsum_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_COLSUM_H_ */
