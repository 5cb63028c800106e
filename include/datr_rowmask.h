/*
 * datr_rowmask.h -- C ABI of the padding-mask kernel of libdatr_b200.so (sm_100a).
 *
 *   datr_zero_masked_rows  <-  `value = value.masked_fill(input_padding_mask[..., None], float(0))` in the reference's
 *                              MSDeformAttn.forward (models/dino/ops/modules/ms_deform_attn.py:96-97) and its autograd
 *                              backward (the same fill applied to the incoming gradient).
 *
 * x [rows, cols] fp32 (contiguous, device memory, caller-owned, 16-byte aligned, cols % 4 == 0) is modified IN PLACE:
 * every row r with mask[r] != 0 (one byte per row, torch.bool layout) is set to zero; other rows are not touched, so
 * the kernel moves `rows` mask bytes plus the masked rows only (the ATen op rewrites the whole tensor: 27 us per call
 * at the encoder's [2, 22223, 256] activation on B200, 51 calls per training step).  Returns 0 or a negative code.
 */
#ifndef DATR_ROWMASK_H_
#define DATR_ROWMASK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DATR_RM_OK = 0, DATR_RM_ERR_BAD_ARGUMENT = -1, DATR_RM_ERR_ALIGNMENT = -2, DATR_RM_ERR_CUDA = -3 };

int datr_zero_masked_rows(float* x, const uint8_t* mask, long long rows, int cols, void* stream);

const char* datr_rowmask_last_error(void);
uint64_t datr_rowmask_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_ROWMASK_H_ */
