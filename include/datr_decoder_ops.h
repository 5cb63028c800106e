/*
 * datr_decoder_ops.h -- C ABI of the small fused kernels of the DINO decoder layer loop (libdatr_b200.so, sm_100a).
 *
 *   datr_sine_embed   the sine position embedding of the decoder's reference boxes, replacing
 *                     gen_sineembed_for_position (reference models/dino/utils.py, called per decoder layer from
 *                     models/dino/deformable_transformer.py TransformerDecoder.forward):
 *                       pos [rows, k] (k = 2: x, y; k = 4: x, y, w, h; normalised), dim_t [128] = 10000 ** (2 * (i // 2) / 128)
 *                       out [rows, 128 * k], blocks ordered (y, x[, w, h]); feature 2j = sin(pos * 2 pi / dim_t[2j]),
 *                       feature 2j + 1 = cos(pos * 2 pi / dim_t[2j + 1]).
 *                     fp32, contiguous, caller-owned device buffers; returns 0 or a negative code.
 */
#ifndef DATR_DECODER_OPS_H_
#define DATR_DECODER_OPS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int datr_sine_embed(const float* pos, const float* dim_t, long long rows, int k, float* out, void* stream);
const char* datr_decoder_ops_last_error(void);
uint64_t datr_decoder_ops_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_DECODER_OPS_H_ */
