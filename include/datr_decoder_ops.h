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

/*
 *   datr_pos_embed_hw   PositionEmbeddingSineHW (reference models/dino/position_encoding.py:62-107) after its cumulative
 *                       sums: y, x [rows] = the (normalised, scaled) cumulative coordinates of every pixel, dim_t_h / dim_t_w
 *                       [feats] = temperature ** (2 * (i // 2) / feats); out [rows, 2 * feats] = [sin / cos interleaved of
 *                       y / dim_t_h | of x / dim_t_w]  (= pos.permute(0, 2, 3, 1) of the reference's [N, 2 * feats, H, W]).
 *   datr_bn_relu_maxpool_nhwc   tail of the frozen ResNet stem (reference backbone.py:97 / torchvision bn1 -> relu -> maxpool):
 *                       out [N, Ho, Wo, C] = max over the 3x3 / stride 2 / padding 1 window of relu(x * scale + shift),
 *                       x [N, H, W, C] NHWC, C % 4 == 0, Ho = (H - 1) / 2 + 1.  No gradient (the stem never trains).
 */
int datr_pos_embed_hw(const float* y, const float* x, const float* dim_t_h, const float* dim_t_w, long long rows, int feats,
                      float* out, void* stream);
int datr_bn_relu_maxpool_nhwc(const float* x, const float* scale, const float* shift, int N, int H, int W, int C, float* out,
                              void* stream);
const char* datr_decoder_ops_last_error(void);
uint64_t datr_decoder_ops_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_DECODER_OPS_H_ */
