/*
 * datr_attn.h -- C ABI of the decoder self-attention softmax kernels of libdatr_b200.so (sm_100a).
 *
 * The reference's decoder layer runs nn.MultiheadAttention over the 900 matching + <=200 de-noising queries with a
 * boolean attention mask (models/dino/deformable_transformer.py:880-897, mask built in dn_components.py:105-121).
 * At T <= 1100 tokens, 8 heads x 32 channels the score matrix of a layer is 77 MB, so the attention is run as
 *   S = Q K^T (library batched GEMM)  ->  P = softmax(scale * S + mask)  ->  O = P V (library batched GEMM)
 * with these two HBM-bound kernels for the softmax and its backward, both IN PLACE on the score matrix:
 *
 *   datr_attn_softmax_forward   s[rows, T] <- softmax_j(scale * s[r, j]  with  -inf where blocked[(r % Tq) * T + j])
 *                               (`blocked` = the reference's bool attn_mask [Tq, T], True = may NOT attend; NULL = none;
 *                               rows = batch * heads * Tq).  Replaces the scale / masked_fill / softmax chain of
 *                               torch.nn.functional.multi_head_attention_forward.
 *   datr_attn_softmax_backward  dp[rows, T] <- scale * p * (dp - sum_j dp[r, j] * p[r, j])   (dS from dP, in place)
 *
 * fp32, contiguous, caller-owned device buffers, T <= 2048.  Algorithmic bytes: forward 8 * rows * T (+ mask),
 * backward 12 * rows * T.  Returns 0 or a negative code.
 */
#ifndef DATR_ATTN_H_
#define DATR_ATTN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DATR_ATTN_OK = 0, DATR_ATTN_ERR_BAD_ARGUMENT = -1, DATR_ATTN_ERR_CUDA = -3 };

int datr_attn_softmax_forward(float* s, const uint8_t* blocked, float scale, long long rows, int T, int Tq, void* stream);
int datr_attn_softmax_backward(const float* p, float* dp, float scale, long long rows, int T, void* stream);

/*
 * Fused decoder self-attention (csrc/attn_fused.cu): S = Q K^T -> masked softmax -> O = P V in one tcgen05 kernel, the score
 * tile kept in tensor memory.  q / k / v are row-strided views of the in-projection outputs: element (n, t, h, c) of q lives
 * at q[(n * T + t) * q_row_stride + h * 32 + c] (32 channels per head; with nn.MultiheadAttention's packed in_proj output
 * [N, T, 2C]: q = base, k = base + C, both with row stride 2C).  Replaces torch.nn.functional.multi_head_attention_forward's
 * bmm / softmax / bmm chain (reference models/dino/deformable_transformer.py:900-908).
 *   datr_attn_mask_words(T)   32-bit words per row of the packed mask (4 per 128-key tile)
 *   datr_attn_pack_mask       blocked [T, T] bytes (True = may NOT attend; NULL = no mask) -> bits [T, words] and, for the
 *                             backward, bits_t [T, words] (the mask transposed: bits along the queries)
 *   datr_attn_fused_forward   out [N, T, H*32];  lse [N, H, T] = log sum_j exp(scale * s_ij) over the attendable keys (nullable);
 *                             p_out [N*H, T, T] = the probabilities (nullable: only the GEMM-based backward needs them)
 * TF32 products, fp32 accumulation / softmax.
 */
int datr_attn_mask_words(int T);
int datr_attn_pack_mask(const uint8_t* blocked, int T, uint32_t* bits, uint32_t* bits_t /* nullable: transposed mask */, void* stream);
int datr_attn_fused_forward(const float* q, long long q_row_stride, const float* k, long long k_row_stride, const float* v,
                            long long v_row_stride, const uint32_t* mask_bits, int N, int H, int T, float scale, float* out,
                            float* lse, float* p_out, void* stream);
/*
 *   datr_attn_fused_backward  dq / dk / dv (row-strided like q / k / v) from dout [N, T, H*32], the forward's out and lse; two
 *                             launches (dQ + delta, then dK + dV), score tiles recomputed in tensor memory; `mask_bits_t` =
 *                             the transposed packed mask (second output of datr_attn_pack_mask); `delta` [N, H, T] scratch.
 */
int datr_attn_fused_backward(const float* q, long long q_row_stride, const float* k, long long k_row_stride, const float* v,
                             long long v_row_stride, const uint32_t* mask_bits, const uint32_t* mask_bits_t, int N, int H,
                             int T, float scale, const float* out, const float* lse, const float* dout, float* delta,
                             float* dq, long long dq_row_stride, float* dk, long long dk_row_stride, float* dv,
                             long long dv_row_stride, void* stream);
const char* datr_attn_fused_last_error(void);
uint64_t datr_attn_fused_launch_count(void);

const char* datr_attn_last_error(void);
uint64_t datr_attn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_ATTN_H_ */
