/*
 * datr_msda.h -- C ABI of the B200 (sm_100a) multi-scale deformable attention
 * library (libdatr_b200.so).  Plain pointers and sizes only; no torch types.
 *
 * Each entry point replaces one function of the reference's native extension
 * `MultiScaleDeformableAttention` (paths relative to the reference repo):
 *
 *   datr_msda_forward   <- ms_deform_attn_forward   models/dino/ops/src/vision.cpp:14,
 *                          ms_deform_attn.h:21-39 (dispatch),
 *                          cuda/ms_deform_attn_cuda.cu:20-80 (host),
 *                          cuda/ms_deform_im2col_cuda.cuh:237-299, :924-954 (kernel + launcher)
 *   datr_msda_backward  <- ms_deform_attn_backward  vision.cpp:15, ms_deform_attn.h:41-60,
 *                          ms_deform_attn_cuda.cu:83-153, ms_deform_im2col_cuda.cuh:87-159,
 *                          :301-920 (6 kernel variants), :957-1327 (launcher)
 *
 * Contract (differences from the reference are deliberate and listed):
 *   - The CALLER owns every buffer (device memory, contiguous, row-major).  The
 *     library allocates nothing and keeps no global mutable state; calls are
 *     re-entrant across host threads and streams.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*; NULL =
 *     legacy default stream) and the call returns without host synchronisation,
 *     like the reference (cu:65,135 use the current torch stream).
 *   - forward: every element of `output` is written (no pre-zeroing needed; the
 *     reference pre-zeroes with at::zeros, cu:54).
 *   - backward: `grad_value` is zero-filled BY THE LIBRARY on `stream` before the
 *     scatter; `grad_sampling_loc` and `grad_attn_weight` are fully overwritten
 *     (the reference allocates the three with zeros, cu:121-123).
 *   - spatial_shapes / level_start_index are int64 arrays IN DEVICE MEMORY, as in
 *     the reference (cuh:240-241).
 *   - Errors are returned as negative codes (the reference only printf()s launch
 *     errors, cuh:948-952); datr_last_error() gives the message for the calling
 *     thread.  The im2col_step batch chunking of the reference host code
 *     (cu:50-75) has no numerical effect and is validated by the Python shim.
 *   - There is no CPU implementation (the reference's CPU entry points throw,
 *     cpu/ms_deform_attn_cpu.cpp:26,39).
 *
 * Shapes:  value [batch, spatial_size, num_heads, channels]
 *          spatial_shapes int64 [num_levels, 2] = (H_l, W_l); level_start_index int64 [num_levels]
 *          sampling_loc [batch, num_query, num_heads, num_levels, num_point, 2] = (x, y) in [0,1]
 *          attn_weight  [batch, num_query, num_heads, num_levels, num_point]
 *          output / grad_output [batch, num_query, num_heads * channels]
 */
#ifndef DATR_MSDA_H_
#define DATR_MSDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element type of value / sampling_loc / attn_weight / output (all the same, as cu:64 dispatches) */
enum { DATR_DTYPE_F32 = 0, DATR_DTYPE_F64 = 1 };

enum {
  DATR_OK = 0,
  DATR_ERR_BAD_ARGUMENT = -1, /* null pointer, non-positive dimension, unknown dtype */
  DATR_ERR_ALIGNMENT = -2,    /* a buffer is not aligned to its element type */
  DATR_ERR_CUDA = -3,         /* memset / kernel launch failed; see datr_last_error() */
  DATR_ERR_UNSUPPORTED = -4   /* fused entry points only: configuration outside fp32 / 32 channels / L*P <= 32 */
};

int datr_msda_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const void* sampling_loc, const void* attn_weight,
                      int batch, int spatial_size, int num_heads, int channels,
                      int num_levels, int num_query, int num_point, int dtype,
                      void* output, void* stream);

int datr_msda_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const void* sampling_loc, const void* attn_weight, const void* grad_output,
                       int batch, int spatial_size, int num_heads, int channels,
                       int num_levels, int num_query, int num_point, int dtype,
                       void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* stream);

/*
 * Module-level fusion: the elementwise prologue of the reference's MSDeformAttn.forward
 * (models/dino/ops/modules/ms_deform_attn.py:99-111) runs inside the kernels, so sampling_locations and
 * attention_weights are never materialised.  Inputs are what that prologue consumes:
 *   sampling_offsets [batch, num_query, num_heads, num_levels, num_point, 2]   output of the offsets Linear (:99)
 *   attn_logits      [batch, num_query, num_heads, num_levels * num_point]     output of the weights Linear (:100)
 *   reference_points [batch, num_query, num_levels, ref_dim], ref_dim = 2 or 4 (:102-111)
 *     ref_dim 2: loc = ref + offsets / (W_l, H_l);   ref_dim 4: loc = ref.xy + offsets / num_point * ref.wh * 0.5
 *   attention weights = softmax of the logits over num_levels * num_point (:101)
 * sampling_offsets and attn_logits may be row-strided views (`*_row_stride` = elements between consecutive
 * queries, 0 = densely packed; everything inside one query's row is contiguous), so both can live in ONE GEMM output
 * [batch * num_query, 3 * heads * levels * points]; the two gradients are written with the same strides and form one
 * GEMM operand for the merged Linear's backward.
 * backward writes grad_value (zero-filled by the library), grad_sampling_offsets and grad_attn_logits (softmax and
 * location chain rules applied); reference_points receive no gradient (they are detached in the DINO transformer).
 * Supported: DATR_DTYPE_F32, channels == 32, num_point in {1,2,4,8}, num_levels * num_point <= 32; anything else
 * returns DATR_ERR_UNSUPPORTED and the caller composes datr_msda_forward / _backward with the torch prologue.
 */
int datr_msda_fused_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                            const void* sampling_offsets, long long offsets_row_stride,
                            const void* attn_logits, long long logits_row_stride, const void* reference_points,
                            int ref_dim, int batch, int spatial_size, int num_heads, int channels,
                            int num_levels, int num_query, int num_point, int dtype, void* output, void* stream);

int datr_msda_fused_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                             const void* sampling_offsets, long long offsets_row_stride,
                             const void* attn_logits, long long logits_row_stride, const void* reference_points,
                             int ref_dim, const void* grad_output, int batch, int spatial_size, int num_heads,
                             int channels, int num_levels, int num_query, int num_point, int dtype,
                             void* grad_value, void* grad_sampling_offsets, void* grad_attn_logits, void* stream);

/*
 * The same two backward entry points with HOST copies of spatial_shapes / level_start_index next to the device arrays
 * (either both or neither; NULL = not available).  With them the library can build one TMA tensor map per level over
 * grad_value and scatter each sample's 2x2-pixel block with ONE `cp.reduce.async.bulk.tensor` instead of four vector
 * reductions through the load/store unit (4 points, <= 4 levels of at least 2x2 pixels; selected by
 * datr_msda_set_backward_stages, off by default); results are the same sums in a different order of additions.  The plain entry points above are these with NULL host arrays.
 */
int datr_msda_backward_hs(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const int64_t* host_spatial_shapes, const int64_t* host_level_start_index,
                          const void* sampling_loc, const void* attn_weight, const void* grad_output,
                          int batch, int spatial_size, int num_heads, int channels,
                          int num_levels, int num_query, int num_point, int dtype,
                          void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* stream);

int datr_msda_fused_backward_hs(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                const int64_t* host_spatial_shapes, const int64_t* host_level_start_index,
                                const void* sampling_offsets, long long offsets_row_stride,
                                const void* attn_logits, long long logits_row_stride, const void* reference_points,
                                int ref_dim, const void* grad_output, int batch, int spatial_size, int num_heads,
                                int channels, int num_levels, int num_query, int num_point, int dtype,
                                void* grad_value, void* grad_sampling_offsets, void* grad_attn_logits, void* stream);

/*
 * Pair-row value maps (B200 layout, no counterpart in the reference).  The forward gather is bound by the number of
 * 128-byte lines the L1TEX pipe serves (4 per sample with fp32 rows), not by bytes.  datr_msda_pack_value_pairs rewrites
 * value [batch, spatial_size, heads, 32] fp32 as 16-bit PAIR rows of the same size: the line of (pixel, head) holds, for
 * each of the 8 lanes that read it, 4 channels of the pixel followed by the same 4 channels of the pixel one column to
 * the right (zeros in the last column of a level), so a bilinear sample needs TWO lines.  datr_msda_fused_forward_pairs is
 * datr_msda_fused_forward on such a map (fp32 accumulation and output; 32 channels, 4 points).  Precision class: bf16
 * storage is inside BASELINE's 1e-2 bar, fp16 storage (saturating) inside 1e-3 for |value| < 65504.
 */
enum { DATR_STORE_F32 = 0, DATR_STORE_BF16_PAIRS = 1, DATR_STORE_FP16_PAIRS = 2 };

int datr_msda_pack_value_pairs(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                               int batch, int spatial_size, int num_heads, int num_levels, int storage, void* pairs,
                               void* stream);

int datr_msda_fused_forward_pairs(const void* pairs, int storage, const int64_t* spatial_shapes,
                                  const int64_t* level_start_index, const void* sampling_offsets,
                                  long long offsets_row_stride, const void* attn_logits, long long logits_row_stride,
                                  const void* reference_points, int ref_dim, int batch, int spatial_size, int num_heads,
                                  int channels, int num_levels, int num_query, int num_point, void* output, void* stream);

/* Scatter variant of the fast backward: 0 = vector reductions, 1 / 2 = TMA reduce with that many staging buffers per
 * warp (default 0 = the faster one on B200, see csrc/msda.cu; environment DATR_MSDA_BWD_STAGES at load time).  Process-wide. */
void datr_msda_set_backward_stages(int stages);
int datr_msda_get_backward_stages(void);

/* Message of the last failing call made by the calling thread ("" if none). */
const char* datr_last_error(void);

/* ABI version of this header (bumped on any signature change). */
int datr_abi_version(void);

/* Number of kernel launches issued by this process through the library (for bench accounting). */
uint64_t datr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_MSDA_H_ */
