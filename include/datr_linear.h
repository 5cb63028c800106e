/*
 * datr_linear.h -- C ABI of the tensor-core linear layer of libdatr_b200.so (sm_100a: TMA + tcgen05 + TMEM).
 *
 *   datr_linear_tf32  <-  every nn.Linear on the DINO transformer hot path, i.e. torch.nn.functional.linear as
 *                         called from the reference's models/dino/ops/modules/ms_deform_attn.py:94-125
 *                         (value_proj, sampling_offsets, attention_weights, output_proj) and
 *                         models/dino/deformable_transformer.py:784-805, :941-947 (linear1 + activation, linear2),
 *                         with the bias add, the ReLU and the residual add that follow it fused into the epilogue;
 *                         and the 1x1 convolutions of the ResNet-50 bottlenecks on NHWC activations
 *                         (models/dino/backbone.py:97 -> torchvision Bottleneck conv1 / conv3 / downsample) with the
 *                         FrozenBatchNorm2d affine (backbone.py:62-72) folded into weight and bias.
 *
 *   relu == 0:  y[M,N] =        x[M,K] . w[N,K]^T + bias[N]   + residual[M,N]
 *   relu == 1:  y[M,N] = max(0, x[M,K] . w[N,K]^T + bias[N] ) + residual[M,N]      (FFN linear1)
 *   relu == 2:  y[M,N] = max(0, x[M,K] . w[N,K]^T + bias[N]   + residual[M,N] )    (ResNet bottleneck output)
 *   relu == 3:  y[M,N] = residual[M,N] > 0 ? x[M,K] . w[N,K]^T + bias[N] : 0            (input gradient through a ReLU:
 *                                                                                        `residual` is the saved activation)
 *
 * All buffers are fp32 device memory, row-major, contiguous, 16-byte aligned, owned by the caller; `bias` and
 * `residual` may be NULL.  Requirements: K % 32 == 0, N % 4 == 0.
 * Products are TF32 (10-bit mantissa, round-to-nearest on the TMA load), accumulation is fp32.  Work is enqueued on
 * `stream` (cudaStream_t as void*), no host synchronisation.  Returns 0 or a negative code;
 * datr_linear_last_error() gives the calling thread's message.  There is no CPU implementation.
 */
#ifndef DATR_LINEAR_H_
#define DATR_LINEAR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  DATR_LINEAR_OK = 0,
  DATR_LINEAR_ERR_BAD_ARGUMENT = -1,
  DATR_LINEAR_ERR_ALIGNMENT = -2,
  DATR_LINEAR_ERR_CUDA = -3
};

int datr_linear_tf32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                     int M, int N, int K, int relu, void* stream);

/*
 * Weight / bias gradient of the same layer:  dw[N,K] = dz[M,N]^T . x[M,K],  db[N] = sum_m dz[m,n]  (db may be NULL).
 * Replaces the `grad_output.t().mm(input)` and `grad_output.sum(0)` of torch.nn.functional.linear's autograd.
 * Both operands are read as they lie in memory (transposed-operand MMA); the M rows are split over CTAs and the
 * partial tiles are reduced with vector atomics into dw / db, which the library zero-fills on `stream` first
 * (summation order, hence the last bits, vary from run to run).  N % 4 == 0, K % 4 == 0.
 */
int datr_linear_wgrad_tf32(const float* dz, const float* x, float* dw, float* db, int M, int N, int K, void* stream);
/* dw += dz^T x, db += column sums of dz: no zero fill -- the reductions land in a buffer that already holds gradient
 * (what autograd's AccumulateGrad does with `param.grad += new` after `grad_output.t().mm(input)` during the reference's
 * `optimizer.zero_grad(); losses.backward()`, engine.py:98-99 -- here in the weight-gradient kernel itself; dw / db are the
 * parameter's slices of the step's flat gradient buffer, zeroed once per step by the caller). */
int datr_linear_wgrad_tf32_acc(const float* dz, const float* x, float* dw, float* db, int M, int N, int K, void* stream);
const char* datr_linear_wgrad_last_error(void);
uint64_t datr_linear_wgrad_launch_count(void);

/*
 * bf16-operand variants (the "bf16" precision class of BASELINE.json, 1e-2): x / w / dz as bf16 in HBM, fp32 accumulation in
 * tensor memory, tcgen05.mma.kind::f16.  Same tiling and epilogues as the TF32 kernels; half the operand bytes per FLOP.
 *   datr_linear_bf16        y = act(x w^T + bias) (+ residual);  y fp32 or bf16 (y_bf16); residual fp32 [M, N], or bf16
 *                           (residual_bf16) when it is the ReLU-mask source of relu == 3.  K % 64 == 0, N % 4 == 0.
 *   datr_linear_wgrad_bf16  dw [N, K] = dz^T x, db [N] = column sums of dz (fp32 outputs, zero-filled by the library).
 */
/* datr_linear_tf32 with the weight given transposed, w_t [K, N] row-major: y = act(x w_t + bias) + residual.  The input
 * gradient of a Linear (dx = dy W, W stored [N_fwd, K_fwd]) without materialising W^T. */
int datr_linear_tf32_bt(const float* x, const float* w_t, const float* bias, const float* residual, float* y,
                        int M, int N, int K, int relu, void* stream);

/* y = (x w_t + residual) where mask > 0, else 0; mask [M, N] fp32, residual [M, N] optional.  Replaces, in the backward of
 * a ResNet bottleneck (reference models/dino/backbone.py:62-72 is torchvision's Bottleneck: out = relu(bn3(conv3(..)) +
 * identity)), three passes autograd runs separately: the input gradient of conv1, its sum with the gradient of the identity
 * branch, and the ReLU backward of the previous block's output. */
int datr_linear_tf32_bt_masked(const float* x, const float* w_t, const float* residual, const float* mask, float* y,
                               int M, int N, int K, void* stream);

int datr_linear_bf16(const void* x, const void* w, const float* bias, const void* residual, int residual_bf16, void* y,
                     int y_bf16, int M, int N, int K, int relu, void* stream);
int datr_linear_wgrad_bf16(const void* dz, const void* x, float* dw, float* db, int M, int N, int K, void* stream);
int datr_linear_wgrad_bf16_acc(const void* dz, const void* x, float* dw, float* db, int M, int N, int K, void* stream);

const char* datr_linear_last_error(void);

/* Number of tensor-core linear kernels launched by this process (bench accounting). */
uint64_t datr_linear_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_LINEAR_H_ */
