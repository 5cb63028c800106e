// layernorm.cu -- LayerNorm over 256-channel rows for sm_100a, forward and backward (HBM-bound elementwise + reduction).
//
// Forward (layernorm256_fwd): one warp per row, the row lives in registers (two coalesced 512-byte loads), mean and
// the centred second moment by shuffle butterflies, y and the row statistics written in the same pass.  Algorithmic
// bytes 4*rows*256*2.  ATen's vectorized_layer_norm_kernel needs 76 us for the 44 446-row encoder activation on B200
// (gpurun_out/dino_step_ops.txt: 26 calls = 1.97 ms per training step), 5x the HBM time of the 91 MB it moves.
//
// The DINO transformer applies nn.LayerNorm(256) after every attention / FFN block (reference
// models/dino/deformable_transformer.py:801-820, :941-994, enc_output_norm :339) to [batch*tokens, 256] activations
// (44 446 rows at 1333x800, batch 2).  ATen's backward splits into a dx kernel and a gamma/beta kernel whose column
// reduction takes ~130 us per call on B200 (profiles/r01d_dino_step_kernels_graphs.txt: 9.0 ms per training step);
// this kernel does dx, dgamma and dbeta in ONE pass over dy and x:
//   * one warp per row, lane owns channels {4*lane..4*lane+3} and {128+4*lane..}: two coalesced 512-byte loads per tensor,
//   * row statistics via 5-step shuffle butterflies, dx written straight back,
//   * dgamma / dbeta (and, optionally, the column sum of dx = the bias gradient of the Linear that produced the
//     normalised tensor) accumulate in registers across the rows of a warp, are combined across the CTA in shared
//     memory and leave as one atomicAdd per channel per CTA.
// Algorithmic bytes per launch: 4*rows*256*3 (dy, x read; dx written) + O(rows) statistics.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_layernorm.h"

namespace {

thread_local char g_ln_err[256] = "";
std::atomic<uint64_t> g_ln_launches{0};

int lnfail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_ln_err, sizeof g_ln_err, fmt, detail);
  return code;
}

constexpr int kC = 256;
constexpr int kWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(kWarps * 32)
layernorm256_fwd(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = lane * 4, c1 = 128 + lane * 4;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c1));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c1));
  const float inv_c = 1.0f / kC;
  const long long stride = (long long)gridDim.x * kWarps;
  long long r = (long long)blockIdx.x * kWarps + warp;
  if (r >= rows) return;
  float4 x0 = __ldg(reinterpret_cast<const float4*>(x + r * kC + c0));
  float4 x1 = __ldg(reinterpret_cast<const float4*>(x + r * kC + c1));
  while (true) {
    const long long rn = r + stride;
    float4 n0 = x0, n1 = x1;
    if (rn < rows) {   // next row's loads in flight while this row is reduced
      n0 = __ldg(reinterpret_cast<const float4*>(x + rn * kC + c0));
      n1 = __ldg(reinterpret_cast<const float4*>(x + rn * kC + c1));
    }
    const float mu = warp_sum(((x0.x + x0.y) + (x0.z + x0.w)) + ((x1.x + x1.y) + (x1.z + x1.w))) * inv_c;
    const float d[8] = {x0.x - mu, x0.y - mu, x0.z - mu, x0.w - mu, x1.x - mu, x1.y - mu, x1.z - mu, x1.w - mu};
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) q = fmaf(d[i], d[i], q);
    const float rs = rsqrtf(warp_sum(q) * inv_c + eps);
    float* yr = y + r * kC;
    *reinterpret_cast<float4*>(yr + c0) = make_float4(fmaf(d[0] * rs, g0.x, b0.x), fmaf(d[1] * rs, g0.y, b0.y),
                                                      fmaf(d[2] * rs, g0.z, b0.z), fmaf(d[3] * rs, g0.w, b0.w));
    *reinterpret_cast<float4*>(yr + c1) = make_float4(fmaf(d[4] * rs, g1.x, b1.x), fmaf(d[5] * rs, g1.y, b1.y),
                                                      fmaf(d[6] * rs, g1.z, b1.z), fmaf(d[7] * rs, g1.w, b1.w));
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
    if (rn >= rows) break;
    r = rn; x0 = n0; x1 = n1;
  }
}

__global__ void __launch_bounds__(kWarps * 32)
layernorm256_bwd(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                 float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx_colsum, int rows) {
  __shared__ float red[3][kWarps][kC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = lane * 4, c1 = 128 + lane * 4;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c1));
  float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ac[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float inv_c = 1.0f / kC;
  for (long long r = (long long)blockIdx.x * kWarps + warp; r < rows; r += (long long)gridDim.x * kWarps) {
    const float* xr = x + r * kC;
    const float* dr = dy + r * kC;
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(xr + c0)), x1 = __ldg(reinterpret_cast<const float4*>(xr + c1));
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(dr + c0)), d1 = __ldg(reinterpret_cast<const float4*>(dr + c1));
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    const float xh[8] = {(x0.x - mu) * rs, (x0.y - mu) * rs, (x0.z - mu) * rs, (x0.w - mu) * rs,
                         (x1.x - mu) * rs, (x1.y - mu) * rs, (x1.z - mu) * rs, (x1.w - mu) * rs};
    const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float s1 = 0.f, s2 = 0.f, w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      w[i] = dv[i] * gv[i];
      s1 += w[i];
      s2 = fmaf(w[i], xh[i], s2);
      ag[i] = fmaf(dv[i], xh[i], ag[i]);
      ab[i] += dv[i];
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = rs * (w[i] - s1 - xh[i] * s2);
      ac[i] += o[i];
    }
    float* oxr = dx + r * kC;
    *reinterpret_cast<float4*>(oxr + c0) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(oxr + c1) = make_float4(o[4], o[5], o[6], o[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (i < 4 ? c0 : c1 - 4) + i;
    red[0][warp][c] = ag[i];
    red[1][warp][c] = ab[i];
    red[2][warp][c] = ac[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kC; c += blockDim.x) {
    float a = 0.f, b = 0.f, s = 0.f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) { a += red[0][k][c]; b += red[1][k][c]; s += red[2][k][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
    if (dx_colsum) atomicAdd(dx_colsum + c, s);
  }
}

}  // namespace

extern "C" {

int datr_layernorm256_backward(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                               float* dx, float* dgamma, float* dbeta, float* dx_colsum, int rows, void* stream_) {
  if (!dy || !x || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta) return lnfail(DATR_LN_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (rows <= 0) return lnfail(DATR_LN_ERR_BAD_ARGUMENT, "rows must be positive%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(dy) || !al16(x) || !al16(gamma) || !al16(dx)) return lnfail(DATR_LN_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t e = cudaMemsetAsync(dgamma, 0, kC * sizeof(float), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, kC * sizeof(float), stream);
  if (e == cudaSuccess && dx_colsum) e = cudaMemsetAsync(dx_colsum, 0, kC * sizeof(float), stream);
  if (e != cudaSuccess) return lnfail(DATR_LN_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  const long long want = ((long long)rows + kWarps - 1) / kWarps;
  const int grid = int(want < 148 * 4 ? want : 148 * 4);     // 4 CTAs (32 warps) per SM, grid-stride over rows
  layernorm256_bwd<<<grid, kWarps * 32, 0, stream>>>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, dx_colsum, rows);
  e = cudaGetLastError();
  if (e != cudaSuccess) return lnfail(DATR_LN_ERR_CUDA, "layernorm256_bwd launch: %s", cudaGetErrorString(e));
  g_ln_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_LN_OK;
}

int datr_layernorm256_forward(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean,
                              float* rstd, int rows, void* stream_) {
  if (!x || !gamma || !beta || !y || !mean || !rstd) return lnfail(DATR_LN_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (rows <= 0) return lnfail(DATR_LN_ERR_BAD_ARGUMENT, "rows must be positive%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(x) || !al16(gamma) || !al16(beta) || !al16(y)) return lnfail(DATR_LN_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long want = ((long long)rows + kWarps - 1) / kWarps;
  const int grid = int(want < 148 * 8 ? want : 148 * 8);     // 8 CTAs (64 warps) per SM, grid-stride over rows
  layernorm256_fwd<<<grid, kWarps * 32, 0, stream>>>(x, gamma, beta, eps, y, mean, rstd, rows);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return lnfail(DATR_LN_ERR_CUDA, "layernorm256_fwd launch: %s", cudaGetErrorString(e));
  g_ln_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_LN_OK;
}

const char* datr_layernorm_last_error(void) { return g_ln_err; }
uint64_t datr_layernorm_launch_count(void) { return g_ln_launches.load(std::memory_order_relaxed); }

}  // extern "C"
