// attn_softmax.cu -- masked softmax over attention-score rows and its backward, in place (sm_100a, HBM-bound).
//
// Decoder self-attention of DINO (reference models/dino/deformable_transformer.py:880-897): 900 matching + up to 200
// de-noising queries, 8 heads x 32 channels, boolean mask.  PyTorch's memory-efficient SDPA kernel (fp32, sm80 code)
// needs 143 us forward / 343 us backward per layer on B200 for 2.5 / 6 GFLOP; with the 77 MB score matrix simply kept
// in HBM the same attention is two library batched GEMMs around ONE pass of this softmax (and three GEMMs + one pass
// in the backward).  One warp owns a row: the row lives in registers (coalesced 128-byte loads, element j = 32*c + lane),
// max / sum / dot by shuffle butterflies, result written back over the input.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_attn.h"

namespace {

thread_local char g_at_err[256] = "";
std::atomic<uint64_t> g_at_launches{0};

int afail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_at_err, sizeof g_at_err, fmt, detail);
  return code;
}

constexpr int kWarps = 8;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

template <int PER, bool kMask>   // PER * 32 >= T
__global__ void __launch_bounds__(kWarps * 32)
softmax_fwd(float* __restrict__ s, const uint8_t* __restrict__ blocked, float scale, long long rows, int T, int Tq) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* row = s + r * T;
  const uint8_t* mrow = kMask ? blocked + (r % Tq) * (long long)T : nullptr;
  // all loads of the row first (clamped index instead of a branch per element: 36 independent loads in flight),
  // then the arithmetic
  float v[PER];
  uint8_t b[PER];
#pragma unroll
  for (int c = 0; c < PER; ++c) v[c] = row[min(c * 32 + lane, T - 1)];
#pragma unroll
  for (int c = 0; c < PER; ++c) b[c] = kMask ? mrow[min(c * 32 + lane, T - 1)] : uint8_t(0);
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int c = 0; c < PER; ++c) {
    const bool live = c * 32 + lane < T && !b[c];
    v[c] = live ? v[c] * scale : -CUDART_INF_F;
    mx = fmaxf(mx, v[c]);
  }
  mx = wmax(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < PER; ++c) {
    v[c] = (c * 32 + lane < T) ? expf(v[c] - mx) : 0.f;    // a fully blocked row gives NaN, like torch
    sum += v[c];
  }
  const float inv = 1.0f / wsum(sum);
#pragma unroll
  for (int c = 0; c < PER; ++c) {
    const int j = c * 32 + lane;
    if (j < T) row[j] = v[c] * inv;
  }
}

template <int PER>
__global__ void __launch_bounds__(kWarps * 32)
softmax_bwd(const float* __restrict__ p, float* __restrict__ dp, float scale, long long rows, int T) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* prow = p + r * T;
  float* drow = dp + r * T;
  float pv[PER], dv[PER];
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < PER; ++c) {
    const int j = c * 32 + lane;
    pv[c] = dv[c] = 0.f;
    if (j < T) { pv[c] = prow[j]; dv[c] = drow[j]; }
    dot = fmaf(pv[c], dv[c], dot);
  }
  dot = wsum(dot);
#pragma unroll
  for (int c = 0; c < PER; ++c) {
    const int j = c * 32 + lane;
    if (j < T) drow[j] = scale * (pv[c] * (dv[c] - dot));
  }
}

}  // namespace

extern "C" {

int datr_attn_softmax_forward(float* s, const uint8_t* blocked, float scale, long long rows, int T, int Tq, void* stream_) {
  if (!s) return afail(DATR_ATTN_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (rows <= 0 || T <= 0 || T > 2048 || Tq <= 0) return afail(DATR_ATTN_ERR_BAD_ARGUMENT, "need rows > 0, 0 < T <= 2048, Tq > 0%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned grid = unsigned((rows + kWarps - 1) / kWarps);
  if (T <= 36 * 32) {
    if (blocked) softmax_fwd<36, true><<<grid, kWarps * 32, 0, stream>>>(s, blocked, scale, rows, T, Tq);
    else softmax_fwd<36, false><<<grid, kWarps * 32, 0, stream>>>(s, blocked, scale, rows, T, Tq);
  } else {
    if (blocked) softmax_fwd<64, true><<<grid, kWarps * 32, 0, stream>>>(s, blocked, scale, rows, T, Tq);
    else softmax_fwd<64, false><<<grid, kWarps * 32, 0, stream>>>(s, blocked, scale, rows, T, Tq);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return afail(DATR_ATTN_ERR_CUDA, "softmax_fwd launch: %s", cudaGetErrorString(e));
  g_at_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_ATTN_OK;
}

int datr_attn_softmax_backward(const float* p, float* dp, float scale, long long rows, int T, void* stream_) {
  if (!p || !dp) return afail(DATR_ATTN_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (rows <= 0 || T <= 0 || T > 2048) return afail(DATR_ATTN_ERR_BAD_ARGUMENT, "need rows > 0, 0 < T <= 2048%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned grid = unsigned((rows + kWarps - 1) / kWarps);
  if (T <= 36 * 32) softmax_bwd<36><<<grid, kWarps * 32, 0, stream>>>(p, dp, scale, rows, T);
  else softmax_bwd<64><<<grid, kWarps * 32, 0, stream>>>(p, dp, scale, rows, T);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return afail(DATR_ATTN_ERR_CUDA, "softmax_bwd launch: %s", cudaGetErrorString(e));
  g_at_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_ATTN_OK;
}

const char* datr_attn_last_error(void) { return g_at_err; }
uint64_t datr_attn_launch_count(void) { return g_at_launches.load(std::memory_order_relaxed); }

}  // extern "C"
