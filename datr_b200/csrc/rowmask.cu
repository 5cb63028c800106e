// rowmask.cu -- zero the padding rows of a token matrix in place (sm_100a).
//
// The reference masks the projected value map with `masked_fill(padding_mask[..., None], 0)`
// (models/dino/ops/modules/ms_deform_attn.py:96-97), an out-of-place elementwise pass over the whole [N*S, 256]
// activation in the forward and again over its gradient in the backward.  Only the masked rows change, so this kernel
// reads the one-byte-per-row mask (a warp ballots 32 rows at a time) and writes zeros to the masked rows only.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_rowmask.h"

namespace {

thread_local char g_rm_err[256] = "";
std::atomic<uint64_t> g_rm_launches{0};

int rmfail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_rm_err, sizeof g_rm_err, fmt, detail);
  return code;
}

__global__ void __launch_bounds__(256)
zero_masked_rows(float* __restrict__ x, const uint8_t* __restrict__ mask, long long rows, int cols) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int vec = cols >> 2;
  for (long long r0 = warp * 32; r0 < rows; r0 += nwarps * 32) {
    const long long r = r0 + lane;
    unsigned hit = __ballot_sync(0xffffffffu, r < rows && mask[r] != 0);
    while (hit) {
      const int k = __ffs(hit) - 1;
      hit &= hit - 1;
      float4* row = reinterpret_cast<float4*>(x + (r0 + k) * cols);
      for (int c = lane; c < vec; c += 32) row[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

}  // namespace

extern "C" {

int datr_zero_masked_rows(float* x, const uint8_t* mask, long long rows, int cols, void* stream_) {
  if (!x || !mask) return rmfail(DATR_RM_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (rows <= 0 || cols <= 0 || (cols & 3)) return rmfail(DATR_RM_ERR_BAD_ARGUMENT, "rows must be positive and cols a positive multiple of 4%s");
  if (reinterpret_cast<uintptr_t>(x) & 15) return rmfail(DATR_RM_ERR_ALIGNMENT, "x must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long want = (rows + 255) / 256;            // 8 warps x 32 rows per CTA
  const int grid = int(want < 148 * 8 ? want : 148 * 8);
  zero_masked_rows<<<grid, 256, 0, stream>>>(x, mask, rows, cols);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return rmfail(DATR_RM_ERR_CUDA, "zero_masked_rows launch: %s", cudaGetErrorString(e));
  g_rm_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_RM_OK;
}

const char* datr_rowmask_last_error(void) { return g_rm_err; }
uint64_t datr_rowmask_launch_count(void) { return g_rm_launches.load(std::memory_order_relaxed); }

}  // extern "C"
