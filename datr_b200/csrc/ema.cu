// ema.cu -- exponential moving average of a whole state dict in one launch (sm_100a, HBM-bound).
//
// The reference's teacher update (models/dino/EMA.py:41-50 and siblings, util/utils.py:391-392) walks the state dict
// in Python: two elementwise kernels per tensor, ~1 280 launches for DINO-4scale.  Here a CTA owns one chunk of one
// tensor (tables in device memory, see include/datr_ema.h); 16-byte vector accesses where the chunk is aligned.
// The arithmetic keeps the reference's roundings: fl(fl(ema * d) + fl((1 - d) * model)), so results are bit-identical.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_ema.h"

namespace {

thread_local char g_ema_err[256] = "";
std::atomic<uint64_t> g_ema_launches{0};

int efail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_ema_err, sizeof g_ema_err, fmt, detail);
  return code;
}

__device__ __forceinline__ float ema1(float e, float m, float d, float omd) {
  return __fadd_rn(__fmul_rn(e, d), __fmul_rn(omd, m));     // no FMA: the reference rounds the two products
}

__global__ void __launch_bounds__(256)
ema_update(const int64_t* __restrict__ segs, const int64_t* __restrict__ chunks, float d, float omd) {
  const int64_t seg = chunks[2 * blockIdx.x], first = chunks[2 * blockIdx.x + 1];
  float* e = reinterpret_cast<float*>(segs[3 * seg]) + first;
  const float* m = reinterpret_cast<const float*>(segs[3 * seg + 1]) + first;
  int64_t n = segs[3 * seg + 2] - first;
  if (n > DATR_EMA_CHUNK) n = DATR_EMA_CHUNK;
  if (((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(m)) & 15) == 0) {
    const int64_t nv = n >> 2;
    for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
      float4 a = reinterpret_cast<float4*>(e)[i];
      const float4 b = __ldg(reinterpret_cast<const float4*>(m) + i);
      a.x = ema1(a.x, b.x, d, omd); a.y = ema1(a.y, b.y, d, omd); a.z = ema1(a.z, b.z, d, omd); a.w = ema1(a.w, b.w, d, omd);
      reinterpret_cast<float4*>(e)[i] = a;
    }
    for (int64_t i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) e[i] = ema1(e[i], __ldg(m + i), d, omd);
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) e[i] = ema1(e[i], __ldg(m + i), d, omd);
  }
}

}  // namespace

extern "C" {

int datr_ema_update(const int64_t* segs, const int64_t* chunks, int n_chunks, float decay, float one_minus_decay,
                    void* stream_) {
  if (!segs || !chunks) return efail(DATR_EMA_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (n_chunks <= 0) return efail(DATR_EMA_ERR_BAD_ARGUMENT, "n_chunks must be positive%s");
  ema_update<<<unsigned(n_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(segs, chunks, decay, one_minus_decay);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return efail(DATR_EMA_ERR_CUDA, "ema_update launch: %s", cudaGetErrorString(e));
  g_ema_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_EMA_OK;
}

const char* datr_ema_last_error(void) { return g_ema_err; }
uint64_t datr_ema_launch_count(void) { return g_ema_launches.load(std::memory_order_relaxed); }

}  // extern "C"
