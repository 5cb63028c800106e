// adamw.cu -- gradient clipping + AdamW over every parameter of the model in ONE launch (sm_100a, HBM-bound).
//
// The reference's step ends with torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1) and torch.optim.AdamW.step()
// (engine.py:108-111, main.py:152): a scaling pass over all gradients plus the optimizer's multi-tensor kernels (torch's
// fused AdamW: 10 launches, 0.68 ms for DINO's 47.8 M parameters on B200 -- five times the time the 765 MB of traffic
// need).  Here a CTA owns one chunk of one parameter (tables in device memory, as in ema.cu); the clipping coefficient
// is read from device memory (no host sync) and applied on the fly, so the separate scaling pass disappears.
// Arithmetic follows torch's fused kernel (ATen/native/cuda/fused_adam_utils.cuh, adam_math, ADAMW mode, no amsgrad,
// no maximize) operation by operation in fp32.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_adamw.h"

namespace {

thread_local char g_aw_err[256] = "";
std::atomic<uint64_t> g_aw_launches{0};

struct Hyper { float beta1, beta2, eps, step_scale /* 1 / bias_correction1 */, bc2_sqrt; };

__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, float lr, float wd, const Hyper& h) {
  p -= lr * wd * p;
  m = m + (1.f - h.beta1) * (g - m);                       // lerp(exp_avg, grad, 1 - beta1), weight < 0.5
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p -= (lr * h.step_scale) * m / denom;
}

// segs: per parameter {param*, grad*, exp_avg*, exp_avg_sq*, numel, lr bits | weight-decay bits << 32}
__global__ void __launch_bounds__(256)
adamw_step(const int64_t* __restrict__ segs, const int64_t* __restrict__ chunks, const float* __restrict__ grad_scale,
           Hyper h) {
  const int64_t seg = chunks[2 * blockIdx.x], first = chunks[2 * blockIdx.x + 1];
  const int64_t* s = segs + 6 * seg;
  float* p = reinterpret_cast<float*>(s[0]) + first;
  const float* g = reinterpret_cast<const float*>(s[1]) + first;
  float* m = reinterpret_cast<float*>(s[2]) + first;
  float* v = reinterpret_cast<float*>(s[3]) + first;
  int64_t n = s[4] - first;
  if (n > DATR_ADAMW_CHUNK) n = DATR_ADAMW_CHUNK;
  const float lr = __int_as_float(int(uint64_t(s[5]) & 0xffffffffu)), wd = __int_as_float(int(uint64_t(s[5]) >> 32));
  const float gs = grad_scale ? __ldg(grad_scale) : 1.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t nv = vec ? (n >> 2) : 0;
  for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    adamw1(pp.x, gg.x * gs, mm.x, vv.x, lr, wd, h); adamw1(pp.y, gg.y * gs, mm.y, vv.y, lr, wd, h);
    adamw1(pp.z, gg.z * gs, mm.z, vv.z, lr, wd, h); adamw1(pp.w, gg.w * gs, mm.w, vv.w, lr, wd, h);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) adamw1(p[i], __ldg(g + i) * gs, m[i], v[i], lr, wd, h);
}

}  // namespace

extern "C" {

int datr_adamw_step(const int64_t* segs, const int64_t* chunks, int n_chunks, const float* grad_scale, float beta1, float beta2,
                    float eps, float bias_correction1, float bias_correction2_sqrt, void* stream_) {
  if (!segs || !chunks || n_chunks <= 0 || !(bias_correction1 > 0.f) || !(bias_correction2_sqrt > 0.f)) {
    snprintf(g_aw_err, sizeof g_aw_err, "datr_adamw_step: null table, n_chunks <= 0 or non-positive bias correction");
    return -1;
  }
  const Hyper h = {beta1, beta2, eps, 1.f / bias_correction1, bias_correction2_sqrt};
  adamw_step<<<unsigned(n_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(segs, chunks, grad_scale, h);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_aw_err, sizeof g_aw_err, "adamw_step launch: %s", cudaGetErrorString(e)); return -3; }
  g_aw_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

const char* datr_adamw_last_error(void) { return g_aw_err; }
uint64_t datr_adamw_launch_count(void) { return g_aw_launches.load(std::memory_order_relaxed); }

}  // extern "C"
