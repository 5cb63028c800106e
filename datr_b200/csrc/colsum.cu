// colsum.cu -- bias gradients of the Linear layers for sm_100a: column sums of tall [rows, cols] fp32 matrices, and
// the ReLU-backward mask fused with the column sum (HBM-bound: one pass over the data).
//
//   datr_colsum            out[c]  = sum_r x[r, c]                       (db of every nn.Linear: rows = 44 446 tokens)
//   datr_relu_bwd_colsum   dz[r,c] = y[r,c] > 0 ? dy[r,c] : 0;  db[c] = sum_r dz[r,c]      (FFN linear1, 2048 columns)
//
// ATen's generic reduce kernel needs ~25 us per call for these shapes on B200 (7.3 ms per DINO training step,
// profiles/r01d_dino_step_kernels_graphs.txt); here a CTA owns a slab of rows, each thread keeps a float4 of
// columns in registers across the slab (coalesced 512-byte warp loads), row-lanes are combined in shared memory and
// each CTA leaves with one atomicAdd per column.  Algorithmic bytes: 4*rows*cols (colsum), 12*rows*cols (relu_bwd).
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_colsum.h"

namespace {

thread_local char g_cs_err[256] = "";
std::atomic<uint64_t> g_cs_launches{0};

int csfail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_cs_err, sizeof g_cs_err, fmt, detail);
  return code;
}

constexpr int kThreads = 256;

// blockDim.x = 256 threads = cg column groups (float4 each) x (256 / cg) row lanes; grid = (row slabs, column blocks)
template <bool kRelu>
__global__ void __launch_bounds__(kThreads)
colsum_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ dz, float* __restrict__ out,
              int rows, int cols, int cg) {
  __shared__ float4 red[kThreads];
  const int lanes = kThreads / cg;
  const int cgi = threadIdx.x % cg, rl = threadIdx.x / cg;
  const int c4 = blockIdx.y * cg + cgi;                 // float4 column index
  const bool live = c4 * 4 < cols;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const size_t stride = (size_t)cols;
    const long long step = (long long)gridDim.x * lanes;
    constexpr int kU = 4;                                   // independent loads in flight per thread
    long long r = (long long)blockIdx.x * lanes + rl;
    for (; r + (kU - 1) * step < rows; r += kU * step) {
      float4 v[kU], a[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const size_t off = (size_t)(r + u * step) * stride + (size_t)c4 * 4;
        v[u] = __ldg(reinterpret_cast<const float4*>(x + off));
        if (kRelu) a[u] = __ldg(reinterpret_cast<const float4*>(y + off));
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (kRelu) {
          v[u].x = a[u].x > 0.f ? v[u].x : 0.f; v[u].y = a[u].y > 0.f ? v[u].y : 0.f;
          v[u].z = a[u].z > 0.f ? v[u].z : 0.f; v[u].w = a[u].w > 0.f ? v[u].w : 0.f;
          *reinterpret_cast<float4*>(dz + (size_t)(r + u * step) * stride + (size_t)c4 * 4) = v[u];
        }
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
    for (; r < rows; r += step) {
      const size_t off = (size_t)r * stride + (size_t)c4 * 4;
      float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
      if (kRelu) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(y + off));
        v.x = a.x > 0.f ? v.x : 0.f; v.y = a.y > 0.f ? v.y : 0.f; v.z = a.z > 0.f ? v.z : 0.f; v.w = a.w > 0.f ? v.w : 0.f;
        *reinterpret_cast<float4*>(dz + off) = v;
      }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0 && live) {
    for (int k = 1; k < lanes; ++k) {
      const float4 o = red[k * cg + cgi];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    float* o = out + (size_t)c4 * 4;
    atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
  }
}

int launch(bool relu, const float* x, const float* y, float* dz, float* out, int rows, int cols, void* stream_) {
  if (!x || !out || (relu && (!y || !dz))) return csfail(DATR_CS_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (rows <= 0 || cols <= 0 || cols % 4 != 0) return csfail(DATR_CS_ERR_BAD_ARGUMENT, "rows > 0 and cols %% 4 == 0 required%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(x) || !al16(out) || (relu && (!al16(y) || !al16(dz)))) return csfail(DATR_CS_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, stream);
  if (e != cudaSuccess) return csfail(DATR_CS_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  const int c4 = cols / 4;
  int cg = 1;
  while (cg < c4 && cg < kThreads) cg <<= 1;            // power of two <= 256 covering the float4 columns
  const int lanes = kThreads / cg;
  const int gy = (c4 + cg - 1) / cg;
  long long gx = (148LL * 3 + gy - 1) / gy;             // ~3 CTAs per SM: few atomics per column, long row sweeps
  const long long max_gx = ((long long)rows + lanes - 1) / lanes;
  if (gx > max_gx) gx = max_gx;
  if (gx < 1) gx = 1;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  if (relu) colsum_kernel<true><<<grid, kThreads, 0, stream>>>(x, y, dz, out, rows, cols, cg);
  else colsum_kernel<false><<<grid, kThreads, 0, stream>>>(x, nullptr, nullptr, out, rows, cols, cg);
  e = cudaGetLastError();
  if (e != cudaSuccess) return csfail(DATR_CS_ERR_CUDA, "colsum_kernel launch: %s", cudaGetErrorString(e));
  g_cs_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_CS_OK;
}

}  // namespace

extern "C" {

int datr_colsum(const float* x, float* out, int rows, int cols, void* stream) {
  return launch(false, x, nullptr, nullptr, out, rows, cols, stream);
}

int datr_relu_bwd_colsum(const float* dy, const float* y, float* dz, float* db, int rows, int cols, void* stream) {
  return launch(true, dy, y, dz, db, rows, cols, stream);
}

const char* datr_colsum_last_error(void) { return g_cs_err; }
uint64_t datr_colsum_launch_count(void) { return g_cs_launches.load(std::memory_order_relaxed); }

}  // extern "C"
