// groupnorm.cu -- GroupNorm over NHWC activations, forward and backward (sm_100a, HBM-bound).
//
// The reference's input projections are Conv2d(1x1 | 3x3 s2) + nn.GroupNorm(32, 256) (models/dino/dino.py:111-126).  ATen's
// CUDA GroupNorm works on NCHW: on the channels_last feature maps of this package it cost a layout copy in, the
// moments + apply kernels, and a layout copy back (and the same again in the backward).  Here the statistics and the
// normalisation run on the NHWC tensor the projection GEMM wrote:
//   forward : stats pass (per (image, group) sum and sum of squares, fp32 per thread, fp64 atomics per CTA) -> finalize
//             (mean, rstd) -> apply  y = (x - mean) * rstd * gamma + beta            : 2 reads + 1 write of the map
//   backward: sums pass (per (image, group) sum dy*gamma and sum dy*gamma*x; per channel dgamma, dbeta) -> apply
//             dx = rstd * (dy * gamma - c2 - xhat * c1),  c1 = mean_g(dy*gamma*xhat), c2 = mean_g(dy*gamma)
// A thread owns 4 consecutive channels of a pixel (channels per group must be a multiple of 4, so a float4 never
// straddles groups); a CTA of 256 threads covers C/4 lanes x (1024/C) pixel rows and walks a slab of pixels.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_groupnorm.h"

namespace {

thread_local char g_gn_err[256] = "";
std::atomic<uint64_t> g_gn_launches{0};

constexpr int kThreads = 256;
constexpr int kSlab = 128;      // pixels per CTA

__device__ __forceinline__ float sum4(const float4& v) { return (v.x + v.y) + (v.z + v.w); }

// stats[n][g] = {sum, sum of squares} (fp64, zeroed by the caller)
__global__ void __launch_bounds__(kThreads)
gn_stats(const float* __restrict__ x, long long HW, int C, int G, double* __restrict__ stats) {
  extern __shared__ float sh[];                     // [2][rows][C/4]
  const int lanes = C / 4, rows = kThreads / lanes;
  const int c4 = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * kSlab, p1 = min(p0 + kSlab, HW);
  const float4* xp = reinterpret_cast<const float4*>(x) + ((long long)n * HW) * lanes + c4;
  float s = 0.f, ss = 0.f;
  if (pr < rows)
    for (long long p = p0 + pr; p < p1; p += rows) {
      const float4 v = __ldg(xp + p * lanes);
      s += sum4(v);
      ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  sh[pr * lanes + c4] = s;
  sh[(rows + pr) * lanes + c4] = ss;
  __syncthreads();
  const int lpg = (C / G) / 4;                      // float4 lanes per group
  if (threadIdx.x < G) {
    double a = 0.0, b = 0.0;
    for (int r = 0; r < rows; ++r)
      for (int l = 0; l < lpg; ++l) {
        a += double(sh[r * lanes + threadIdx.x * lpg + l]);
        b += double(sh[(rows + r) * lanes + threadIdx.x * lpg + l]);
      }
    atomicAdd(stats + ((long long)n * G + threadIdx.x) * 2, a);
    atomicAdd(stats + ((long long)n * G + threadIdx.x) * 2 + 1, b);
  }
}

__global__ void gn_finalize(const double* __restrict__ stats, int NG, double count, float eps, float* __restrict__ mean,
                            float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NG) return;
  const double m = stats[2 * i] / count;
  double var = stats[2 * i + 1] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = float(m);
  rstd[i] = float(1.0 / sqrt(var + double(eps)));
}

__global__ void __launch_bounds__(kThreads)
gn_apply(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
         const float* __restrict__ gamma, const float* __restrict__ beta, long long HW, int C, int G, long long total4,
         float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index over [N, HW, C/4]
  if (i >= total4) return;
  const int lanes = C / 4;
  const int c4 = int(i % lanes);
  const long long n = i / (HW * lanes);
  const int g = (c4 * 4) / (C / G);
  const float m = __ldg(mean + n * G + g), r = __ldg(rstd + n * G + g);
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  float4 o;
  o.x = (v.x - m) * r * ga.x + be.x; o.y = (v.y - m) * r * ga.y + be.y;
  o.z = (v.z - m) * r * ga.z + be.z; o.w = (v.w - m) * r * ga.w + be.w;
  reinterpret_cast<float4*>(y)[i] = o;
}

// sums[n][g] = {sum dy*gamma, sum dy*gamma*x} (fp64); dgamma[c] += sum dy*xhat, dbeta[c] += sum dy (fp32 atomics per CTA)
__global__ void __launch_bounds__(kThreads)
gn_bwd_sums(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
            const float* __restrict__ rstd, const float* __restrict__ gamma, long long HW, int C, int G,
            double* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sh[];                     // [10][rows][C/4]
  const int lanes = C / 4, rows = kThreads / lanes;
  const int c4 = threadIdx.x % lanes, pr = threadIdx.x / lanes;
  const int n = blockIdx.y;
  const int g = (c4 * 4) / (C / G);
  const long long p0 = (long long)blockIdx.x * kSlab, p1 = min(p0 + kSlab, HW);
  const long long base = ((long long)n * HW) * lanes + c4;
  const float m = __ldg(mean + n * G + g), r = __ldg(rstd + n * G + g);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
  float a1 = 0.f, a2 = 0.f;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
  if (pr < rows)
    for (long long p = p0 + pr; p < p1; p += rows) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(dy) + base + p * lanes);
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + base + p * lanes);
      const float4 t = make_float4(d.x * ga.x, d.y * ga.y, d.z * ga.z, d.w * ga.w);
      a1 += sum4(t);
      a2 += (t.x * v.x + t.y * v.y) + (t.z * v.z + t.w * v.w);
      dg.x += d.x * (v.x - m) * r; dg.y += d.y * (v.y - m) * r; dg.z += d.z * (v.z - m) * r; dg.w += d.w * (v.w - m) * r;
      db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
    }
  const int stride = rows * lanes;
  float* cell = sh + pr * lanes + c4;
  cell[0] = a1; cell[stride] = a2;
  cell[2 * stride] = dg.x; cell[3 * stride] = dg.y; cell[4 * stride] = dg.z; cell[5 * stride] = dg.w;
  cell[6 * stride] = db.x; cell[7 * stride] = db.y; cell[8 * stride] = db.z; cell[9 * stride] = db.w;
  __syncthreads();
  const int lpg = (C / G) / 4;
  if (threadIdx.x < G) {
    double a = 0.0, b = 0.0;
    for (int rr = 0; rr < rows; ++rr)
      for (int l = 0; l < lpg; ++l) {
        a += double(sh[rr * lanes + threadIdx.x * lpg + l]);
        b += double(sh[stride + rr * lanes + threadIdx.x * lpg + l]);
      }
    atomicAdd(sums + ((long long)n * G + threadIdx.x) * 2, a);
    atomicAdd(sums + ((long long)n * G + threadIdx.x) * 2 + 1, b);
  }
  // per-channel partials: thread t < C handles channel t
  for (int c = threadIdx.x; c < C; c += kThreads) {
    const int l = c / 4, e = c % 4;
    float sg = 0.f, sb = 0.f;
    for (int rr = 0; rr < rows; ++rr) {
      sg += sh[(2 + e) * stride + rr * lanes + l];
      sb += sh[(6 + e) * stride + rr * lanes + l];
    }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
  }
}

__global__ void __launch_bounds__(kThreads)
gn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
             const float* __restrict__ rstd, const float* __restrict__ gamma, const double* __restrict__ sums, long long HW,
             int C, int G, long long total4, double count, float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int lanes = C / 4;
  const int c4 = int(i % lanes);
  const long long n = i / (HW * lanes);
  const int g = (c4 * 4) / (C / G);
  const float m = __ldg(mean + n * G + g), r = __ldg(rstd + n * G + g);
  const double s1 = sums[(n * G + g) * 2], s2 = sums[(n * G + g) * 2 + 1];
  const float c2 = float(s1 / count);                                  // mean of dy*gamma
  const float c1 = float((s2 - double(m) * s1) * double(r) / count);  // mean of dy*gamma*xhat
  const float4 d = __ldg(reinterpret_cast<const float4*>(dy) + i), v = __ldg(reinterpret_cast<const float4*>(x) + i);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
  float4 o;
  o.x = r * (d.x * ga.x - c2 - (v.x - m) * r * c1); o.y = r * (d.y * ga.y - c2 - (v.y - m) * r * c1);
  o.z = r * (d.z * ga.z - c2 - (v.z - m) * r * c1); o.w = r * (d.w * ga.w - c2 - (v.w - m) * r * c1);
  reinterpret_cast<float4*>(dx)[i] = o;
}

int check(const void* a, const void* b, int N, long long HW, int C, int G) {
  if (!a || !b || N <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G != 0 || (C / G) % 4 != 0 || C % 4 != 0 || C > 1024 ||
      kThreads % (C / 4) != 0 || G > kThreads) {
    snprintf(g_gn_err, sizeof g_gn_err, "datr_groupnorm: null pointer or unsupported shape (C %% G == 0, (C / G) %% 4 == 0, 256 %% (C / 4) == 0)");
    return -1;
  }
  return 0;
}

int done(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_gn_err, sizeof g_gn_err, "%s: %s", what, cudaGetErrorString(e)); return -3; }
  g_gn_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace

extern "C" {

int datr_groupnorm_nhwc_forward(const float* x, const float* gamma, const float* beta, int N, long long HW, int C, int G, float eps,
                                float* y, float* mean, float* rstd, double* scratch, void* stream_) {
  if (int rc = check(x, y, N, HW, C, G)) return rc;
  if (!gamma || !beta || !mean || !rstd || !scratch) { snprintf(g_gn_err, sizeof g_gn_err, "datr_groupnorm: null pointer"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * N * G, st) != cudaSuccess) return done("cudaMemsetAsync");
  const dim3 grid(unsigned((HW + kSlab - 1) / kSlab), unsigned(N));
  const size_t smem = sizeof(float) * 2 * kThreads;
  gn_stats<<<grid, kThreads, smem, st>>>(x, HW, C, G, scratch);
  gn_finalize<<<(N * G + 127) / 128, 128, 0, st>>>(scratch, N * G, double(HW) * (C / G), eps, mean, rstd);
  const long long total4 = (long long)N * HW * (C / 4);
  gn_apply<<<unsigned((total4 + kThreads - 1) / kThreads), kThreads, 0, st>>>(x, mean, rstd, gamma, beta, HW, C, G, total4, y);
  return done("groupnorm forward");
}

int datr_groupnorm_nhwc_backward(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int N,
                                 long long HW, int C, int G, float* dx, float* dgamma, float* dbeta, double* scratch,
                                 void* stream_) {
  if (int rc = check(dy, dx, N, HW, C, G)) return rc;
  if (!x || !mean || !rstd || !gamma || !dgamma || !dbeta || !scratch) { snprintf(g_gn_err, sizeof g_gn_err, "datr_groupnorm: null pointer"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * N * G, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st);
  if (e != cudaSuccess) return done("cudaMemsetAsync");
  const dim3 grid(unsigned((HW + kSlab - 1) / kSlab), unsigned(N));
  const size_t smem = sizeof(float) * 10 * kThreads;
  gn_bwd_sums<<<grid, kThreads, smem, st>>>(dy, x, mean, rstd, gamma, HW, C, G, scratch, dgamma, dbeta);
  const long long total4 = (long long)N * HW * (C / 4);
  gn_bwd_apply<<<unsigned((total4 + kThreads - 1) / kThreads), kThreads, 0, st>>>(dy, x, mean, rstd, gamma, scratch, HW, C, G,
                                                                                  total4, double(HW) * (C / G), dx);
  return done("groupnorm backward");
}

const char* datr_groupnorm_last_error(void) { return g_gn_err; }
uint64_t datr_groupnorm_launch_count(void) { return g_gn_launches.load(std::memory_order_relaxed); }

}  // extern "C"
