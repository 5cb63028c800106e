// msda.cu -- multi-scale deformable attention for NVIDIA B200 (sm_100a).
//
// Forward: out[b,q,m,:] = sum_{l,p} attn[b,q,m,l,p] * bilinear(value[b, level l, :, m, :], loc[b,q,m,l,p])
// Backward: grad_value (scatter-add), grad_sampling_loc, grad_attn_weight.
// Semantics follow the reference CUDA op (file:line relative to the reference repo,
// models/dino/ops/src/cuda/ms_deform_im2col_cuda.cuh): pixel coordinate = loc*size - 0.5 (:285-286),
// a sample contributes only if -1 < y < H and -1 < x < W (:288), the four corners are individually
// bounds-checked (:56-79), grad_loc = (W*d/dx, H*d/dy) * grad_out * attn (:157-158),
// grad_attn = <grad_out, bilinear value> (:156).
//
// Design (this file is not a translation of the reference kernels):
//   * fp32, 32 channels/head (the DINO configuration): eight lanes own one (b,q,m) row, each lane
//     holds a float4 of channels, so one corner of one sample is a single 128-byte line read by one
//     quarter-warp, and a warp instruction covers four rows.  Corner addresses are clamped and the
//     bounds test is folded into the weights, which makes every load unconditional: the P sample
//     points of a level are issued as one batch of 4*P independent 16-byte loads per lane.
//   * a CTA owns 32 consecutive queries of ONE head so that neighbouring queries (which sample
//     neighbouring pixels in the encoder) share L1 lines.
//   * backward: channel reductions for grad_loc / grad_attn are 8-lane shuffle butterflies (the
//     reference stages them in shared memory and sums serially on thread 0, cuh:377-393);
//     grad_value uses 16-byte vector reductions (red.global.add.v4.f32), one per corner per lane,
//     instead of 4 scalar atomics.
//   * the P = 4 forward kernel (the DINO configuration; `_p4c` below) additionally removes the
//     per-lane redundancy of the sample geometry: the 8 lanes of a row split its L*P samples between
//     them (coalesced loc / attn loads), compute the geometry once and publish it to the other lanes
//     as 16-byte records in a conflict-free per-warp shared-memory table.  Measured on B200 at the
//     DINO-4scale encoder shape this halves the issued instructions (158 M -> 82 M warp
//     instructions) and leaves the kernel bound by L1 line throughput: one 128-byte wavefront per
//     (row, corner), 22.8 M per call, ~68 % of the l1tex data-pipe peak (profiles/).  For the
//     backward kernel the same restructuring measured slower (the kernel is bound by L2 vector
//     reductions: 82.5 M red sectors per call), so it keeps the direct per-lane geometry.
//   * any other channel count, and fp64, run the generic warp-per-row kernels below.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "datr_msda.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof g_err, fmt, detail);
  return code;
}

// ------------------------------------------------------------------------------------------------
// Sample geometry shared by all kernels.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Tap {
  int o[4];    // clamped pixel index y*W+x of corners (y0,x0) (y0,x1) (y1,x0) (y1,x1)
  bool in[4];  // corner lies inside the map AND the sample passes the (-1,H)x(-1,W) guard
  T ly, lx, hy, hx;
};

template <typename T>
__device__ __forceinline__ Tap<T> locate(T locx, T locy, int H, int W) {
  Tap<T> t;
  const T y = locy * T(H) - T(0.5);
  const T x = locx * T(W) - T(0.5);
  const bool ok = (y > T(-1)) && (x > T(-1)) && (y < T(H)) && (x < T(W));
  const T yf = floor(y), xf = floor(x);
  // out-of-range (or NaN) coordinates are rejected by `ok`; clamp before the int conversion
  const int y0 = ok ? int(yf) : 0, x0 = ok ? int(xf) : 0;
  t.ly = y - yf; t.lx = x - xf; t.hy = T(1) - t.ly; t.hx = T(1) - t.lx;
  const bool y0in = ok && y0 >= 0, y1in = ok && y0 + 1 <= H - 1;
  const bool x0in = x0 >= 0, x1in = x0 + 1 <= W - 1;
  const int y0c = max(y0, 0), y1c = min(y0 + 1, H - 1);
  const int x0c = max(x0, 0), x1c = min(x0 + 1, W - 1);
  t.o[0] = y0c * W + x0c; t.o[1] = y0c * W + x1c; t.o[2] = y1c * W + x0c; t.o[3] = y1c * W + x1c;
  t.in[0] = y0in && x0in; t.in[1] = y0in && x1in; t.in[2] = y1in && x0in; t.in[3] = y1in && x1in;
  return t;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ordered (volatile) variant: keeps a batch of loads ahead of the arithmetic that consumes them
__device__ __forceinline__ float4 ldg4_ordered(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

constexpr int kRowsPerCta = 32;  // 256 threads / 8 lanes per row

// ------------------------------------------------------------------------------------------------
// fp32, D = 32 forward.
// ------------------------------------------------------------------------------------------------
template <int kP>
__global__ void __launch_bounds__(256)
msda_fwd_f32_d32(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                 const float* __restrict__ attn, int N, int S, int M, int L, int Lq,
                 float* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  const int m = blockIdx.x % M;
  const long long bq = (long long)(blockIdx.x / M) * kRowsPerCta + (threadIdx.x >> 3);
  if (bq >= (long long)N * Lq) return;
  const int b = int(bq / Lq);
  const long long row = bq * M + m;
  const int rs = M * 32;  // floats between consecutive pixels
  const float* vb = value + (long long)b * S * rs + m * 32 + sub * 4;
  const float* lp = loc + row * L * kP * 2;
  const float* ap = attn + row * L * kP;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < L; ++l) {
    const int H = int(__ldg(shapes + 2 * l)), W = int(__ldg(shapes + 2 * l + 1));
    const float* vl = vb + (long long)__ldg(lstart + l) * rs;
    float4 v[kP][4];
    float w[kP][4];
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(lp) + l * kP + p);
      const float a = __ldg(ap + l * kP + p);
      const Tap<float> t = locate<float>(xy.x, xy.y, H, W);
      const float wy0 = t.hy * a, wy1 = t.ly * a;
      w[p][0] = t.in[0] ? wy0 * t.hx : 0.f;
      w[p][1] = t.in[1] ? wy0 * t.lx : 0.f;
      w[p][2] = t.in[2] ? wy1 * t.hx : 0.f;
      w[p][3] = t.in[3] ? wy1 * t.lx : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) v[p][i] = ldg4(vl + (long long)t.o[i] * rs);
    }
#pragma unroll
    for (int p = 0; p < kP; ++p)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc.x = fmaf(w[p][i], v[p][i].x, acc.x);
        acc.y = fmaf(w[p][i], v[p][i].y, acc.y);
        acc.z = fmaf(w[p][i], v[p][i].z, acc.z);
        acc.w = fmaf(w[p][i], v[p][i].w, acc.w);
      }
  }
  *reinterpret_cast<float4*>(out + row * 32 + sub * 4) = acc;
}

// ------------------------------------------------------------------------------------------------
// fp32, D = 32 backward.
// ------------------------------------------------------------------------------------------------
template <int kP>
__global__ void __launch_bounds__(256)
msda_bwd_f32_d32(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                 const float* __restrict__ attn, const float* __restrict__ grad_out,
                 int N, int S, int M, int L, int Lq,
                 float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  const int sub = threadIdx.x & 7;
  const int m = blockIdx.x % M;
  long long bq = (long long)(blockIdx.x / M) * kRowsPerCta + (threadIdx.x >> 3);
  // keep whole warps alive for the shuffles: out-of-range rows redo the last row and skip all writes
  const bool live = bq < (long long)N * Lq;
  if (!live) bq = (long long)N * Lq - 1;
  const int b = int(bq / Lq);
  const long long row = bq * M + m;
  const int rs = M * 32;
  const long long voff = (long long)b * S * rs + m * 32 + sub * 4;
  const float* vb = value + voff;
  float* gvb = grad_value + voff;
  const float* lp = loc + row * L * kP * 2;
  const float* ap = attn + row * L * kP;
  const float4 g = ldg4(grad_out + row * 32 + sub * 4);

  for (int l = 0; l < L; ++l) {
    const int H = int(__ldg(shapes + 2 * l)), W = int(__ldg(shapes + 2 * l + 1));
    const long long lo = (long long)__ldg(lstart + l) * rs;
    const float* vl = vb + lo;
    float* gvl = gvb + lo;
    float keep_a = 0.f, keep_x = 0.f, keep_y = 0.f;  // lane `p` keeps the results of sample p
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(lp) + l * kP + p);
      const float a = __ldg(ap + l * kP + p);
      const Tap<float> t = locate<float>(xy.x, xy.y, H, W);
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = ldg4(vl + (long long)t.o[i] * rs);
        if (!t.in[i]) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float4 tv = make_float4(g.x * a, g.y * a, g.z * a, g.w * a);
      const float cw[4] = {t.hy * t.hx, t.hy * t.lx, t.ly * t.hx, t.ly * t.lx};
      if (live) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (t.in[i]) red_add4(gvl + (long long)t.o[i] * rs, cw[i] * tv.x, cw[i] * tv.y, cw[i] * tv.z, cw[i] * tv.w);
      }
      float4 val, dy, dx;
#define DATR_MIX(c)                                                                  \
  val.c = cw[0] * v[0].c + cw[1] * v[1].c + cw[2] * v[2].c + cw[3] * v[3].c;         \
  dy.c = t.hx * (v[2].c - v[0].c) + t.lx * (v[3].c - v[1].c);                        \
  dx.c = t.hy * (v[1].c - v[0].c) + t.ly * (v[3].c - v[2].c);
      DATR_MIX(x) DATR_MIX(y) DATR_MIX(z) DATR_MIX(w)
#undef DATR_MIX
      const float pa = group8_sum(dot4(g, val));
      const float px = group8_sum(dot4(tv, dx)) * float(W);
      const float py = group8_sum(dot4(tv, dy)) * float(H);
      if (sub == p) { keep_a = pa; keep_x = px; keep_y = py; }
    }
    if (live && sub < kP) {
      grad_attn[row * L * kP + l * kP + sub] = keep_a;
      reinterpret_cast<float2*>(grad_loc)[row * L * kP + l * kP + sub] = make_float2(keep_x, keep_y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32, D = 32, P = 4 forward: the 8 lanes of a row split its L*P samples between them, compute the
// sample geometry once and publish it through a per-warp shared-memory table (a chunk = 16 samples
// = 4 levels).
// ------------------------------------------------------------------------------------------------
constexpr int kChunk = 16;                    // samples per chunk

struct LevelGeom { int H, W, start; };

__device__ __forceinline__ void load_levels(LevelGeom* sh, const int64_t* __restrict__ shapes,
                                            const int64_t* __restrict__ lstart, int L) {
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    sh[l].H = int(__ldg(shapes + 2 * l));
    sh[l].W = int(__ldg(shapes + 2 * l + 1));
    sh[l].start = int(__ldg(lstart + l));
  }
  __syncthreads();
}

constexpr int kMaxLevels = 32;

// ------------------------------------------------------------------------------------------------
// Compact geometry table (16 bytes per sample): {pixel index | flags << 26, lx, ly, attn}.
// flags: bits 0-3 corner (y0x0, y0x1, y1x0, y1x1) contributes; bit 4: x1 is a distinct pixel
// (x1c = x0c + 1); bit 5: y1 is a distinct row.  Lane `sub` prepares samples sub and sub + 8 of the
// chunk, so the table writes are conflict-free; rows of a warp are 17 slots apart (bank offset 4).
// ------------------------------------------------------------------------------------------------
constexpr int kCRow = kChunk + 1;

__device__ __forceinline__ uint4 pack_tap(float locx, float locy, float a, const LevelGeom& g, bool live) {
  const Tap<float> t = locate<float>(locx, locy, g.H, g.W);
  unsigned flags = (t.in[0] ? 1u : 0u) | (t.in[1] ? 2u : 0u) | (t.in[2] ? 4u : 0u) | (t.in[3] ? 8u : 0u);
  if (!live) flags = 0u;
  flags |= (t.o[1] != t.o[0]) ? 16u : 0u;
  flags |= (t.o[2] != t.o[0]) ? 32u : 0u;
  uint4 q;
  q.x = unsigned(g.start + t.o[0]) | (flags << 26);
  q.y = __float_as_uint(t.lx); q.z = __float_as_uint(t.ly); q.w = __float_as_uint(a);
  return q;
}

template <int kBatch>
__global__ void __launch_bounds__(256)
msda_fwd_f32_d32_p4c(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                     const float* __restrict__ attn, int N, int S, int M, int L, int Lq,
                     float* __restrict__ out) {
  __shared__ __align__(16) uint4 taps[8][4][kCRow];
  __shared__ LevelGeom geom[kMaxLevels];
  load_levels(geom, shapes, lstart, L);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 3, sub = lane & 7;
  const int m = blockIdx.x % M;
  long long bq = (long long)(blockIdx.x / M) * kRowsPerCta + warp * 4 + r;
  const bool live = bq < (long long)N * Lq;
  if (!live) bq = (long long)N * Lq - 1;
  const int b = int(bq / Lq);
  const long long row = bq * M + m;
  const int rs = M * 32;
  const int LP = L * 4;
  const float* vb = value + (long long)b * S * rs + m * 32 + sub * 4;
  uint4* mytaps = &taps[warp][r][0];

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = 0; s0 < LP; s0 += kChunk) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int s = s0 + sub + 8 * k;
      if (s < LP) {
        const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + row * LP + s);
        const float a = __ldg(attn + row * LP + s);
        mytaps[sub + 8 * k] = pack_tap(xy.x, xy.y, a, geom[s >> 2], true);
      }
    }
    __syncwarp();
    const int n = min(kChunk, LP - s0);
    for (int j0 = 0; j0 < n; j0 += 4) {
      const int W = geom[(s0 + j0) >> 2].W;
#pragma unroll
      for (int jb = 0; jb < 4; jb += kBatch) {
        float4 v[kBatch][4];
        float w[kBatch][4];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const uint4 q = mytaps[j0 + jb + j];
          const unsigned flags = q.x >> 26;
          const int base = int(q.x & 0x03ffffffu);
          const int dx = (flags >> 4) & 1, dy = (flags & 32u) ? W : 0;
          const float lx = __uint_as_float(q.y), ly = __uint_as_float(q.z), a = __uint_as_float(q.w);
          const float wy0 = (1.f - ly) * a, wy1 = ly * a, hx = 1.f - lx;
          w[j][0] = (flags & 1u) ? wy0 * hx : 0.f;
          w[j][1] = (flags & 2u) ? wy0 * lx : 0.f;
          w[j][2] = (flags & 4u) ? wy1 * hx : 0.f;
          w[j][3] = (flags & 8u) ? wy1 * lx : 0.f;
          const float* p00 = vb + (long long)base * rs;
          v[j][0] = ldg4_ordered(p00);
          v[j][1] = ldg4_ordered(p00 + dx * rs);
          v[j][2] = ldg4_ordered(p00 + (long long)dy * rs);
          v[j][3] = ldg4_ordered(p00 + (long long)(dy + dx) * rs);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc.x = fmaf(w[j][i], v[j][i].x, acc.x);
            acc.y = fmaf(w[j][i], v[j][i].y, acc.y);
            acc.z = fmaf(w[j][i], v[j][i].z, acc.z);
            acc.w = fmaf(w[j][i], v[j][i].w, acc.w);
          }
      }
    }
    __syncwarp();
  }
  if (live) *reinterpret_cast<float4*>(out + row * 32 + sub * 4) = acc;
}

// ------------------------------------------------------------------------------------------------
// Generic kernels: any channel count / point count, float or double.  One warp per (b,q,m) row,
// lanes stride over channels.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const T* __restrict__ loc, const T* __restrict__ attn,
                 long long rows, int S, int M, int D, int L, int Lq, int P, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int m = int(row % M);
  const int b = int(row / ((long long)M * Lq));
  const long long rs = (long long)M * D;
  const T* vb = value + (long long)b * S * rs + (long long)m * D;
  for (int c0 = 0; c0 < D; c0 += 32) {
    const int c = c0 + lane;
    T acc = 0;
    for (int l = 0; l < L; ++l) {
      const int H = int(shapes[2 * l]), W = int(shapes[2 * l + 1]);
      const T* vl = vb + lstart[l] * rs;
      for (int p = 0; p < P; ++p) {
        const long long k = (row * L + l) * P + p;
        const Tap<T> t = locate<T>(loc[2 * k], loc[2 * k + 1], H, W);
        const T a = attn[k];
        if (c < D) {
          const T cw[4] = {t.hy * t.hx, t.hy * t.lx, t.ly * t.hx, t.ly * t.lx};
          T s = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (t.in[i]) s += cw[i] * vl[(long long)t.o[i] * rs + c];
          acc += s * a;
        }
      }
    }
    if (c < D) out[row * D + c] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
msda_bwd_generic(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const T* __restrict__ loc, const T* __restrict__ attn,
                 const T* __restrict__ grad_out, long long rows, int S, int M, int D, int L, int Lq, int P,
                 T* __restrict__ grad_value, T* __restrict__ grad_loc, T* __restrict__ grad_attn) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;  // whole warps share a row: no partial-warp exit
  const int m = int(row % M);
  const int b = int(row / ((long long)M * Lq));
  const long long rs = (long long)M * D;
  const long long voff = (long long)b * S * rs + (long long)m * D;
  for (int l = 0; l < L; ++l) {
    const int H = int(shapes[2 * l]), W = int(shapes[2 * l + 1]);
    const long long lo = voff + lstart[l] * rs;
    for (int p = 0; p < P; ++p) {
      const long long k = (row * L + l) * P + p;
      const Tap<T> t = locate<T>(loc[2 * k], loc[2 * k + 1], H, W);
      const T a = attn[k];
      const T cw[4] = {t.hy * t.hx, t.hy * t.lx, t.ly * t.hx, t.ly * t.lx};
      T pa = 0, px = 0, py = 0;
      for (int c = lane; c < D; c += 32) {
        const T top = grad_out[row * D + c], tv = top * a;
        T v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = 0;
          if (t.in[i]) {
            const long long e = lo + (long long)t.o[i] * rs + c;
            v[i] = value[e];
            atomicAdd(grad_value + e, cw[i] * tv);
          }
        }
        pa += top * (cw[0] * v[0] + cw[1] * v[1] + cw[2] * v[2] + cw[3] * v[3]);
        px += tv * (t.hy * (v[1] - v[0]) + t.ly * (v[3] - v[2]));
        py += tv * (t.hx * (v[2] - v[0]) + t.lx * (v[3] - v[1]));
      }
      pa = warp_sum(pa); px = warp_sum(px) * T(W); py = warp_sum(py) * T(H);
      if (lane == 0) { grad_attn[k] = pa; grad_loc[2 * k] = px; grad_loc[2 * k + 1] = py; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side.
// ------------------------------------------------------------------------------------------------
bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int check_common(const void* value, const int64_t* shapes, const int64_t* lstart, const void* loc,
                 const void* attn, int N, int S, int M, int D, int L, int Lq, int P, int dtype) {
  if (!value || !shapes || !lstart || !loc || !attn) return fail(DATR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (N <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0)
    return fail(DATR_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (dtype != DATR_DTYPE_F32 && dtype != DATR_DTYPE_F64) return fail(DATR_ERR_BAD_ARGUMENT, "unknown dtype%s");
  const size_t es = dtype == DATR_DTYPE_F32 ? 4 : 8;
  if (!aligned(value, es) || !aligned(loc, es) || !aligned(attn, es) || !aligned(shapes, 8) || !aligned(lstart, 8))
    return fail(DATR_ERR_ALIGNMENT, "buffer not aligned to its element type%s");
  if ((long long)N * Lq * M * L * P * 2 > (1LL << 40)) return fail(DATR_ERR_BAD_ARGUMENT, "problem too large%s");
  return DATR_OK;
}

int after_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return DATR_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_OK;
}

bool fast_ok(int D, int P, int dtype, const void* a, const void* b, const void* c, const void* d) {
  return dtype == DATR_DTYPE_F32 && D == 32 && (P == 4 || P == 1 || P == 2 || P == 8) && aligned(a, 16) &&
         aligned(b, 16) && aligned(c, 8) && aligned(d, 16);
}

}  // namespace

extern "C" {

int datr_msda_forward(const void* value, const int64_t* shapes, const int64_t* lstart, const void* loc,
                      const void* attn, int N, int S, int M, int D, int L, int Lq, int P, int dtype, void* out,
                      void* stream_) {
  if (int rc = check_common(value, shapes, lstart, loc, attn, N, S, M, D, L, Lq, P, dtype)) return rc;
  if (!out) return fail(DATR_ERR_BAD_ARGUMENT, "null output pointer%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long rows = (long long)N * Lq * M;
  if (fast_ok(D, P, dtype, value, out, loc, attn)) {
    const long long ctas = (((long long)N * Lq + kRowsPerCta - 1) / kRowsPerCta) * M;
    if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
    const float* v = static_cast<const float*>(value);
    const float* lc = static_cast<const float*>(loc);
    const float* at = static_cast<const float*>(attn);
    float* o = static_cast<float*>(out);
#define DATR_FWD(PP) msda_fwd_f32_d32<PP><<<(unsigned)ctas, 256, 0, stream>>>(v, shapes, lstart, lc, at, N, S, M, L, Lq, o)
    if (P == 4 && L <= kMaxLevels && (long long)S < (1LL << 26)) {
      msda_fwd_f32_d32_p4c<2><<<(unsigned)ctas, 256, 0, stream>>>(v, shapes, lstart, lc, at, N, S, M, L, Lq, o);
      return after_launch("msda_fwd_f32_d32_p4c");
    }
    switch (P) {
      case 1: DATR_FWD(1); break;
      case 2: DATR_FWD(2); break;
      case 4: DATR_FWD(4); break;
      default: DATR_FWD(8); break;
    }
#undef DATR_FWD
    return after_launch("msda_fwd_f32_d32");
  }
  const long long ctas = (rows + 7) / 8;
  if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
  if (dtype == DATR_DTYPE_F32)
    msda_fwd_generic<float><<<(unsigned)ctas, 256, 0, stream>>>(
        static_cast<const float*>(value), shapes, lstart, static_cast<const float*>(loc),
        static_cast<const float*>(attn), rows, S, M, D, L, Lq, P, static_cast<float*>(out));
  else
    msda_fwd_generic<double><<<(unsigned)ctas, 256, 0, stream>>>(
        static_cast<const double*>(value), shapes, lstart, static_cast<const double*>(loc),
        static_cast<const double*>(attn), rows, S, M, D, L, Lq, P, static_cast<double*>(out));
  return after_launch("msda_fwd_generic");
}

int datr_msda_backward(const void* value, const int64_t* shapes, const int64_t* lstart, const void* loc,
                       const void* attn, const void* grad_out, int N, int S, int M, int D, int L, int Lq, int P,
                       int dtype, void* grad_value, void* grad_loc, void* grad_attn, void* stream_) {
  if (int rc = check_common(value, shapes, lstart, loc, attn, N, S, M, D, L, Lq, P, dtype)) return rc;
  if (!grad_out || !grad_value || !grad_loc || !grad_attn) return fail(DATR_ERR_BAD_ARGUMENT, "null gradient pointer%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t es = dtype == DATR_DTYPE_F32 ? 4 : 8;
  const cudaError_t me = cudaMemsetAsync(grad_value, 0, es * (size_t)N * S * M * D, stream);
  if (me != cudaSuccess) return fail(DATR_ERR_CUDA, "cudaMemsetAsync(grad_value): %s", cudaGetErrorString(me));
  const long long rows = (long long)N * Lq * M;
  if (fast_ok(D, P, dtype, value, grad_out, grad_loc, grad_attn) && aligned(grad_value, 16) && aligned(loc, 8) &&
      aligned(attn, 4)) {
    const long long ctas = (((long long)N * Lq + kRowsPerCta - 1) / kRowsPerCta) * M;
    if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
    const float* v = static_cast<const float*>(value);
    const float* lc = static_cast<const float*>(loc);
    const float* at = static_cast<const float*>(attn);
    const float* go = static_cast<const float*>(grad_out);
    float* gv = static_cast<float*>(grad_value);
    float* gl = static_cast<float*>(grad_loc);
    float* ga = static_cast<float*>(grad_attn);
#define DATR_BWD(PP) \
  msda_bwd_f32_d32<PP><<<(unsigned)ctas, 256, 0, stream>>>(v, shapes, lstart, lc, at, go, N, S, M, L, Lq, gv, gl, ga)
    switch (P) {
      case 1: DATR_BWD(1); break;
      case 2: DATR_BWD(2); break;
      case 4: DATR_BWD(4); break;
      default: DATR_BWD(8); break;
    }
#undef DATR_BWD
    return after_launch("msda_bwd_f32_d32");
  }
  const long long ctas = (rows + 7) / 8;
  if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
  if (dtype == DATR_DTYPE_F32)
    msda_bwd_generic<float><<<(unsigned)ctas, 256, 0, stream>>>(
        static_cast<const float*>(value), shapes, lstart, static_cast<const float*>(loc),
        static_cast<const float*>(attn), static_cast<const float*>(grad_out), rows, S, M, D, L, Lq, P,
        static_cast<float*>(grad_value), static_cast<float*>(grad_loc), static_cast<float*>(grad_attn));
  else
    msda_bwd_generic<double><<<(unsigned)ctas, 256, 0, stream>>>(
        static_cast<const double*>(value), shapes, lstart, static_cast<const double*>(loc),
        static_cast<const double*>(attn), static_cast<const double*>(grad_out), rows, S, M, D, L, Lq, P,
        static_cast<double*>(grad_value), static_cast<double*>(grad_loc), static_cast<double*>(grad_attn));
  return after_launch("msda_bwd_generic");
}

const char* datr_last_error(void) { return g_err; }
int datr_abi_version(void) { return 1; }
uint64_t datr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
