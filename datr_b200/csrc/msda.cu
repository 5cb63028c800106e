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
//     quarter-warp, and a warp instruction covers four rows.  A CTA owns 32 consecutive queries of
//     ONE head so that neighbouring queries (which sample neighbouring pixels in the encoder) share
//     L1 lines.
//   * sample geometry is computed ONCE per sample (the 8 lanes of a row split its L*P samples) and
//     published to the row's lanes through a conflict-free shared-memory slot table; the 2x2 pixel
//     block of a sample is anchored so that its four addresses are level-uniform offsets of one
//     anchor pixel and always in range -- bounds handling lives in the slot weights, every load is
//     unconditional, and the inner loop is LDS + mad.wide + LDG.128 + FMA only (see the comment
//     above `place`).  First B200 profile of the previous per-lane-geometry kernels showed both
//     directions issue-bound (65-72 % issue slots, 40 % of them address arithmetic; profiles/r01a_*).
//   * backward: channel reductions for grad_loc / grad_attn are 8-lane shuffle butterflies over
//     three per-lane partial sums (the reference stages them in shared memory and sums serially on
//     thread 0, cuh:377-393); grad_value uses 16-byte vector reductions (red.global.add.v4.f32),
//     one per slot per lane, predicated off for empty slots, instead of 4 scalar atomics.
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
// Sample geometry of the generic kernels.
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

// ------------------------------------------------------------------------------------------------
// fp32, D = 32 kernels (the DINO configuration).
//
// Work split: a CTA of 256 threads owns 32 consecutive queries of ONE head; eight lanes own one
// (b,q,m) row and each lane holds a float4 of its 32 channels, so a corner of a sample is one
// 128-byte line read by a quarter-warp.
//
// Stage 1 (per warp, once per row): the 8 lanes of a row split its L*P samples, load loc / attn
// coalesced, and turn each sample into a SLOT record in a per-warp shared-memory table:
//   the 2x2 pixel block is anchored at (by,bx) = clamp((y0,x0), 0, size-2), so its four pixels sit
//   at the level-uniform offsets {0, 1, W, W+1} (or 0 where a level is one pixel wide/high) and are
//   always inside the level -- every load is unconditional and needs ONE mad.wide per corner;
//   what would have been bounds tests becomes the weight of each slot: slot i of an axis carries
//   weight h (=1-l, d/dcoord = -1), weight l (d/dcoord = +1) or nothing, depending on which of the
//   sample's two corners landed on it (cuh:56-79 bounds rules, :288 sample guard).
// Stage 2: every lane walks the table: 1-3 LDS.128, 4 mad.wide, 4 LDG.128 (+4 predicated
//   RED.128 in backward) per sample, and pure FMAs.
// ------------------------------------------------------------------------------------------------
struct LevelGeom { int H, W, start; };
constexpr int kMaxLevels = 32;
constexpr int kRowsPerCta = 32;  // 256 threads / 8 lanes per row
constexpr int kMaxTaps = 32;     // L*P limit of the fast path

struct Slots {
  int pix;                  // pixel index (level start included) of the anchor (by,bx)
  float a;                  // attention weight (0 if the sample is rejected)
  float wy0, wy1, wx0, wx1; // interpolation weight carried by each slot (0 = slot unused)
  float sy0, sy1, sx0, sx1; // d(weight)/d(coordinate) of each slot: -1, +1 or 0
};

__device__ __forceinline__ Slots place(float locx, float locy, float a, const LevelGeom& g) {
  const int H = g.H, W = g.W;
  const float y = locy * float(H) - 0.5f;
  const float x = locx * float(W) - 0.5f;
  const bool ok = (y > -1.f) && (x > -1.f) && (y < float(H)) && (x < float(W));  // cuh:288; false for NaN
  const float yf = floorf(y), xf = floorf(x);
  const int y0 = ok ? int(yf) : 0, x0 = ok ? int(xf) : 0;
  const float ly = y - yf, lx = x - xf, hy = 1.f - ly, hx = 1.f - lx;
  const int by = min(max(y0, 0), max(H - 2, 0)), bx = min(max(x0, 0), max(W - 2, 0));
  Slots s;
  s.pix = g.start + by * W + bx;
  s.a = ok ? a : 0.f;
  const bool y_on0 = ok && y0 == by, y1_on0 = ok && y0 + 1 == by;          // slot 0 = row by
  s.wy0 = y_on0 ? hy : (y1_on0 ? ly : 0.f);
  s.sy0 = y_on0 ? -1.f : (y1_on0 ? 1.f : 0.f);
  const bool row1 = ok && H >= 2;                                          // slot 1 = row by+1
  const bool y1_on1 = row1 && y0 == by, y_on1 = row1 && y0 == by + 1;
  s.wy1 = y1_on1 ? ly : (y_on1 ? hy : 0.f);
  s.sy1 = y1_on1 ? 1.f : (y_on1 ? -1.f : 0.f);
  const bool x_on0 = ok && x0 == bx, x1_on0 = ok && x0 + 1 == bx;
  s.wx0 = x_on0 ? hx : (x1_on0 ? lx : 0.f);
  s.sx0 = x_on0 ? -1.f : (x1_on0 ? 1.f : 0.f);
  const bool col1 = ok && W >= 2;
  const bool x1_on1 = col1 && x0 == bx, x_on1 = col1 && x0 == bx + 1;
  s.wx1 = x1_on1 ? lx : (x_on1 ? hx : 0.f);
  s.sx1 = x1_on1 ? 1.f : (x_on1 ? -1.f : 0.f);
  return s;
}

__device__ __forceinline__ void load_levels(LevelGeom* sh, const int64_t* __restrict__ shapes,
                                            const int64_t* __restrict__ lstart, int L) {
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    sh[l].H = int(__ldg(shapes + 2 * l));
    sh[l].W = int(__ldg(shapes + 2 * l + 1));
    sh[l].start = int(__ldg(lstart + l));
  }
  __syncthreads();
}

// base + pix * stride_bytes in one IMAD.WIDE (volatile: keeps ptxas from splitting it into a shared
// product plus a 64-bit add per corner, which doubles the address instructions of the inner loop)
__device__ __forceinline__ const float* pixel_ptr(const float* base, int pix, int stride_bytes) {
  unsigned long long r;
  asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(pix), "r"(stride_bytes), "l"(base));
  return reinterpret_cast<const float*>(r);
}

// ordered (volatile) 16-byte read-only load: keeps a batch of loads ahead of the arithmetic
__device__ __forceinline__ float4 ldg4_ordered(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 16-byte vector reduction, skipped when the slot weight is zero
__device__ __forceinline__ void red_add4_if(const float* p, float w, const float4& g) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.neu.f32 q, %5, 0f00000000;\n\t"
      "@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
      ::"l"(p), "f"(w * g.x), "f"(w * g.y), "f"(w * g.z), "f"(w * g.w), "f"(w) : "memory");
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__device__ __forceinline__ float group8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}

// ------------------------------------------------------------------------------------------------
// Module-level fusion (kMode 1 / 2): the kernels take what the reference's MSDeformAttn.forward feeds
// into its elementwise prologue (models/dino/ops/modules/ms_deform_attn.py:99-111) instead of the
// materialised sampling_locations / attention_weights:
//   raw sampling offsets [N,Lq,M,L,P,2], raw attention logits [N,Lq,M,L*P], reference points [N,Lq,L,R]
//   kMode 1 (R = 2, encoder):  loc = ref + off / (W_l, H_l)                       (:102-105)
//   kMode 2 (R = 4, decoder):  loc = ref.xy + off / P * ref.wh * 0.5              (:106-108)
//   attn = softmax over the L*P logits of a (query, head) row                     (:101)
// with the same operation order as the torch expressions.  The 8 lanes of a row hold its samples
// (lane `sub` owns samples sub, sub+8, ...), so the softmax is two 8-lane shuffle butterflies.
// ------------------------------------------------------------------------------------------------
constexpr int kChunks = 4;  // kMaxTaps / 8

struct RowTaps {
  float x[kChunks], y[kChunks], a[kChunks];  // location and softmax weight of samples c*8+sub
};

// `off` / `logit` point at the (query, head) row: offsets[bq][m][0][0][0], logits[bq][m][0]
template <int kMode, int kP>
__device__ __forceinline__ RowTaps fused_taps(const float* __restrict__ off, const float* __restrict__ logit,
                                              const float* __restrict__ ref, long long bq, int sub,
                                              int L, int LP, const LevelGeom* geom) {
  RowTaps t;
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int s = c * 8 + sub;
    t.x[c] = t.y[c] = 0.f;
    t.a[c] = -INFINITY;
    if (s < LP) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(off) + s);
      t.x[c] = xy.x; t.y[c] = xy.y;
      t.a[c] = __ldg(logit + s);
      mx = fmaxf(mx, t.a[c]);
    }
  }
  mx = group8_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    t.a[c] = (c * 8 + sub < LP) ? expf(t.a[c] - mx) : 0.f;
    sum += t.a[c];
  }
  sum = group8_sum(sum);
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int s = c * 8 + sub;
    t.a[c] = t.a[c] / sum;
    if (s < LP) {
      const int l = s / kP;
      if (kMode == 1) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(ref) + bq * L + l);
        t.x[c] = r.x + t.x[c] / float(geom[l].W);
        t.y[c] = r.y + t.y[c] / float(geom[l].H);
      } else {
        const float4 r = __ldg(reinterpret_cast<const float4*>(ref) + bq * L + l);
        t.x[c] = r.x + ((t.x[c] / float(kP)) * r.z) * 0.5f;
        t.y[c] = r.y + ((t.y[c] / float(kP)) * r.w) * 0.5f;
      }
    }
  }
  return t;
}

// row stride (in 16-byte records) of the slot tables: 4 rows of a warp read 4 different records per
// LDS.128; they fall on disjoint bank groups iff stride mod 8 is not 0 or 4.
__host__ __device__ inline int table_stride(int taps) {
  int s = taps + 1;
  if ((s & 3) == 0) ++s;
  return s;
}
__host__ __device__ inline int pix_stride(int taps) { return taps | 1; }

constexpr int kGeomBytes = 512;  // kMaxLevels * sizeof(LevelGeom) rounded up

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int kP, int kBatch, int kMode>
__global__ void __launch_bounds__(256)
msda_fwd_f32_d32(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                 const float* __restrict__ attn, const float* __restrict__ ref, long long ostride, long long lstride,
                 int N, int S, int M, int L, int Lq, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  LevelGeom* geom = reinterpret_cast<LevelGeom*>(smem);
  const int LP = L * kP, ws = table_stride(LP), ps = pix_stride(LP);
  uint4* wtab = reinterpret_cast<uint4*>(smem + kGeomBytes);            // [32][ws] {w00,w01,w10,w11} * attn
  int* ptab = reinterpret_cast<int*>(wtab + kRowsPerCta * ws);          // [32][ps] anchor pixel
  load_levels(geom, shapes, lstart, L);

  const int r = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int m = blockIdx.x % M;
  long long bq = (long long)(blockIdx.x / M) * kRowsPerCta + r;
  const bool live = bq < (long long)N * Lq;
  if (!live) bq = (long long)N * Lq - 1;
  const int b = int(bq / Lq);
  const long long row = bq * M + m;
  const int rs4 = M * 32 * 4;  // bytes between consecutive pixels
  const float* vb = value + (long long)b * S * (M * 32) + m * 32 + sub * 4;
  uint4* myw = wtab + r * ws;
  int* myp = ptab + r * ps;

  auto publish = [&](int s, float x, float y, float a) {
    const Slots t = place(x, y, a, geom[s / kP]);
    const float ay0 = t.wy0 * t.a, ay1 = t.wy1 * t.a;
    myw[s] = make_uint4(__float_as_uint(ay0 * t.wx0), __float_as_uint(ay0 * t.wx1),
                        __float_as_uint(ay1 * t.wx0), __float_as_uint(ay1 * t.wx1));
    myp[s] = t.pix;
  };
  if constexpr (kMode == 0) {
    for (int s = sub; s < LP; s += 8) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + row * LP + s);
      publish(s, xy.x, xy.y, __ldg(attn + row * LP + s));
    }
  } else {
    // fused modes: rows of the offsets / logits tensors may be strided (both can live in one merged GEMM output)
    const RowTaps t = fused_taps<kMode, kP>(loc + bq * ostride + (long long)m * LP * 2, attn + bq * lstride + (long long)m * LP,
                                            ref, bq, sub, L, LP, geom);
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
      if (c * 8 + sub < LP) publish(c * 8 + sub, t.x[c], t.y[c], t.a[c]);
  }
  __syncwarp();

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < L; ++l) {
    const int W = geom[l].W, H = geom[l].H;
    const int d01 = W >= 2 ? rs4 : 0;
    const long long d10 = H >= 2 ? (long long)W * rs4 : 0;
    const float* vb01 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(vb) + d01);
    const float* vb10 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(vb) + d10);
    const float* vb11 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(vb10) + d01);
#pragma unroll
    for (int p0 = 0; p0 < kP; p0 += kBatch) {
      float4 v[kBatch][4];
      uint4 w[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int s = l * kP + p0 + j;
        const int pix = myp[s];
        w[j] = myw[s];
        v[j][0] = ldg4_ordered(pixel_ptr(vb, pix, rs4));
        v[j][1] = ldg4_ordered(pixel_ptr(vb01, pix, rs4));
        v[j][2] = ldg4_ordered(pixel_ptr(vb10, pix, rs4));
        v[j][3] = ldg4_ordered(pixel_ptr(vb11, pix, rs4));
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const float wj[4] = {__uint_as_float(w[j].x), __uint_as_float(w[j].y), __uint_as_float(w[j].z),
                             __uint_as_float(w[j].w)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc.x = fmaf(wj[i], v[j][i].x, acc.x);
          acc.y = fmaf(wj[i], v[j][i].y, acc.y);
          acc.z = fmaf(wj[i], v[j][i].z, acc.z);
          acc.w = fmaf(wj[i], v[j][i].w, acc.w);
        }
      }
    }
  }
  if (live) *reinterpret_cast<float4*>(out + row * 32 + sub * 4) = acc;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int kP, int kMode>
__global__ void __launch_bounds__(256, 4)
msda_bwd_f32_d32(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                 const float* __restrict__ attn, const float* __restrict__ ref, const float* __restrict__ grad_out,
                 long long ostride, long long lstride, int N, int S, int M, int L, int Lq,
                 float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  extern __shared__ __align__(16) unsigned char smem[];
  LevelGeom* geom = reinterpret_cast<LevelGeom*>(smem);
  const int LP = L * kP, ws = table_stride(LP);
  uint4* tab0 = reinterpret_cast<uint4*>(smem + kGeomBytes);   // [32][ws] {pix, a, a*W, a*H}
  uint4* tab1 = tab0 + kRowsPerCta * ws;                       // {wy0, wy1, wx0, wx1}
  uint4* tab2 = tab1 + kRowsPerCta * ws;                       // {sy0, sy1, sx0, sx1}
  load_levels(geom, shapes, lstart, L);

  const int r = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int m = blockIdx.x % M;
  long long bq = (long long)(blockIdx.x / M) * kRowsPerCta + r;
  // keep whole warps alive for the shuffles: out-of-range rows redo the last row and skip all writes
  const bool live = bq < (long long)N * Lq;
  if (!live) bq = (long long)N * Lq - 1;
  const int b = int(bq / Lq);
  const long long row = bq * M + m;
  const int rs4 = M * 32 * 4;
  const long long voff = (long long)b * S * (M * 32) + m * 32 + sub * 4;
  const float* vb = value + voff;
  const float* gb = grad_value + voff;
  uint4* my0 = tab0 + r * ws;
  uint4* my1 = tab1 + r * ws;
  uint4* my2 = tab2 + r * ws;

  auto publish = [&](int s, float x, float y, float a) {
    const LevelGeom g = geom[s / kP];
    const Slots t = place(x, y, a, g);
    const float al = live ? t.a : 0.f;  // dead rows scatter nothing
    my0[s] = make_uint4(unsigned(t.pix), __float_as_uint(al), __float_as_uint(t.a * float(g.W)),
                        __float_as_uint(t.a * float(g.H)));
    my1[s] = make_uint4(__float_as_uint(t.wy0), __float_as_uint(t.wy1), __float_as_uint(t.wx0), __float_as_uint(t.wx1));
    my2[s] = make_uint4(__float_as_uint(t.sy0), __float_as_uint(t.sy1), __float_as_uint(t.sx0), __float_as_uint(t.sx1));
  };
  float soft[kChunks] = {0.f, 0.f, 0.f, 0.f};  // fused modes: softmax weight of samples c*8+sub
  if constexpr (kMode == 0) {
    for (int s = sub; s < LP; s += 8) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + row * LP + s);
      publish(s, xy.x, xy.y, __ldg(attn + row * LP + s));
    }
  } else {
    const RowTaps t = fused_taps<kMode, kP>(loc + bq * ostride + (long long)m * LP * 2, attn + bq * lstride + (long long)m * LP,
                                            ref, bq, sub, L, LP, geom);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      soft[c] = t.a[c];
      if (c * 8 + sub < LP) publish(c * 8 + sub, t.x[c], t.y[c], t.a[c]);
    }
  }
  const float4 g = ldg4(grad_out + row * 32 + sub * 4);
  __syncwarp();

  float keep_a = 0.f, keep_x = 0.f, keep_y = 0.f;  // lane (s & 7) keeps the results of sample s
  for (int l = 0; l < L; ++l) {
    const int W = geom[l].W, H = geom[l].H;
    const int d01 = W >= 2 ? rs4 : 0;
    const long long d10 = H >= 2 ? (long long)W * rs4 : 0;
    const long long d11 = d10 + d01;
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const int s = l * kP + p;
      const uint4 q0 = my0[s], q1 = my1[s], q2 = my2[s];
      const int pix = int(q0.x);
      const float a = __uint_as_float(q0.y);
      const float wy0 = __uint_as_float(q1.x), wy1 = __uint_as_float(q1.y);
      const float wx0 = __uint_as_float(q1.z), wx1 = __uint_as_float(q1.w);
      const float* p00 = pixel_ptr(vb, pix, rs4);
      const float4 v00 = ldg4_ordered(p00);
      const float4 v01 = ldg4_ordered(reinterpret_cast<const float*>(reinterpret_cast<const char*>(p00) + d01));
      const float4 v10 = ldg4_ordered(reinterpret_cast<const float*>(reinterpret_cast<const char*>(p00) + d10));
      const float4 v11 = ldg4_ordered(reinterpret_cast<const float*>(reinterpret_cast<const char*>(p00) + d11));
      // grad_value: slot weight * attn * grad_out, one 16-byte reduction per slot (skipped if weight 0)
      const float* g00 = pixel_ptr(gb, pix, rs4);
      const float ay0 = wy0 * a, ay1 = wy1 * a;
      red_add4_if(g00, ay0 * wx0, g);
      red_add4_if(reinterpret_cast<const float*>(reinterpret_cast<const char*>(g00) + d01), ay0 * wx1, g);
      red_add4_if(reinterpret_cast<const float*>(reinterpret_cast<const char*>(g00) + d10), ay1 * wx0, g);
      red_add4_if(reinterpret_cast<const float*>(reinterpret_cast<const char*>(g00) + d11), ay1 * wx1, g);
      // <grad_out, slot value> over this lane's 4 channels
      const float e00 = dot4(g, v00), e01 = dot4(g, v01), e10 = dot4(g, v10), e11 = dot4(g, v11);
      const float r0 = fmaf(wx1, e01, wx0 * e00), r1 = fmaf(wx1, e11, wx0 * e10);       // interpolate along x
      const float sx0 = __uint_as_float(q2.z), sx1 = __uint_as_float(q2.w);
      const float t0 = fmaf(sx1, e01, sx0 * e00), t1 = fmaf(sx1, e11, sx0 * e10);       // differentiate along x
      float pa = fmaf(wy1, r1, wy0 * r0);                                                // cuh:156
      float px = fmaf(wy1, t1, wy0 * t0);                                                // cuh:157 (x)
      float py = fmaf(__uint_as_float(q2.y), r1, __uint_as_float(q2.x) * r0);            // cuh:158 (y)
      pa = group8_sum(pa); px = group8_sum(px); py = group8_sum(py);
      if constexpr (kMode == 0) {
        if (sub == (s & 7)) { keep_a = pa; keep_x = px * __uint_as_float(q0.z); keep_y = py * __uint_as_float(q0.w); }
        if ((s & 7) == 7 || s == LP - 1) {
          const int s0 = s & ~7;
          if (live && s0 + sub <= s) {
            grad_attn[row * LP + s0 + sub] = keep_a;
            reinterpret_cast<float2*>(grad_loc)[row * LP + s0 + sub] = make_float2(keep_x, keep_y);
          }
        }
      } else {
        // every lane of the row has read record s (the shuffles above are warp-synchronous): reuse its slot of
        // table 2 for {d/d attn, d/d loc.x, d/d loc.y}; the owner lane of the sample collects it after the loop
        if (sub == (s & 7))
          my2[s] = make_uint4(__float_as_uint(pa), __float_as_uint(px * __uint_as_float(q0.z)),
                              __float_as_uint(py * __uint_as_float(q0.w)), 0u);
      }
    }
  }
  if constexpr (kMode != 0) {
    // softmax backward (d logit_s = a_s * (d a_s - sum_t a_t * d a_t)) and the chain rule of the location formula
    (void)keep_a; (void)keep_x; (void)keep_y;
    __syncwarp();
    uint4 res[kChunks];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      res[c] = make_uint4(0u, 0u, 0u, 0u);
      if (c * 8 + sub < LP) res[c] = my2[c * 8 + sub];
      dot = fmaf(soft[c], __uint_as_float(res[c].x), dot);
    }
    dot = group8_sum(dot);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 8 + sub;
      if (live && s < LP) {
        const int l = s / kP;
        float gx = __uint_as_float(res[c].y), gy = __uint_as_float(res[c].z);
        if (kMode == 1) {
          gx = gx / float(geom[l].W);
          gy = gy / float(geom[l].H);
        } else {
          const float4 rr = __ldg(reinterpret_cast<const float4*>(ref) + bq * L + l);
          gx = ((gx * 0.5f) * rr.z) / float(kP);
          gy = ((gy * 0.5f) * rr.w) / float(kP);
        }
        // gradients use the row strides of their inputs (a merged offsets+logits gradient is one GEMM operand)
        grad_attn[bq * lstride + (long long)m * LP + s] = soft[c] * (__uint_as_float(res[c].x) - dot);
        *reinterpret_cast<float2*>(grad_loc + bq * ostride + ((long long)m * LP + s) * 2) = make_float2(gx, gy);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Run kernels: encoder self-attention, where the queries ARE the pixels of the value maps.
//
// Neighbouring queries of one head sample neighbouring pixels of every level (their reference points are one pixel
// apart on their own level, 1/2, 1/4, ... pixel apart on the coarser ones, and the offsets come from a linear map of
// smoothly varying features).  The row kernels above give that locality to the caches only: a 2x2 footprint costs four
// 128-byte L1TEX wavefronts and -- in the backward -- four 128-byte reductions sent to L2, whether or not the previous
// query of the same (head, level, point) just touched the same pixels.  Both kernels sit on those two walls
// (profiles/r01o_msda_fused_ln_rowmask_ncu_full.txt: l1tex 2.07 cycles per line forward; 82.5 M red sectors = 2.64 GB
// on the SM->L2 path backward).
//
// Here the loop nest is transposed.  An 8-lane group owns ONE sample slot (level l, point p) of one head and WALKS a
// run of kRunLen consecutive queries, keeping the current 2x2 footprint in registers: the four value rows (float4 per
// lane) and, in the backward, four gradient accumulators.  With d = anchor(q) - anchor(q-1):
//   d == 0  same footprint: no load, no reduction, contributions add up in registers
//   d == 1  footprint moved one pixel right: the right column becomes the left column (register moves), two loads;
//           backward sends the two pixels that left the footprint to L2 (2 reds instead of 4)
//   else    four loads; backward flushes the four accumulators
// so L1TEX wavefronts and red traffic shrink by the run coherence of the data (up to 16x on the coarsest level of the
// encoder), and never grow: incoherent locations cost what the row kernels cost.
// Geometry (slot tables) is still computed once per sample by the 8 lanes of a (query, head) row (stage 1, shared
// with the row kernels: `place`, fused prologue) and published through shared memory; the forward's sum over slots is
// a two-step shuffle over the four groups of a warp plus a shared-memory partial per warp; the backward's
// per-sample results return to the row mapping through the slot table for the softmax / location chain rules.
// ------------------------------------------------------------------------------------------------
constexpr int kRunLen = 32;     // consecutive queries walked by one group
constexpr int kRunsPerCta = 2;  // runs (of the same head) per CTA
constexpr int kRunRows = kRunLen * kRunsPerCta;
constexpr int kFarPixel = -2;   // "no footprint yet": any real anchor is >= 2 away

__device__ __forceinline__ void red_add4_nz(const float* p, const float4& a) {
  if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
}

__device__ __forceinline__ const float* byte_off(const float* p, long long bytes) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p) + bytes);
}

// threads of a run: 8 lanes per slot, slots padded to whole warps
__host__ __device__ inline int run_threads(int taps) { return ((taps + 3) & ~3) * 8; }

template <int kP, int kMode>
__global__ void __maxnreg__(64)
msda_fwd_runs_f32_d32(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                      const float* __restrict__ attn, const float* __restrict__ ref, long long ostride, long long lstride,
                      int N, int S, int M, int L, int Lq, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  LevelGeom* geom = reinterpret_cast<LevelGeom*>(smem);
  const int LP = L * kP, ws = table_stride(LP), ps = pix_stride(LP);
  const int tpr = run_threads(LP), nw = tpr >> 5;
  uint4* wtab = reinterpret_cast<uint4*>(smem + kGeomBytes);                 // [64][ws] {w00,w01,w10,w11} * attn
  float4* part = reinterpret_cast<float4*>(wtab + kRunRows * ws);           // [runs][nw][kRunLen][8] warp partials
  int* ptab = reinterpret_cast<int*>(part + kRunsPerCta * nw * kRunLen * 8); // [64][ps] anchor pixel
  load_levels(geom, shapes, lstart, L);

  const int sub = threadIdx.x & 7;
  const int m = blockIdx.x % M;
  const int tiles = (Lq + kRunRows - 1) / kRunRows;
  const int tile = blockIdx.x / M;
  const int b = tile / tiles, q0 = (tile % tiles) * kRunRows;
  const int rs4 = M * 32 * 4;
  const float* vb = value + (long long)b * S * (M * 32) + m * 32 + sub * 4;

  // stage 1: the 8 lanes of a (query, head) row publish its slots (rows past the end repeat the last query)
  for (int rr = threadIdx.x >> 3; rr < kRunRows; rr += blockDim.x >> 3) {
    const long long bq = (long long)b * Lq + min(q0 + rr, Lq - 1);
    const long long row = bq * M + m;
    uint4* myw = wtab + rr * ws;
    int* myp = ptab + rr * ps;
    auto publish = [&](int s, float x, float y, float a) {
      const Slots t = place(x, y, a, geom[s / kP]);
      const float ay0 = t.wy0 * t.a, ay1 = t.wy1 * t.a;
      myw[s] = make_uint4(__float_as_uint(ay0 * t.wx0), __float_as_uint(ay0 * t.wx1),
                          __float_as_uint(ay1 * t.wx0), __float_as_uint(ay1 * t.wx1));
      myp[s] = t.pix;
    };
    if constexpr (kMode == 0) {
      for (int s = sub; s < LP; s += 8) {
        const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + row * LP + s);
        publish(s, xy.x, xy.y, __ldg(attn + row * LP + s));
      }
    } else {
      const RowTaps t = fused_taps<kMode, kP>(loc + bq * ostride + (long long)m * LP * 2, attn + bq * lstride + (long long)m * LP,
                                              ref, bq, sub, L, LP, geom);
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        if (c * 8 + sub < LP) publish(c * 8 + sub, t.x[c], t.y[c], t.a[c]);
    }
  }
  __syncthreads();

  // stage 2: group `slot` of a run walks the run's queries
  {
    const int run = threadIdx.x / tpr, tr = threadIdx.x - run * tpr;
    const int slot = tr >> 3, wir = tr >> 5;
    const bool used = slot < LP;               // padding groups of the last warp add zeros
    const int s = used ? slot : 0;
    const LevelGeom g = geom[s / kP];
    const bool wide = g.W >= 2;
    const int d01 = wide ? rs4 : 0;
    const long long d10 = g.H >= 2 ? (long long)g.W * rs4 : 0;
    const long long d11 = d10 + d01;
    const int n = max(0, min(kRunLen, Lq - (q0 + run * kRunLen)));
    const uint4* wrow = wtab + (run * kRunLen) * ws + s;
    const int* prow = ptab + (run * kRunLen) * ps + s;
    float4* prt = part + ((run * nw + wir) * kRunLen) * 8 + sub;
    int cur = kFarPixel;
    float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
    for (int i = 0; i < n; ++i) {
      const int pix = prow[i * ps];
      const uint4 w = wrow[i * ws];
      const int d = pix - cur;
      if (d != 0) {
        const bool shift = wide && d == 1;
        const float* p00 = pixel_ptr(vb, pix, rs4);
        if (shift) { v00 = v01; v10 = v11; }
        else { v00 = ldg4_ordered(p00); v10 = ldg4_ordered(byte_off(p00, d10)); }
        v01 = ldg4_ordered(byte_off(p00, d01));
        v11 = ldg4_ordered(byte_off(p00, d11));
        cur = pix;
      }
      const float w00 = used ? __uint_as_float(w.x) : 0.f, w01 = used ? __uint_as_float(w.y) : 0.f;
      const float w10 = used ? __uint_as_float(w.z) : 0.f, w11 = used ? __uint_as_float(w.w) : 0.f;
      float4 a;
      a.x = fmaf(w11, v11.x, fmaf(w10, v10.x, fmaf(w01, v01.x, w00 * v00.x)));
      a.y = fmaf(w11, v11.y, fmaf(w10, v10.y, fmaf(w01, v01.y, w00 * v00.y)));
      a.z = fmaf(w11, v11.z, fmaf(w10, v10.z, fmaf(w01, v01.z, w00 * v00.z)));
      a.w = fmaf(w11, v11.w, fmaf(w10, v10.w, fmaf(w01, v01.w, w00 * v00.w)));
      // sum over the four slots of this warp
      a.x += __shfl_xor_sync(0xffffffffu, a.x, 8);  a.y += __shfl_xor_sync(0xffffffffu, a.y, 8);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, 8);  a.w += __shfl_xor_sync(0xffffffffu, a.w, 8);
      a.x += __shfl_xor_sync(0xffffffffu, a.x, 16); a.y += __shfl_xor_sync(0xffffffffu, a.y, 16);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, 16); a.w += __shfl_xor_sync(0xffffffffu, a.w, 16);
      if ((threadIdx.x & 31) < 8) prt[i * 8] = a;
    }
  }
  __syncthreads();

  // stage 3: sum the warps of a run, one (query, head) row per 8 lanes
  for (int rr = threadIdx.x >> 3; rr < kRunRows; rr += blockDim.x >> 3) {
    const int q = q0 + rr;
    if (q >= Lq) continue;
    const int run = rr / kRunLen, i = rr - run * kRunLen;
    const float4* prt = part + ((run * nw) * kRunLen + i) * 8 + sub;
    float4 acc = prt[0];
    for (int w = 1; w < nw; ++w) {
      const float4 t = prt[w * kRunLen * 8];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *reinterpret_cast<float4*>(out + (((long long)b * Lq + q) * M + m) * 32 + sub * 4) = acc;
  }
}

template <int kP, int kMode>
__global__ void __maxnreg__(80)
msda_bwd_runs_f32_d32(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                      const float* __restrict__ attn, const float* __restrict__ ref, const float* __restrict__ grad_out,
                      long long ostride, long long lstride, int N, int S, int M, int L, int Lq,
                      float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  extern __shared__ __align__(16) unsigned char smem[];
  LevelGeom* geom = reinterpret_cast<LevelGeom*>(smem);
  const int LP = L * kP, ws = table_stride(LP);
  const int tpr = run_threads(LP);
  uint4* tab0 = reinterpret_cast<uint4*>(smem + kGeomBytes);   // [64][ws] {pix, a, softmax weight, -}
  uint4* tab1 = tab0 + kRunRows * ws;                          // {wy0, wy1, wx0, wx1}
  uint4* tab2 = tab1 + kRunRows * ws;                          // {sy0, sy1, sx0, sx1}, then {d/d attn, d/d x, d/d y}
  float4* gtile = reinterpret_cast<float4*>(tab2 + kRunRows * ws);  // [64][8] grad_out rows of this head
  load_levels(geom, shapes, lstart, L);

  const int sub = threadIdx.x & 7;
  const int m = blockIdx.x % M;
  const int tiles = (Lq + kRunRows - 1) / kRunRows;
  const int tile = blockIdx.x / M;
  const int b = tile / tiles, q0 = (tile % tiles) * kRunRows;
  const int rs4 = M * 32 * 4;
  const long long voff = (long long)b * S * (M * 32) + m * 32 + sub * 4;
  const float* vb = value + voff;
  const float* gb = grad_value + voff;

  for (int rr = threadIdx.x >> 3; rr < kRunRows; rr += blockDim.x >> 3) {
    const long long bq = (long long)b * Lq + min(q0 + rr, Lq - 1);
    const long long row = bq * M + m;
    uint4* my0 = tab0 + rr * ws;
    uint4* my1 = tab1 + rr * ws;
    uint4* my2 = tab2 + rr * ws;
    auto publish = [&](int s, float x, float y, float a, float soft) {
      const Slots t = place(x, y, a, geom[s / kP]);
      my0[s] = make_uint4(unsigned(t.pix), __float_as_uint(t.a), __float_as_uint(soft), 0u);
      my1[s] = make_uint4(__float_as_uint(t.wy0), __float_as_uint(t.wy1), __float_as_uint(t.wx0), __float_as_uint(t.wx1));
      my2[s] = make_uint4(__float_as_uint(t.sy0), __float_as_uint(t.sy1), __float_as_uint(t.sx0), __float_as_uint(t.sx1));
    };
    if constexpr (kMode == 0) {
      for (int s = sub; s < LP; s += 8) {
        const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + row * LP + s);
        publish(s, xy.x, xy.y, __ldg(attn + row * LP + s), 0.f);
      }
    } else {
      const RowTaps t = fused_taps<kMode, kP>(loc + bq * ostride + (long long)m * LP * 2, attn + bq * lstride + (long long)m * LP,
                                              ref, bq, sub, L, LP, geom);
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        if (c * 8 + sub < LP) publish(c * 8 + sub, t.x[c], t.y[c], t.a[c], t.a[c]);
    }
    gtile[rr * 8 + sub] = ldg4(grad_out + row * 32 + sub * 4);
  }
  __syncthreads();

  {
    const int run = threadIdx.x / tpr, tr = threadIdx.x - run * tpr;
    const int slot = tr >> 3;
    const bool used = slot < LP;
    const int s = used ? slot : 0;
    const LevelGeom g = geom[s / kP];
    const bool wide = g.W >= 2;
    const int d01 = wide ? rs4 : 0;
    const long long d10 = g.H >= 2 ? (long long)g.W * rs4 : 0;
    const long long d11 = d10 + d01;
    const float fW = float(g.W), fH = float(g.H);
    const int n = max(0, min(kRunLen, Lq - (q0 + run * kRunLen)));   // padding groups walk along with weight 0
    const int base = (run * kRunLen) * ws + s;
    const float4* grow = gtile + (run * kRunLen) * 8 + sub;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = kFarPixel;
    float4 v00 = zero, v01 = zero, v10 = zero, v11 = zero;
    float4 c00 = zero, c01 = zero, c10 = zero, c11 = zero;   // gradient of the footprint's four pixels
    for (int i = 0; i < n; ++i) {
      const uint4 q0r = tab0[base + i * ws], q1 = tab1[base + i * ws], q2 = tab2[base + i * ws];
      const float4 go = grow[i * 8];
      const int pix = int(q0r.x);
      const float a = used ? __uint_as_float(q0r.y) : 0.f;
      const int d = pix - cur;
      if (d != 0) {
        const bool shift = wide && d == 1;
        const float* gp = pixel_ptr(gb, cur, rs4);
        red_add4_nz(gp, c00);
        red_add4_nz(byte_off(gp, d10), c10);
        const float* p00 = pixel_ptr(vb, pix, rs4);
        if (shift) {
          c00 = c01; c10 = c11; v00 = v01; v10 = v11;
        } else {
          red_add4_nz(byte_off(gp, d01), c01);
          red_add4_nz(byte_off(gp, d11), c11);
          c00 = zero; c10 = zero;
          v00 = ldg4_ordered(p00); v10 = ldg4_ordered(byte_off(p00, d10));
        }
        c01 = zero; c11 = zero;
        v01 = ldg4_ordered(byte_off(p00, d01));
        v11 = ldg4_ordered(byte_off(p00, d11));
        cur = pix;
      }
      const float wy0 = __uint_as_float(q1.x), wy1 = __uint_as_float(q1.y);
      const float wx0 = __uint_as_float(q1.z), wx1 = __uint_as_float(q1.w);
      const float ay0 = wy0 * a, ay1 = wy1 * a;
      const float w00 = ay0 * wx0, w01 = ay0 * wx1, w10 = ay1 * wx0, w11 = ay1 * wx1;
      c00.x = fmaf(w00, go.x, c00.x); c00.y = fmaf(w00, go.y, c00.y); c00.z = fmaf(w00, go.z, c00.z); c00.w = fmaf(w00, go.w, c00.w);
      c01.x = fmaf(w01, go.x, c01.x); c01.y = fmaf(w01, go.y, c01.y); c01.z = fmaf(w01, go.z, c01.z); c01.w = fmaf(w01, go.w, c01.w);
      c10.x = fmaf(w10, go.x, c10.x); c10.y = fmaf(w10, go.y, c10.y); c10.z = fmaf(w10, go.z, c10.z); c10.w = fmaf(w10, go.w, c10.w);
      c11.x = fmaf(w11, go.x, c11.x); c11.y = fmaf(w11, go.y, c11.y); c11.z = fmaf(w11, go.z, c11.z); c11.w = fmaf(w11, go.w, c11.w);
      const float e00 = dot4(go, v00), e01 = dot4(go, v01), e10 = dot4(go, v10), e11 = dot4(go, v11);
      const float r0 = fmaf(wx1, e01, wx0 * e00), r1 = fmaf(wx1, e11, wx0 * e10);
      const float sx0 = __uint_as_float(q2.z), sx1 = __uint_as_float(q2.w);
      const float t0 = fmaf(sx1, e01, sx0 * e00), t1 = fmaf(sx1, e11, sx0 * e10);
      float pa = fmaf(wy1, r1, wy0 * r0);                                                // cuh:156
      float px = fmaf(wy1, t1, wy0 * t0);                                                // cuh:157 (x)
      float py = fmaf(__uint_as_float(q2.y), r1, __uint_as_float(q2.x) * r0);            // cuh:158 (y)
      pa = group8_sum(pa); px = group8_sum(px); py = group8_sum(py);
      // all 8 lanes of the group have read record i (the shuffles are warp-synchronous): reuse its slot of table 2
      if (sub == 0 && used)
        tab2[base + i * ws] = make_uint4(__float_as_uint(pa), __float_as_uint(px * (a * fW)), __float_as_uint(py * (a * fH)), 0u);
    }
    const float* gp = pixel_ptr(gb, cur, rs4);
    red_add4_nz(gp, c00);
    red_add4_nz(byte_off(gp, d01), c01);
    red_add4_nz(byte_off(gp, d10), c10);
    red_add4_nz(byte_off(gp, d11), c11);
  }
  __syncthreads();

  // stage 3: back to one (query, head) row per 8 lanes: softmax / location chain rules, coalesced gradient rows
  for (int rr = threadIdx.x >> 3; rr < kRunRows; rr += blockDim.x >> 3) {
    const int q = q0 + rr;
    const bool live = q < Lq;
    const long long bq = (long long)b * Lq + min(q, Lq - 1);
    const long long row = bq * M + m;
    const uint4* r0 = tab0 + rr * ws;
    const uint4* r2 = tab2 + rr * ws;
    if constexpr (kMode == 0) {
      if (live)
        for (int s = sub; s < LP; s += 8) {
          const uint4 res = r2[s];
          grad_attn[row * LP + s] = __uint_as_float(res.x);
          reinterpret_cast<float2*>(grad_loc)[row * LP + s] = make_float2(__uint_as_float(res.y), __uint_as_float(res.z));
        }
    } else {
      uint4 res[kChunks];
      float soft[kChunks];
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        res[c] = make_uint4(0u, 0u, 0u, 0u);
        soft[c] = 0.f;
        if (c * 8 + sub < LP) { res[c] = r2[c * 8 + sub]; soft[c] = __uint_as_float(r0[c * 8 + sub].z); }
        dot = fmaf(soft[c], __uint_as_float(res[c].x), dot);
      }
      dot = group8_sum(dot);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int s = c * 8 + sub;
        if (live && s < LP) {
          const int l = s / kP;
          float gx = __uint_as_float(res[c].y), gy = __uint_as_float(res[c].z);
          if (kMode == 1) {
            gx = gx / float(geom[l].W);
            gy = gy / float(geom[l].H);
          } else {
            const float4 rr4 = __ldg(reinterpret_cast<const float4*>(ref) + bq * L + l);
            gx = ((gx * 0.5f) * rr4.z) / float(kP);
            gy = ((gy * 0.5f) * rr4.w) / float(kP);
          }
          grad_attn[bq * lstride + (long long)m * LP + s] = soft[c] * (__uint_as_float(res[c].x) - dot);
          *reinterpret_cast<float2*>(grad_loc + bq * ostride + ((long long)m * LP + s) * 2) = make_float2(gx, gy);
        }
      }
    }
  }
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

bool fast_ok(int D, int P, int L, int dtype, const void* a, const void* b, const void* c, const void* d) {
  return dtype == DATR_DTYPE_F32 && D == 32 && (P == 4 || P == 1 || P == 2 || P == 8) && L <= kMaxLevels &&
         L * P <= kMaxTaps && aligned(a, 16) && aligned(b, 16) && aligned(c, 8) && aligned(d, 16);
}

// The backward slot tables of L*P = 32 taps need 51 KB of dynamic shared memory: opt in once per device.
int allow_big_smem() {
  static std::atomic<uint64_t> done{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(DATR_ERR_CUDA, "cudaGetDevice failed%s");
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return DATR_OK;
  const int bytes = 64 * 1024;
  cudaError_t e = cudaSuccess;
#define DATR_OPT(K) if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
#define DATR_OPT_MODES(PP) DATR_OPT((msda_bwd_f32_d32<PP, 0>)); DATR_OPT((msda_bwd_f32_d32<PP, 1>)); DATR_OPT((msda_bwd_f32_d32<PP, 2>))
  DATR_OPT_MODES(1); DATR_OPT_MODES(2); DATR_OPT_MODES(4); DATR_OPT_MODES(8);
#undef DATR_OPT_MODES
#undef DATR_OPT
  if (e != cudaSuccess) return fail(DATR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  done.fetch_or(bit, std::memory_order_release);
  return DATR_OK;
}

// Kernel family of the fp32 / D = 32 path: 0 = auto (run kernels when the queries are the pixels of the value maps,
// i.e. Lq == S: encoder self-attention), 1 = always the row kernels, 2 = run kernels wherever they exist.
// Process-wide tuning knob (datr_msda_set_strategy); initialised from the environment variable DATR_MSDA_STRATEGY.
std::atomic<int> g_strategy{-1};

int strategy() {
  int v = g_strategy.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("DATR_MSDA_STRATEGY");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
    g_strategy.store(v, std::memory_order_relaxed);
  }
  return v;
}

bool use_runs(int mode, int P, int L, int Lq, int S) {
  if (P != 4 || mode == 2 || L * P > kMaxTaps) return false;  // instantiated for the DINO encoder (4 points, 2-d references)
  const int st = strategy();
  return st == 2 || (st == 0 && Lq == S);
}

size_t runs_smem_fwd(int LP) {
  return kGeomBytes + (size_t)kRunRows * table_stride(LP) * 16 + (size_t)kRunsPerCta * (run_threads(LP) / 32) * kRunLen * 128 +
         (size_t)kRunRows * pix_stride(LP) * 4;
}
size_t runs_smem_bwd(int LP) { return kGeomBytes + (size_t)kRunRows * table_stride(LP) * 48 + (size_t)kRunRows * 128; }

int allow_runs_smem() {
  static std::atomic<uint64_t> done{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(DATR_ERR_CUDA, "cudaGetDevice failed%s");
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return DATR_OK;
  const int fb = (int)runs_smem_fwd(kMaxTaps), bb = (int)runs_smem_bwd(kMaxTaps);
  cudaError_t e = cudaFuncSetAttribute(msda_fwd_runs_f32_d32<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(msda_fwd_runs_f32_d32<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(msda_bwd_runs_f32_d32<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bb);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(msda_bwd_runs_f32_d32<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bb);
  if (e != cudaSuccess) return fail(DATR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  done.fetch_or(bit, std::memory_order_release);
  return DATR_OK;
}

long long runs_ctas(int N, int Lq, int M) { return (long long)N * ((Lq + kRunRows - 1) / kRunRows) * M; }

// launchers of the fp32 / D = 32 kernels; mode 0 = materialised locations + weights, 1 / 2 = fused prologue (R = 2 / 4)
int launch_fwd_fast(int mode, const float* v, const int64_t* shapes, const int64_t* lstart, const float* lc,
                    const float* at, const float* ref, long long ostride, long long lstride, int N, int S, int M, int L,
                    int Lq, int P, float* o, cudaStream_t stream) {
  if (use_runs(mode, P, L, Lq, S)) {
    const long long rc_ = runs_ctas(N, Lq, M);
    if (rc_ > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
    if (int rc = allow_runs_smem()) return rc;
    const int threads = kRunsPerCta * run_threads(L * P);
    const size_t sm = runs_smem_fwd(L * P);
    if (mode == 0)
      msda_fwd_runs_f32_d32<4, 0><<<(unsigned)rc_, threads, sm, stream>>>(v, shapes, lstart, lc, at, ref, ostride, lstride, N, S, M, L, Lq, o);
    else
      msda_fwd_runs_f32_d32<4, 1><<<(unsigned)rc_, threads, sm, stream>>>(v, shapes, lstart, lc, at, ref, ostride, lstride, N, S, M, L, Lq, o);
    return after_launch("msda_fwd_runs_f32_d32");
  }
  const long long ctas = (((long long)N * Lq + kRowsPerCta - 1) / kRowsPerCta) * M;
  if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
  const int LP = L * P;
  const size_t smem = kGeomBytes + (size_t)kRowsPerCta * (table_stride(LP) * 16 + pix_stride(LP) * 4);
#define DATR_FWD(PP, BB, MM) \
  msda_fwd_f32_d32<PP, BB, MM><<<(unsigned)ctas, 256, smem, stream>>>(v, shapes, lstart, lc, at, ref, ostride, lstride, N, S, M, L, Lq, o)
#define DATR_FWD_P(MM)                  \
  switch (P) {                          \
    case 1: DATR_FWD(1, 1, MM); break;  \
    case 2: DATR_FWD(2, 2, MM); break;  \
    case 4: DATR_FWD(4, 2, MM); break;  \
    default: DATR_FWD(8, 2, MM); break; \
  }
  if (mode == 0) { DATR_FWD_P(0) } else if (mode == 1) { DATR_FWD_P(1) } else { DATR_FWD_P(2) }
#undef DATR_FWD_P
#undef DATR_FWD
  return after_launch("msda_fwd_f32_d32");
}

int launch_bwd_fast(int mode, const float* v, const int64_t* shapes, const int64_t* lstart, const float* lc,
                    const float* at, const float* ref, const float* go, long long ostride, long long lstride, int N, int S,
                    int M, int L, int Lq, int P, float* gv, float* gl, float* ga, cudaStream_t stream) {
  if (use_runs(mode, P, L, Lq, S)) {
    const long long rc_ = runs_ctas(N, Lq, M);
    if (rc_ > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
    if (int rc = allow_runs_smem()) return rc;
    const int threads = kRunsPerCta * run_threads(L * P);
    const size_t sm = runs_smem_bwd(L * P);
    if (mode == 0)
      msda_bwd_runs_f32_d32<4, 0><<<(unsigned)rc_, threads, sm, stream>>>(v, shapes, lstart, lc, at, ref, go, ostride, lstride, N, S, M, L, Lq, gv, gl, ga);
    else
      msda_bwd_runs_f32_d32<4, 1><<<(unsigned)rc_, threads, sm, stream>>>(v, shapes, lstart, lc, at, ref, go, ostride, lstride, N, S, M, L, Lq, gv, gl, ga);
    return after_launch("msda_bwd_runs_f32_d32");
  }
  const long long ctas = (((long long)N * Lq + kRowsPerCta - 1) / kRowsPerCta) * M;
  if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
  const size_t smem = kGeomBytes + (size_t)kRowsPerCta * table_stride(L * P) * 48;
  if (int rc = allow_big_smem()) return rc;
#define DATR_BWD(PP, MM) \
  msda_bwd_f32_d32<PP, MM><<<(unsigned)ctas, 256, smem, stream>>>(v, shapes, lstart, lc, at, ref, go, ostride, lstride, N, S, M, L, Lq, gv, gl, ga)
#define DATR_BWD_P(MM)               \
  switch (P) {                       \
    case 1: DATR_BWD(1, MM); break;  \
    case 2: DATR_BWD(2, MM); break;  \
    case 4: DATR_BWD(4, MM); break;  \
    default: DATR_BWD(8, MM); break; \
  }
  if (mode == 0) { DATR_BWD_P(0) } else if (mode == 1) { DATR_BWD_P(1) } else { DATR_BWD_P(2) }
#undef DATR_BWD_P
#undef DATR_BWD
  return after_launch("msda_bwd_f32_d32");
}

int check_fused(const void* ref, int ref_dim, int D, int P, int L, int dtype) {
  if (!ref) return fail(DATR_ERR_BAD_ARGUMENT, "null reference_points pointer%s");
  if (ref_dim != 2 && ref_dim != 4) return fail(DATR_ERR_BAD_ARGUMENT, "reference_points must have 2 or 4 components%s");
  if (dtype != DATR_DTYPE_F32 || D != 32 || !(P == 1 || P == 2 || P == 4 || P == 8) || L > kMaxLevels || L * P > kMaxTaps)
    return fail(DATR_ERR_UNSUPPORTED, "the fused entry points cover fp32, 32 channels per head, 1/2/4/8 points, L*P <= 32%s");
  if (!aligned(ref, ref_dim == 2 ? 8 : 16)) return fail(DATR_ERR_ALIGNMENT, "reference_points not aligned to one point%s");
  return DATR_OK;
}

// row strides (elements between consecutive queries) of the offsets / logits tensors; 0 = densely packed
int check_strides(long long* ostride, long long* lstride, int M, int L, int P) {
  const long long taps = (long long)M * L * P;
  if (*ostride == 0) *ostride = taps * 2;
  if (*lstride == 0) *lstride = taps;
  if (*ostride < taps * 2 || *lstride < taps || (*ostride & 1))
    return fail(DATR_ERR_BAD_ARGUMENT, "row strides must cover one row (and the offsets stride must be even)%s");
  return DATR_OK;
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
  if (fast_ok(D, P, L, dtype, value, out, loc, attn))
    return launch_fwd_fast(0, static_cast<const float*>(value), shapes, lstart, static_cast<const float*>(loc),
                           static_cast<const float*>(attn), nullptr, 0, 0, N, S, M, L, Lq, P, static_cast<float*>(out), stream);
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
  if (fast_ok(D, P, L, dtype, value, grad_out, grad_loc, grad_attn) && aligned(grad_value, 16) && aligned(loc, 8) &&
      aligned(attn, 4))
    return launch_bwd_fast(0, static_cast<const float*>(value), shapes, lstart, static_cast<const float*>(loc),
                           static_cast<const float*>(attn), nullptr, static_cast<const float*>(grad_out), 0, 0, N, S, M, L, Lq, P,
                           static_cast<float*>(grad_value), static_cast<float*>(grad_loc), static_cast<float*>(grad_attn),
                           stream);
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

int datr_msda_fused_forward(const void* value, const int64_t* shapes, const int64_t* lstart, const void* offsets,
                            long long ostride, const void* logits, long long lstride, const void* ref, int ref_dim, int N,
                            int S, int M, int D, int L, int Lq, int P, int dtype, void* out, void* stream_) {
  if (int rc = check_common(value, shapes, lstart, offsets, logits, N, S, M, D, L, Lq, P, dtype)) return rc;
  if (!out) return fail(DATR_ERR_BAD_ARGUMENT, "null output pointer%s");
  if (int rc = check_fused(ref, ref_dim, D, P, L, dtype)) return rc;
  if (int rc = check_strides(&ostride, &lstride, M, L, P)) return rc;
  if (!aligned(value, 16) || !aligned(out, 16) || !aligned(offsets, 8))
    return fail(DATR_ERR_ALIGNMENT, "value / output must be 16-byte aligned, offsets 8-byte aligned%s");
  return launch_fwd_fast(ref_dim == 2 ? 1 : 2, static_cast<const float*>(value), shapes, lstart,
                         static_cast<const float*>(offsets), static_cast<const float*>(logits),
                         static_cast<const float*>(ref), ostride, lstride, N, S, M, L, Lq, P, static_cast<float*>(out),
                         static_cast<cudaStream_t>(stream_));
}

int datr_msda_fused_backward(const void* value, const int64_t* shapes, const int64_t* lstart, const void* offsets,
                             long long ostride, const void* logits, long long lstride, const void* ref, int ref_dim,
                             const void* grad_out, int N, int S, int M, int D, int L, int Lq, int P, int dtype,
                             void* grad_value, void* grad_offsets, void* grad_logits, void* stream_) {
  if (int rc = check_common(value, shapes, lstart, offsets, logits, N, S, M, D, L, Lq, P, dtype)) return rc;
  if (!grad_out || !grad_value || !grad_offsets || !grad_logits) return fail(DATR_ERR_BAD_ARGUMENT, "null gradient pointer%s");
  if (int rc = check_fused(ref, ref_dim, D, P, L, dtype)) return rc;
  if (int rc = check_strides(&ostride, &lstride, M, L, P)) return rc;
  if (!aligned(value, 16) || !aligned(grad_out, 16) || !aligned(grad_value, 16) || !aligned(offsets, 8) ||
      !aligned(grad_offsets, 8))
    return fail(DATR_ERR_ALIGNMENT, "value / grad_output / grad_value must be 16-byte aligned, offsets 8-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const cudaError_t me = cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)N * S * M * D, stream);
  if (me != cudaSuccess) return fail(DATR_ERR_CUDA, "cudaMemsetAsync(grad_value): %s", cudaGetErrorString(me));
  return launch_bwd_fast(ref_dim == 2 ? 1 : 2, static_cast<const float*>(value), shapes, lstart,
                         static_cast<const float*>(offsets), static_cast<const float*>(logits),
                         static_cast<const float*>(ref), static_cast<const float*>(grad_out), ostride, lstride, N, S, M, L,
                         Lq, P, static_cast<float*>(grad_value), static_cast<float*>(grad_offsets),
                         static_cast<float*>(grad_logits), stream);
}

int datr_msda_set_strategy(int s) {
  if (s < 0 || s > 2) return fail(DATR_ERR_BAD_ARGUMENT, "strategy must be 0 (auto), 1 (row kernels) or 2 (run kernels)%s");
  g_strategy.store(s, std::memory_order_relaxed);
  return DATR_OK;
}
int datr_msda_get_strategy(void) { return strategy(); }

const char* datr_last_error(void) { return g_err; }
int datr_abi_version(void) { return 1; }
uint64_t datr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
