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
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
  int anchor;               // by << 16 | bx
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
  s.anchor = (by << 16) | bx;
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
// TMA scatter of the backward (kStages > 0 variants; opt-in, NOT the default).  Measured on B200
// (tools/probes/msda_tma_red_probe.cu, profiles/r02w_msda_tma_red_probe.txt): 22.8 M corner lines take 322 us as
// red.global.add.v4.f32 and 321 us as 5.7 M `cp.reduce.async.bulk.tensor` boxes of 2 x 2 pixels -- the wall of the
// scatter is the L2 atomic units (9 TB/s of fp32 adds), not the SM's load/store path.  In the full kernel the staging
// (4 STS.128 + proxy fence + 2 warp barriers per sample, 3 instead of 4 CTAs/SM) costs more than taking the reductions
// out of L1TEX saves: 502 / 513 us (1 / 2 buffers) against 447 us (profiles/r02x_msda_bwd_tma_variants.txt).
// One tensor map per level over grad_value: dims {32 channels, heads, W, H, batch}, box {32, 1, 2, 2, 1}; the four
// weighted rows of a sample are staged in shared memory ([corner][32 floats], the box layout) and lane 0 of the row
// issues ONE reduce per sample.
// ------------------------------------------------------------------------------------------------
constexpr int kTmaLevels = 4;
struct LevelMaps { CUtensorMap m[kTmaLevels]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tma_red_add_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// kStore = 0: value rows are fp32 [N,S,M,32] (the reference's layout).  kStore = 1 / 2: `value` points at PAIR rows
// (datr_msda_pack_value_pairs): the 128-byte line of (pixel, head) holds, per lane, 4 channels of the pixel and the same
// 4 channels of its right-hand neighbour as bf16 (1) / fp16 (2), so a sample costs TWO line gathers instead of four --
// the forward is bound by the L1TEX line rate, not by bytes.
template <int kStore>
__device__ __forceinline__ void unpack_pair(const float4& raw, float4& left, float4& right) {
  const uint32_t a = __float_as_uint(raw.x), b = __float_as_uint(raw.y), c = __float_as_uint(raw.z), d = __float_as_uint(raw.w);
  if constexpr (kStore == 1) {
    left = make_float4(__uint_as_float(a << 16), __uint_as_float(a & 0xffff0000u), __uint_as_float(b << 16),
                       __uint_as_float(b & 0xffff0000u));
    right = make_float4(__uint_as_float(c << 16), __uint_as_float(c & 0xffff0000u), __uint_as_float(d << 16),
                        __uint_as_float(d & 0xffff0000u));
  } else {
    const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&a)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&b));
    const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&c)), r1 = __half22float2(*reinterpret_cast<const __half2*>(&d));
    left = make_float4(l0.x, l0.y, l1.x, l1.y);
    right = make_float4(r0.x, r0.y, r1.x, r1.y);
  }
}

template <int kP, int kBatch, int kMode, int kMinBlocks = 0, int kStore = 0>
__global__ void __launch_bounds__(256, kMinBlocks)
msda_fwd_f32_d32(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                 const float* __restrict__ attn, const float* __restrict__ ref, long long ostride, long long lstride,
                 int N, int S, int M, int L, int Lq, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
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
        if constexpr (kStore == 0) {
          v[j][1] = ldg4_ordered(pixel_ptr(vb01, pix, rs4));
          v[j][2] = ldg4_ordered(pixel_ptr(vb10, pix, rs4));
          v[j][3] = ldg4_ordered(pixel_ptr(vb11, pix, rs4));
        } else {
          v[j][2] = ldg4_ordered(pixel_ptr(vb10, pix, rs4));
        }
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const float wj[4] = {__uint_as_float(w[j].x), __uint_as_float(w[j].y), __uint_as_float(w[j].z),
                             __uint_as_float(w[j].w)};
        if constexpr (kStore != 0) {
          const float4 top = v[j][0], bottom = v[j][2];
          unpack_pair<kStore>(top, v[j][0], v[j][1]);
          unpack_pair<kStore>(bottom, v[j][2], v[j][3]);
        }
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
template <int kP, int kMode, int kStages = 0>   // kStages > 0: TMA scatter with that many staging buffers per warp
__global__ void __launch_bounds__(256, 4)
msda_bwd_f32_d32(const __grid_constant__ LevelMaps maps, const float* __restrict__ value, const int64_t* __restrict__ shapes,
                 const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                 const float* __restrict__ attn, const float* __restrict__ ref, const float* __restrict__ grad_out,
                 long long ostride, long long lstride, int N, int S, int M, int L, int Lq,
                 float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  extern __shared__ __align__(128) unsigned char smem[];
  LevelGeom* geom = reinterpret_cast<LevelGeom*>(smem);
  const int LP = L * kP, ws = table_stride(LP);
  uint4* tab0 = reinterpret_cast<uint4*>(smem + kGeomBytes);   // [32][ws] {pix, a, a*W, a*H}
  uint4* tab1 = tab0 + kRowsPerCta * ws;                       // {wy0, wy1, wx0, wx1}
  uint4* tab2 = tab1 + kRowsPerCta * ws;                       // {sy0, sy1, sx0, sx1}
  constexpr bool kTma = kStages > 0;
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
  // TMA variant: staging [warp][buffer][row of the warp][corner][32 floats] behind the tables, 128-byte aligned
  float* stage = nullptr;
  if constexpr (kTma) {
    const uint32_t tables = kGeomBytes + 3u * kRowsPerCta * ws * 16u;
    stage = reinterpret_cast<float*>(smem + ((tables + 127u) & ~127u)) + (threadIdx.x >> 5) * (kStages * 4 * 128) +
            (r & 3) * 128 + sub * 4;
  }

  auto publish = [&](int s, float x, float y, float a) {
    const LevelGeom g = geom[s / kP];
    const Slots t = place(x, y, a, g);
    const float al = live ? t.a : 0.f;  // dead rows scatter nothing
    if constexpr (kTma)   // {pix, a, anchor (by << 16 | bx) = box coordinates of the reduce, -}
      my0[s] = make_uint4(unsigned(t.pix), __float_as_uint(al), unsigned(t.anchor), 0u);
    else
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
      const float ay0 = wy0 * a, ay1 = wy1 * a;
      if constexpr (kTma) {
        // grad_value: the four weighted rows go to the staging buffer of this (warp, sample parity) and lane 0 of the
        // row hands them to the TMA unit as ONE 2x2-pixel reduce.  The buffer was last used kStages samples ago.
        float* st = stage + (s % kStages) * (4 * 128);
        if (sub == 0) bulk_wait_read<kStages - 1>();
        __syncwarp();
        const float c00 = ay0 * wx0, c01 = ay0 * wx1, c10 = ay1 * wx0, c11 = ay1 * wx1;
        *reinterpret_cast<float4*>(st) = make_float4(c00 * g.x, c00 * g.y, c00 * g.z, c00 * g.w);
        *reinterpret_cast<float4*>(st + 32) = make_float4(c01 * g.x, c01 * g.y, c01 * g.z, c01 * g.w);
        *reinterpret_cast<float4*>(st + 64) = make_float4(c10 * g.x, c10 * g.y, c10 * g.z, c10 * g.w);
        *reinterpret_cast<float4*>(st + 96) = make_float4(c11 * g.x, c11 * g.y, c11 * g.z, c11 * g.w);
        fence_async_smem();
        __syncwarp();
        if (sub == 0) {
          if (a != 0.f) tma_red_add_5d(&maps.m[l], st, 0, m, int(q0.z & 0xffffu), int(q0.z >> 16), b);
          bulk_commit();
        }
      } else {
        // grad_value: slot weight * attn * grad_out, one 16-byte reduction per slot (skipped if weight 0)
        const float* g00 = pixel_ptr(gb, pix, rs4);
        red_add4_if(g00, ay0 * wx0, g);
        red_add4_if(reinterpret_cast<const float*>(reinterpret_cast<const char*>(g00) + d01), ay0 * wx1, g);
        red_add4_if(reinterpret_cast<const float*>(reinterpret_cast<const char*>(g00) + d10), ay1 * wx0, g);
        red_add4_if(reinterpret_cast<const float*>(reinterpret_cast<const char*>(g00) + d11), ay1 * wx1, g);
      }
      // <grad_out, slot value> over this lane's 4 channels
      const float e00 = dot4(g, v00), e01 = dot4(g, v01), e10 = dot4(g, v10), e11 = dot4(g, v11);
      const float r0 = fmaf(wx1, e01, wx0 * e00), r1 = fmaf(wx1, e11, wx0 * e10);       // interpolate along x
      const float sx0 = __uint_as_float(q2.z), sx1 = __uint_as_float(q2.w);
      const float t0 = fmaf(sx1, e01, sx0 * e00), t1 = fmaf(sx1, e11, sx0 * e10);       // differentiate along x
      float pa = fmaf(wy1, r1, wy0 * r0);                                                // cuh:156
      float px = fmaf(wy1, t1, wy0 * t0);                                                // cuh:157 (x)
      float py = fmaf(__uint_as_float(q2.y), r1, __uint_as_float(q2.x) * r0);            // cuh:158 (y)
      pa = group8_sum(pa); px = group8_sum(px); py = group8_sum(py);
      // attn * (W, H): read from table 0, or rebuilt from the same two factors when the TMA variant keeps the anchor there
      const float aW = kTma ? a * float(W) : __uint_as_float(q0.z), aH = kTma ? a * float(H) : __uint_as_float(q0.w);
      if constexpr (kMode == 0) {
        if (sub == (s & 7)) { keep_a = pa; keep_x = px * aW; keep_y = py * aH; }
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
          my2[s] = make_uint4(__float_as_uint(pa), __float_as_uint(px * aW), __float_as_uint(py * aH), 0u);
      }
    }
  }
  if constexpr (kTma) {
    if (sub == 0) bulk_wait_read<0>();   // the staging buffers must outlive the last reads of the TMA unit
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
// Pair rows: [N,S,M,32] fp32 -> [N,S,M,8][left 4 channels, right 4 channels] 16-bit, where `right` is the pixel one
// column further in the same level row (zeros in the last column: the kernels anchor every 2x2 block at x <= W-2).
// ------------------------------------------------------------------------------------------------
template <int kStore>
__device__ __forceinline__ uint2 pack4(const float4& v) {
  if constexpr (kStore == 1) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  } else {
    const __half2 a = __floats2half2_rn(fminf(fmaxf(v.x, -65504.f), 65504.f), fminf(fmaxf(v.y, -65504.f), 65504.f));
    const __half2 b = __floats2half2_rn(fminf(fmaxf(v.z, -65504.f), 65504.f), fminf(fmaxf(v.w, -65504.f), 65504.f));
    return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

template <int kStore>
__global__ void __launch_bounds__(256)
msda_pack_pairs(const float* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                long long total, int S, int M, int L, uint4* __restrict__ pairs) {
  __shared__ LevelGeom geom[kMaxLevels];
  load_levels(geom, shapes, lstart, L);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte output = (n, s, m, lane)
  if (t >= total) return;
  const int per_pixel = M * 8;
  const int s = int((t / per_pixel) % S);
  int W = 1, x = 0;
  for (int l = 0; l < L; ++l)
    if (s >= geom[l].start && s < geom[l].start + geom[l].H * geom[l].W) { W = geom[l].W; x = (s - geom[l].start) % W; }
  const float4 left = __ldg(reinterpret_cast<const float4*>(value) + t);
  float4 right = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x + 1 < W) right = __ldg(reinterpret_cast<const float4*>(value) + t + per_pixel);
  const uint2 a = pack4<kStore>(left), b = pack4<kStore>(right);
  pairs[t] = make_uint4(a.x, a.y, b.x, b.y);
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
  const int bytes = 96 * 1024;
  cudaError_t e = cudaSuccess;
#define DATR_OPT(K) if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
#define DATR_OPT_MODES(PP) DATR_OPT((msda_bwd_f32_d32<PP, 0>)); DATR_OPT((msda_bwd_f32_d32<PP, 1>)); DATR_OPT((msda_bwd_f32_d32<PP, 2>))
  DATR_OPT_MODES(1); DATR_OPT_MODES(2); DATR_OPT_MODES(4); DATR_OPT_MODES(8);
#undef DATR_OPT_MODES
#define DATR_OPT_TMA(SS) DATR_OPT((msda_bwd_f32_d32<4, 0, SS>)); DATR_OPT((msda_bwd_f32_d32<4, 1, SS>)); DATR_OPT((msda_bwd_f32_d32<4, 2, SS>))
  DATR_OPT_TMA(1); DATR_OPT_TMA(2);
#undef DATR_OPT_TMA
#undef DATR_OPT
  // tuning hook: preferred shared-memory carve-out (percent of the 228 KB) of the DINO-configuration kernels
  if (const char* c = getenv("DATR_MSDA_CARVEOUT")) {
    const int pct = atoi(c);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(msda_bwd_f32_d32<4, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(msda_bwd_f32_d32<4, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(msda_bwd_f32_d32<4, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }
  if (e != cudaSuccess) return fail(DATR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  done.fetch_or(bit, std::memory_order_release);
  return DATR_OK;
}

// launchers of the fp32 / D = 32 kernels; mode 0 = materialised locations + weights, 1 / 2 = fused prologue (R = 2 / 4)
int launch_fwd_fast(int mode, const float* v, const int64_t* shapes, const int64_t* lstart, const float* lc,
                    const float* at, const float* ref, long long ostride, long long lstride, int N, int S, int M, int L,
                    int Lq, int P, float* o, cudaStream_t stream, int store = 0) {
  const long long ctas = (((long long)N * Lq + kRowsPerCta - 1) / kRowsPerCta) * M;
  if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
  const int LP = L * P;
  const size_t smem = kGeomBytes + (size_t)kRowsPerCta * (table_stride(LP) * 16 + pix_stride(LP) * 4);
#define DATR_FWD(PP, BB, MM) \
  msda_fwd_f32_d32<PP, BB, MM><<<(unsigned)ctas, 256, smem, stream>>>(v, shapes, lstart, lc, at, ref, ostride, lstride, N, S, M, L, Lq, o)
#define DATR_FWD4(BB, MM, KK, SS) \
  msda_fwd_f32_d32<4, BB, MM, KK, SS><<<(unsigned)ctas, 256, smem, stream>>>(v, shapes, lstart, lc, at, ref, ostride, lstride, N, S, M, L, Lq, o)
  // 4 points (DINO): measured on B200 (profiles/r02d_msda_fwd_variants.txt) -- long calls (encoder) are fastest with
  // 2 samples of loads in flight at 5 CTAs/SM (171 vs 177 us at config 2, 409 vs 433 us at 5 scales), short calls
  // (decoder, a few hundred CTAs per SM wave) with 4 samples in flight at 4 CTAs/SM (19.5 vs 22.5 us)
  const bool short_call = (long long)Lq * 4 < (long long)S;
#define DATR_FWD_P(MM)                  \
  switch (P) {                          \
    case 1: DATR_FWD(1, 1, MM); break;  \
    case 2: DATR_FWD(2, 2, MM); break;  \
    case 4:                             \
      if (store == 1) DATR_FWD4(4, MM, 4, 1); else if (store == 2) DATR_FWD4(4, MM, 4, 2); \
      else if (short_call) DATR_FWD4(4, MM, 4, 0); else DATR_FWD4(2, MM, 5, 0); \
      break;                            \
    default: DATR_FWD(8, 2, MM); break; \
  }
  if (store != 0 && P != 4) return fail(DATR_ERR_UNSUPPORTED, "pair-row value maps are implemented for 4 points%s");
  if (mode == 0) { DATR_FWD_P(0) } else if (mode == 1) { DATR_FWD_P(1) } else { DATR_FWD_P(2) }
#undef DATR_FWD_P
#undef DATR_FWD4
#undef DATR_FWD
  return after_launch("msda_fwd_f32_d32");
}

// Backward scatter variant: 0 = red.global.add.v4.f32 per corner, 1 / 2 = TMA reduce per sample with that many staging
// buffers per warp.  DATR_MSDA_BWD_STAGES overrides the default at load time; datr_msda_set_backward_stages() at run time.
std::atomic<int> g_bwd_stages{-1};
int bwd_stages() {
  int v = g_bwd_stages.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("DATR_MSDA_BWD_STAGES");
    v = e ? atoi(e) : 0;   // measured: 447 us (0) vs 502 / 513 us (1 / 2) at the config-2 encoder call
    if (v < 0 || v > 2) v = 0;
    g_bwd_stages.store(v, std::memory_order_relaxed);
  }
  return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return static_cast<EncodeTiledFn>(nullptr);
  }();
  return fn;
}

// One tensor map per level over grad_value [N, S, M, 32]: dims {32, M, W, H, N}, box {32, 1, 2, 2, 1}.  Returns false if
// the geometry is outside what the TMA variant covers (then the red.global kernel runs).
bool encode_level_maps(LevelMaps* maps, float* gv, const int64_t* hshapes, const int64_t* hstart, int N, int S, int M, int L) {
  EncodeTiledFn enc = encode_fn();
  if (!enc || L > kTmaLevels || M > 256) return false;
  memset(maps, 0, sizeof *maps);
  for (int l = 0; l < L; ++l) {
    const int64_t H = hshapes[2 * l], W = hshapes[2 * l + 1], st = hstart[l];
    if (H < 2 || W < 2 || H > 65535 || W > 65535 || st < 0 || st + H * W > S) return false;
    const cuuint64_t row = (cuuint64_t)M * 32 * 4;
    const cuuint64_t gdim[5] = {32, (cuuint64_t)M, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t gstr[4] = {32 * 4, row, (cuuint64_t)W * row, (cuuint64_t)S * row};
    const cuuint32_t box[5] = {32, 1, 2, 2, 1}, estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&maps->m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, gv + st * (int64_t)M * 32, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {     // no context on this thread yet: bind it, retry
      cudaFree(nullptr);
      r = enc(&maps->m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, gv + st * (int64_t)M * 32, gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return false;
  }
  return true;
}

// `hshapes` / `hstart`: host copies of spatial_shapes / level_start_index (or null): needed to build the tensor maps of the
// TMA scatter; without them the red.global kernel runs.
int launch_bwd_fast(int mode, const float* v, const int64_t* shapes, const int64_t* lstart, const float* lc,
                    const float* at, const float* ref, const float* go, long long ostride, long long lstride, int N, int S,
                    int M, int L, int Lq, int P, float* gv, float* gl, float* ga, cudaStream_t stream,
                    const int64_t* hshapes = nullptr, const int64_t* hstart = nullptr) {
  const long long ctas = (((long long)N * Lq + kRowsPerCta - 1) / kRowsPerCta) * M;
  if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
  size_t smem = kGeomBytes + (size_t)kRowsPerCta * table_stride(L * P) * 48;
  if (int rc = allow_big_smem()) return rc;
  static const LevelMaps no_maps = {};
  const int stages = bwd_stages();
  if (stages > 0 && P == 4 && hshapes && hstart) {
    LevelMaps maps;
    if (encode_level_maps(&maps, gv, hshapes, hstart, N, S, M, L)) {
      smem = ((smem + 127) & ~size_t(127)) + (size_t)8 * stages * 4 * 512;
#define DATR_BWD_TMA(MM, SS) \
  msda_bwd_f32_d32<4, MM, SS><<<(unsigned)ctas, 256, smem, stream>>>(maps, v, shapes, lstart, lc, at, ref, go, ostride, lstride, N, S, M, L, Lq, gv, gl, ga)
#define DATR_BWD_TMA_S(MM) if (stages == 1) DATR_BWD_TMA(MM, 1); else DATR_BWD_TMA(MM, 2)
      if (mode == 0) { DATR_BWD_TMA_S(0); } else if (mode == 1) { DATR_BWD_TMA_S(1); } else { DATR_BWD_TMA_S(2); }
#undef DATR_BWD_TMA_S
#undef DATR_BWD_TMA
      return after_launch("msda_bwd_f32_d32 (TMA scatter)");
    }
  }
#define DATR_BWD(PP, MM) \
  msda_bwd_f32_d32<PP, MM><<<(unsigned)ctas, 256, smem, stream>>>(no_maps, v, shapes, lstart, lc, at, ref, go, ostride, lstride, N, S, M, L, Lq, gv, gl, ga)
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
  return datr_msda_backward_hs(value, shapes, lstart, nullptr, nullptr, loc, attn, grad_out, N, S, M, D, L, Lq, P, dtype,
                               grad_value, grad_loc, grad_attn, stream_);
}

int datr_msda_backward_hs(const void* value, const int64_t* shapes, const int64_t* lstart, const int64_t* host_shapes,
                          const int64_t* host_lstart, const void* loc, const void* attn, const void* grad_out, int N, int S,
                          int M, int D, int L, int Lq, int P, int dtype, void* grad_value, void* grad_loc, void* grad_attn,
                          void* stream_) {
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
                           stream, host_shapes, host_lstart);
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

int datr_msda_pack_value_pairs(const void* value, const int64_t* shapes, const int64_t* lstart, int N, int S, int M, int L,
                               int storage, void* pairs, void* stream_) {
  if (!value || !shapes || !lstart || !pairs) return fail(DATR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (N <= 0 || S <= 0 || M <= 0 || L <= 0 || L > kMaxLevels) return fail(DATR_ERR_BAD_ARGUMENT, "bad dimensions%s");
  if (storage != DATR_STORE_BF16_PAIRS && storage != DATR_STORE_FP16_PAIRS)
    return fail(DATR_ERR_BAD_ARGUMENT, "storage must be DATR_STORE_BF16_PAIRS or DATR_STORE_FP16_PAIRS%s");
  if (!aligned(value, 16) || !aligned(pairs, 16)) return fail(DATR_ERR_ALIGNMENT, "value / pairs must be 16-byte aligned%s");
  const long long total = (long long)N * S * M * 8;
  const long long ctas = (total + 255) / 256;
  if (ctas > 0x7fffffffLL) return fail(DATR_ERR_BAD_ARGUMENT, "grid too large%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (storage == DATR_STORE_BF16_PAIRS)
    msda_pack_pairs<1><<<(unsigned)ctas, 256, 0, stream>>>(static_cast<const float*>(value), shapes, lstart, total, S, M, L,
                                                           static_cast<uint4*>(pairs));
  else
    msda_pack_pairs<2><<<(unsigned)ctas, 256, 0, stream>>>(static_cast<const float*>(value), shapes, lstart, total, S, M, L,
                                                           static_cast<uint4*>(pairs));
  return after_launch("msda_pack_pairs");
}

int datr_msda_fused_forward_pairs(const void* pairs, int storage, const int64_t* shapes, const int64_t* lstart,
                                  const void* offsets, long long ostride, const void* logits, long long lstride,
                                  const void* ref, int ref_dim, int N, int S, int M, int D, int L, int Lq, int P, void* out,
                                  void* stream_) {
  if (int rc = check_common(pairs, shapes, lstart, offsets, logits, N, S, M, D, L, Lq, P, DATR_DTYPE_F32)) return rc;
  if (!out) return fail(DATR_ERR_BAD_ARGUMENT, "null output pointer%s");
  if (storage != DATR_STORE_BF16_PAIRS && storage != DATR_STORE_FP16_PAIRS)
    return fail(DATR_ERR_BAD_ARGUMENT, "storage must be DATR_STORE_BF16_PAIRS or DATR_STORE_FP16_PAIRS%s");
  if (int rc = check_fused(ref, ref_dim, D, P, L, DATR_DTYPE_F32)) return rc;
  if (int rc = check_strides(&ostride, &lstride, M, L, P)) return rc;
  if (!aligned(pairs, 16) || !aligned(out, 16) || !aligned(offsets, 8))
    return fail(DATR_ERR_ALIGNMENT, "pairs / output must be 16-byte aligned, offsets 8-byte aligned%s");
  return launch_fwd_fast(ref_dim == 2 ? 1 : 2, static_cast<const float*>(pairs), shapes, lstart,
                         static_cast<const float*>(offsets), static_cast<const float*>(logits),
                         static_cast<const float*>(ref), ostride, lstride, N, S, M, L, Lq, P, static_cast<float*>(out),
                         static_cast<cudaStream_t>(stream_), storage);
}

int datr_msda_fused_backward(const void* value, const int64_t* shapes, const int64_t* lstart, const void* offsets,
                             long long ostride, const void* logits, long long lstride, const void* ref, int ref_dim,
                             const void* grad_out, int N, int S, int M, int D, int L, int Lq, int P, int dtype,
                             void* grad_value, void* grad_offsets, void* grad_logits, void* stream_) {
  return datr_msda_fused_backward_hs(value, shapes, lstart, nullptr, nullptr, offsets, ostride, logits, lstride, ref, ref_dim,
                                     grad_out, N, S, M, D, L, Lq, P, dtype, grad_value, grad_offsets, grad_logits, stream_);
}

int datr_msda_fused_backward_hs(const void* value, const int64_t* shapes, const int64_t* lstart, const int64_t* host_shapes,
                                const int64_t* host_lstart, const void* offsets, long long ostride, const void* logits,
                                long long lstride, const void* ref, int ref_dim, const void* grad_out, int N, int S, int M,
                                int D, int L, int Lq, int P, int dtype, void* grad_value, void* grad_offsets,
                                void* grad_logits, void* stream_) {
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
                         static_cast<float*>(grad_logits), stream, host_shapes, host_lstart);
}

const char* datr_last_error(void) { return g_err; }
int datr_abi_version(void) { return 2; }
void datr_msda_set_backward_stages(int stages) { g_bwd_stages.store(stages < 0 ? -1 : (stages > 2 ? 2 : stages), std::memory_order_relaxed); }
int datr_msda_get_backward_stages(void) { return bwd_stages(); }
uint64_t datr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
