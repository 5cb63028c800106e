// linear_tf32.cu -- Y = act(X · Wᵀ + bias) + residual on the 5th-generation tensor cores (sm_100a).
//
// The dense contractions of the DINO transformer -- value / offset / attention-weight / output projections of
// MSDeformAttn (reference models/dino/ops/modules/ms_deform_attn.py:94-125), the encoder/decoder FFN
// (models/dino/deformable_transformer.py:784-805, :941-947) and the two-stage heads -- are nn.Linear layers
// over M = batch * tokens (44 446 at 1333x800, batch 2) rows with K = 256 / 2048.  The reference runs them through
// cuBLAS SIMT fp32; this kernel is the B200-native replacement:
//
//   * operands stay fp32 in HBM (no cast pass, autograd sees ordinary fp32 tensors); TMA (cp.async.bulk.tensor,
//     128-byte swizzle, TFLOAT32 tensor maps = round-to-nearest on load) stages 128 x 32 tiles of X and BN x 32 tiles
//     of W into a shared-memory ring guarded by full/empty mbarriers;
//   * one elected thread issues tcgen05.mma.kind::tf32 (M = 128, N = BN, K = 8 per instruction, both operands
//     K-major straight from the swizzled tiles); the fp32 accumulator lives in tensor memory (BN columns);
//   * four epilogue warps read the accumulator back with tcgen05.ld (32 lanes x 32 columns per instruction),
//     transpose each 32 x 32 chunk through padded shared memory, add the bias, apply ReLU, add the residual and
//     store whole 128-byte rows -- the bias/activation/residual passes of the reference (three extra reads + writes
//     of the [M, N] activation) never touch HBM;
//   * persistent CTAs (one per SM) walk the output tiles; the accumulator is double-buffered in TMEM so the epilogue
//     of one tile overlaps the TMA/MMA main loop of the next;
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue (a warp may only
//     touch TMEM lanes 32*(warp%4) .. +31, so four consecutive warps cover the 128 accumulator rows; two such groups
//     split the tile's columns).
//
// Numerics: TF32 products (10-bit mantissa), fp32 accumulation: ~3e-4 relative on K = 256 contractions, inside the
// 1e-2 reduced-precision bar of BASELINE.json (the strict-fp32 parity tests keep cuBLAS SIMT fp32).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "datr_linear.h"
#include "tcgen05_common.cuh"

namespace {

thread_local char g_lin_err[512] = "";
std::atomic<uint64_t> g_lin_launches{0};

int lfail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_lin_err, sizeof g_lin_err, fmt, detail);
  return code;
}

using namespace datr_tc;
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kEpiWarps = 8;

// four consecutive outputs as bf16 (round to nearest even): element offset `off` of a bf16 matrix that starts at `y`
__device__ __forceinline__ void store_bf16x4(float* y, size_t off, const float4& t) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(t.x, t.y), b = __floats2bfloat162_rn(t.z, t.w);
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(y) + off) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// MN-major 32-bit B tile: {32 columns x 32 rows} boxes 4096 bytes apart (leading byte offset), swizzle atom of 4 rows
// (512 bytes = stride byte offset), layout type SWIZZLE_128B_BASE32B, descriptor version 1 (see wgrad_tf32.cu)
__device__ __forceinline__ uint64_t mnmajor_b_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(4096 >> 4) << 16) | (uint64_t(512 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(1) << 61);
}

constexpr int kStoreTile = 8192;   // bf16-output epilogue: per warp, up to two {64 columns x 32 rows} boxes of 4 KB

template <int BN, int STAGES, bool kBF16 = false>
struct Smem {
  static constexpr int kA = BM * BK * 4, kB = BN * BK * 4, kStage = kA + kB;
  static constexpr int kEpi = kEpiWarps * (kBF16 ? kStoreTile : kStageTile);
  static constexpr int kBars = 1024;  // barriers + TMEM slot
  static constexpr int kTotal = STAGES * kStage + kEpi + kBars + 1024 /* alignment slack */;
};

// ---------------------------------------------------------------------------------------------------------------
// kernel: persistent, one CTA per SM, static round-robin over 128 x BN output tiles (tiles that share an X row
// block are adjacent in the order, so concurrently running CTAs hit the same X tile in L2).  The accumulator is
// double-buffered in tensor memory (2 x BN columns): the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
// ---------------------------------------------------------------------------------------------------------------
// kBF16: operands are bf16 in HBM (64 elements per 128-byte swizzle row, tcgen05.mma.kind::f16 with K = 16: the same
// bytes per pipeline stage feed twice the FLOPs, which is what the shared-memory-port-bound TF32 loop lacks); accumulation
// stays fp32.  `flags` bit 0: the output is written as bf16; bit 1: `residual` (the ReLU-mask source of relu == 3) is bf16.
// kBT (TF32 only): the weight operand is given TRANSPOSED, w_t [K, N] row-major -- what the input gradient of a Linear
// needs (dx = dy . W with W stored [N_fwd, K_fwd]): its tiles are MN-major B operands (TMA boxes of {32 columns x 32 rows}
// with the 32-byte-atom swizzle, transposed-operand bit of the instruction descriptor, as in wgrad_tf32.cu), so the
// backward no longer materialises W^T with a copy kernel per layer and step.
template <int BN, int STAGES, bool kBF16 = false, bool kBT = false, bool kMask = false>
__global__ void __launch_bounds__(kThreads, 1)
linear_tf32_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                   const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ y,
                   int M, int N, int K, int relu, int flags, const __grid_constant__ CUtensorMap tma_y,
                   const __grid_constant__ CUtensorMap tma_r) {
  using L = Smem<BN, STAGES, kBF16>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi = reinterpret_cast<float*>(smem + STAGES * L::kStage);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage + L::kEpi);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* res_full = acc_empty + 2;      // [kEpiWarps] bf16-output epilogue: mask tile of the warp has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BKe = kBF16 ? 64 : BK;       // operand elements per 128-byte row
  const int kblocks = K / BKe;
  const bool out_bf16 = (flags & 1) != 0, res_bf16 = (flags & 2) != 0;
  const int n_tiles = (N + BN - 1) / BN, m_tiles = (M + BM - 1) / BM;
  const int tiles = n_tiles * m_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_w) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, kEpiWarps); }
    for (int s = 0; s < kEpiWarps; ++s) mbar_init(res_full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(full + s, L::kStage);
          unsigned char* a = smem + s * L::kStage;
          // concurrently running tiles start their K sweep at different k-blocks, so that the CTAs of a wave do not
          // all pull the same W lines out of the same L2 slices at the same time (the sum is order-independent)
          const int kk = (kb + tile) % kblocks;
          tma_load_2d(a, &tma_x, kk * BKe, m0, full + s);
          if constexpr (kBT) {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) tma_load_2d(a + L::kA + c * 4096, &tma_w, n0 + 32 * c, kk * BK, full + s);
          } else {
            tma_load_2d(a + L::kA, &tma_w, kk * BKe, n0, full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = kBF16 ? bf16_idesc<BN>() : (kBT ? (tf32_idesc<BN>() | (1u << 16)) : tf32_idesc<BN>());
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
        const uint32_t as = ti & 1;
        mbar_wait(acc_empty + as, ((ti >> 1) & 1) ^ 1);     // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          mbar_wait(full + s, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a = smem_u32(smem + s * L::kStage);
          const uint64_t ad = kmajor_sw128_desc(a);
          const uint64_t bd = kBT ? mnmajor_b_desc(a + L::kA) : kmajor_sw128_desc(a + L::kA);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) { // +32 bytes along K inside the swizzle row = +2 in the address field
            if constexpr (kBF16) umma_bf16(tmem_d, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0);
            else if constexpr (kBT) umma_tf32(tmem_d, ad + uint64_t(k * 2), bd + uint64_t(k * 64), idesc, (kb | k) != 0);   // 8 rows of w_t = +1024 bytes
            else umma_tf32(tmem_d, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(empty + s);                // frees the stage once these MMAs have read it
        }
        umma_commit(acc_full + as);              // accumulator complete
      }
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 and one half of the tile's columns.  Each 32 x 32
    // accumulator chunk goes through a padded shared-memory tile so that global stores (and residual loads) are whole
    // 128-byte rows; the residual rows of the next chunk are fetched while the current one is processed.
    const int lane_base = (warp & 3) * 32;
    const int half = (warp - 2) >> 2;                        // warps 2-5: columns [0, BN/2), warps 6-9: [BN/2, BN)
    if constexpr (kBF16) {
      if (out_bf16) {
        // bf16 output: every thread owns one accumulator row; 32 columns = 64 bytes go straight into a 128-byte-swizzled
        // {64 columns x 32 rows} box (16-byte chunk c of row r at chunk c ^ (r & 7): the 8 threads of a quarter-warp cover
        // all 32 banks, no transpose needed) and the TMA unit writes whole lines -- per-lane 8-byte stores occupy the
        // store path like 16-byte ones and made this epilogue slower than the fp32 one.  The ReLU-backward mask (relu == 3)
        // arrives the same way: the saved bf16 activation tile is TMA-loaded into the box before the accumulator is ready.
        constexpr int kGroups = BN / 128;                    // 64-column groups per warp
        unsigned char* stg = reinterpret_cast<unsigned char*>(epi) + (warp - 2) * kStoreTile;
        uint64_t* rbar = res_full + (warp - 2);
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
          const int m0 = (tile / n_tiles) * BM + lane_base, n0 = (tile % n_tiles) * BN + half * (BN / 2);
          const uint32_t as = ti & 1;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // last tile's stores have read the boxes
          __syncwarp();
          if (relu == 3 && lane == 0) {
            mbar_expect_tx(rbar, kGroups * 4096);
#pragma unroll
            for (int g = 0; g < kGroups; ++g) tma_load_2d(stg + g * 4096, &tma_r, n0 + g * 64, m0, rbar);
          }
          mbar_wait(acc_full + as, (ti >> 1) & 1);
          tc_fence_after();
          if (relu == 3) mbar_wait(rbar, ti & 1);
          const uint32_t tmem_d = tmem_base + as * BN + half * (BN / 2) + (uint32_t(lane_base) << 16);
#pragma unroll
          for (int c = 0; c < BN / 2; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_d + uint32_t(c), v);
            if (c + 32 >= BN / 2) {                           // last chunk read: hand the buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_empty + as);
            }
            unsigned char* rowp = stg + (c >> 6) * 4096 + lane * 128;
            const int chunk0 = (c & 32) ? 4 : 0;               // first 16-byte chunk of these 32 columns inside the 128-byte row
#pragma unroll
            for (int j = 0; j < 4; ++j) {                     // 8 columns = one 16-byte chunk of bf16
              float t[8];
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (bias) {
                b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + 8 * j));
                b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + 8 * j + 4));
              }
              t[0] = __uint_as_float(v[8 * j]) + b0.x; t[1] = __uint_as_float(v[8 * j + 1]) + b0.y;
              t[2] = __uint_as_float(v[8 * j + 2]) + b0.z; t[3] = __uint_as_float(v[8 * j + 3]) + b0.w;
              t[4] = __uint_as_float(v[8 * j + 4]) + b1.x; t[5] = __uint_as_float(v[8 * j + 5]) + b1.y;
              t[6] = __uint_as_float(v[8 * j + 6]) + b1.z; t[7] = __uint_as_float(v[8 * j + 7]) + b1.w;
              uint4* slot = reinterpret_cast<uint4*>(rowp + (((chunk0 + j) ^ (lane & 7)) << 4));
              if (relu == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = fmaxf(t[e], 0.f);
              } else if (relu == 3) {                         // keep where the saved activation is positive (bf16: sign bit clear, not zero)
                const uint4 hm = *slot;
                const uint32_t hw[4] = {hm.x, hm.y, hm.z, hm.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const uint32_t hb = (e & 1) ? (hw[e >> 1] >> 16) : (hw[e >> 1] & 0xffffu);
                  t[e] = (hb != 0u && (hb & 0x8000u) == 0u) ? t[e] : 0.f;
                }
              }
              const __nv_bfloat162 p0 = __floats2bfloat162_rn(t[0], t[1]), p1 = __floats2bfloat162_rn(t[2], t[3]);
              const __nv_bfloat162 p2 = __floats2bfloat162_rn(t[4], t[5]), p3 = __floats2bfloat162_rn(t[6], t[7]);
              *slot = make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                                 *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int g = 0; g < kGroups; ++g)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&tma_y), "r"(smem_u32(stg + g * 4096)), "r"(n0 + g * 64), "r"(m0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA exits
        goto epilogue_done;
      }
    }
    float* tile_s = epi + (warp - 2) * (kStageTile / 4);
    const int tr = lane >> 3, tc = (lane & 7) * 4;           // transposed role: row tr + 4*j, columns tc .. tc+3
    constexpr int kCols = BN / 2;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN + half * kCols;
      const uint32_t as = ti & 1;
      const int row0 = m0 + lane_base + tr;
      float4 res[8];
      [[maybe_unused]] float4 msk[kMask ? 8 : 1];
      auto fetch_residual = [&](int c) {
        const int col = n0 + c + tc;
        if constexpr (kMask) {                                // the mask rows travel with the residual rows, one chunk ahead
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = row0 + 4 * j;
            msk[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < M && col + 4 <= N) msk[j] = __ldg(reinterpret_cast<const float4*>(bias + (size_t)row * N + col));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int row = row0 + 4 * j;
          res[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (residual && row < M && col + 4 <= N) {
            if (res_bf16) {
              const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(residual) + (size_t)row * N + col));
              res[j] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                                   __uint_as_float(u.y & 0xffff0000u));
            } else {
              res[j] = __ldg(reinterpret_cast<const float4*>(residual + (size_t)row * N + col));
            }
          }
        }
      };
      fetch_residual(0);                                      // zeros when there is no residual
      mbar_wait(acc_full + as, (ti >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN + half * kCols + (uint32_t(lane_base) << 16);
#pragma unroll 1
      for (int c = 0; c < kCols; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_d + uint32_t(c), v);
        if (c + 32 >= kCols) {                                // last chunk read: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + as);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(tile_s + lane * kStagePitch + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int col = n0 + c + tc;
        float4 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = *reinterpret_cast<const float4*>(tile_s + (tr + 4 * j) * kStagePitch + tc);
        __syncwarp();
        if (col < N) {
          const bool vec = col + 4 <= N;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && !kMask) {                               // kMask (relu == 4): `bias` carries the [M, N] ReLU-mask source
            if (vec) b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
            else { b4.x = __ldg(bias + col); if (col + 1 < N) b4.y = __ldg(bias + col + 1); if (col + 2 < N) b4.z = __ldg(bias + col + 2); }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = row0 + 4 * j;
            if (row >= M) break;
            float4 t = o[j];
            t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w;
            if (relu == 1) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
            const size_t off = (size_t)row * N + col;
            if (vec && relu == 3) {          // ReLU-backward mask taken from the `residual` tensor (the saved activation)
              t.x = res[j].x > 0.f ? t.x : 0.f; t.y = res[j].y > 0.f ? t.y : 0.f;
              t.z = res[j].z > 0.f ? t.z : 0.f; t.w = res[j].w > 0.f ? t.w : 0.f;
              if (out_bf16) store_bf16x4(y, off, t); else *reinterpret_cast<float4*>(y + off) = t;
            } else if (vec) {
              t.x += res[j].x; t.y += res[j].y; t.z += res[j].z; t.w += res[j].w;
              if (relu == 2) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
              if constexpr (kMask) {         // (x W + residual) kept where the mask source is positive
                const float4 m4 = msk[j];
                t.x = m4.x > 0.f ? t.x : 0.f; t.y = m4.y > 0.f ? t.y : 0.f;
                t.z = m4.z > 0.f ? t.z : 0.f; t.w = m4.w > 0.f ? t.w : 0.f;
              }
              if (out_bf16) store_bf16x4(y, off, t); else *reinterpret_cast<float4*>(y + off) = t;
            } else {
              const float ov[4] = {t.x, t.y, t.z, t.w};
              for (int e = 0; e < 4 && col + e < N; ++e) {
                const float r1 = residual ? __ldg(residual + off + e) : 0.f;
                float o1 = relu == 3 ? (r1 > 0.f ? ov[e] : 0.f) : ov[e] + r1;
                if (kMask && !(__ldg(bias + off + e) > 0.f)) o1 = 0.f;
                y[off + e] = relu == 2 ? fmaxf(o1, 0.f) : o1;
              }
            }
          }
        }
        if ((residual || kMask) && c + 32 < kCols) fetch_residual(c + 32);
      }
    }
  }
epilogue_done:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// [rows, cols] fp32 row-major matrix -> tensor map with a (box_rows x 32 columns) box, 128-byte swizzle, zero fill
int make_map(CUtensorMap* map, const void* base, int rows, int cols, int box_rows, bool bf16 = false) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return lfail(DATR_LINEAR_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(cols) * (bf16 ? 2 : 4)};
  const cuuint32_t box[2] = {cuuint32_t(bf16 ? 64 : BK), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2,
                         const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_lin_err, sizeof g_lin_err, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return DATR_LINEAR_ERR_CUDA;
  }
  return DATR_LINEAR_OK;
}

template <int BN, int STAGES, bool kBF16 = false, bool kBT = false, bool kMask = false>
int launch(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, const float* residual, float* y, int M, int N,
           int K, int relu, cudaStream_t stream, int flags = 0, const CUtensorMap* my = nullptr, const CUtensorMap* mr = nullptr) {
  using L = Smem<BN, STAGES, kBF16>;
  static const CUtensorMap no_map = {};
  static std::atomic<uint64_t> opted{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(opted.load(std::memory_order_acquire) & bit)) {
    const cudaError_t e = cudaFuncSetAttribute(linear_tf32_kernel<BN, STAGES, kBF16, kBT, kMask>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return lfail(DATR_LINEAR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    opted.fetch_or(bit, std::memory_order_release);
  }
  static std::atomic<int> sm_count[64];
  int sms = sm_count[dev & 63].load(std::memory_order_relaxed);
  if (sms == 0) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    sm_count[dev & 63].store(sms, std::memory_order_relaxed);
  }
  const long long tiles = (long long)((N + BN - 1) / BN) * ((M + BM - 1) / BM);
  const unsigned grid = unsigned(tiles < sms ? tiles : sms);
  linear_tf32_kernel<BN, STAGES, kBF16, kBT, kMask><<<grid, kThreads, L::kTotal, stream>>>(mx, mw, bias, residual, y, M, N, K, relu, flags, my ? *my : no_map,
                                                                               mr ? *mr : no_map);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return lfail(DATR_LINEAR_ERR_CUDA, "linear_tf32_kernel launch: %s", cudaGetErrorString(e));
  g_lin_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_LINEAR_OK;
}

}  // namespace

extern "C" {

int datr_linear_tf32(const float* x, const float* w, const float* bias, const float* residual, float* y, int M, int N,
                     int K, int relu, void* stream_) {
  if (!x || !w || !y) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (M <= 0 || N <= 0 || K <= 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (K % BK != 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "K must be a multiple of 32%s");
  if (N % 4 != 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "N must be a multiple of 4%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(x) || !al16(w) || !al16(y) || (bias && !al16(bias)) || (residual && !al16(residual)))
    return lfail(DATR_LINEAR_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUtensorMap mx, mw;
  static const int force_bn = getenv("DATR_LINEAR_BN") ? atoi(getenv("DATR_LINEAR_BN")) : 0;   // tuning hook
  // 128-wide tiles fill the 148 SMs better at N <= 256 and waste no MMA columns when N mod 256 is in (0, 128]
  const bool wide = force_bn ? force_bn == 256 : (N > 256 && (N % 256 == 0 || N % 256 > 128));
  if (int rc = make_map(&mx, x, M, K, BM)) return rc;
  if (int rc = make_map(&mw, w, N, K, wide ? 256 : 128)) return rc;
  return wide ? launch<256, 3>(mx, mw, bias, residual, y, M, N, K, relu, stream)
              : launch<128, 5>(mx, mw, bias, residual, y, M, N, K, relu, stream);
}

// y = act(x . w_t + bias) + residual with the weight given transposed, w_t [K, N] row-major (TF32 products): the input
// gradient of a Linear without a transposed copy of its weight.
static int linear_bt(const float* x, const float* w_t, const float* bias, const float* residual, float* y, int M, int N,
                     int K, int relu, void* stream_);

int datr_linear_tf32_bt(const float* x, const float* w_t, const float* bias, const float* residual, float* y, int M, int N,
                        int K, int relu, void* stream_) {
  if (relu < 0 || relu > 3) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "relu must be 0..3%s");
  return linear_bt(x, w_t, bias, residual, y, M, N, K, relu, stream_);
}

// y = (x . w_t + residual) where mask > 0, else 0 (mask [M, N]; residual optional): the input gradient of a layer whose
// input is a ReLU output with further consumers -- their gradient arrives as `residual`, the ReLU's own backward mask (its
// output is the layer's saved input) is applied here, so neither the accumulation nor the mask is a separate pass.
int datr_linear_tf32_bt_masked(const float* x, const float* w_t, const float* residual, const float* mask, float* y, int M,
                               int N, int K, void* stream_) {
  if (!mask) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "null mask%s");
  return linear_bt(x, w_t, mask, residual, y, M, N, K, 4, stream_);      // the kernel reads the mask through `bias`
}

static int linear_bt(const float* x, const float* w_t, const float* bias, const float* residual, float* y, int M, int N,
                     int K, int relu, void* stream_) {
  if (!x || !w_t || !y) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (M <= 0 || N <= 0 || K <= 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (K % BK != 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "K must be a multiple of 32%s");
  if (N % 4 != 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "N must be a multiple of 4%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(x) || !al16(w_t) || !al16(y) || (bias && !al16(bias)) || (residual && !al16(residual)))
    return lfail(DATR_LINEAR_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EncodeTiledFn enc = encode_fn();
  if (!enc) return lfail(DATR_LINEAR_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  CUtensorMap mx, mw;
  if (int rc = make_map(&mx, x, M, K, BM)) return rc;
  const cuuint64_t gdim[2] = {cuuint64_t(N), cuuint64_t(K)};
  const cuuint64_t gstride[1] = {cuuint64_t(N) * 4};
  const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
  const CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(w_t), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_lin_err, sizeof g_lin_err, "cuTensorMapEncodeTiled (transposed weight) failed (CUresult %d)", int(r));
    return DATR_LINEAR_ERR_CUDA;
  }
  const bool wide = N > 256 && (N % 256 == 0 || N % 256 > 128);
  if (relu == 4)
    return wide ? launch<256, 3, false, true, true>(mx, mw, bias, residual, y, M, N, K, relu, stream)
                : launch<128, 5, false, true, true>(mx, mw, bias, residual, y, M, N, K, relu, stream);
  return wide ? launch<256, 3, false, true>(mx, mw, bias, residual, y, M, N, K, relu, stream)
              : launch<128, 5, false, true>(mx, mw, bias, residual, y, M, N, K, relu, stream);
}

// bf16 operands (x [M, K], w [N, K] as bf16), fp32 accumulation; y fp32 or bf16 (y_bf16); `residual` [M, N] fp32, or bf16
// (residual_bf16) when it is the ReLU-mask source of relu == 3.
int datr_linear_bf16(const void* x, const void* w, const float* bias, const void* residual, int residual_bf16, void* y,
                     int y_bf16, int M, int N, int K, int relu, void* stream_) {
  if (!x || !w || !y) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (M <= 0 || N <= 0 || K <= 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (K % 64 != 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "K must be a multiple of 64 for bf16 operands%s");
  if (N % 4 != 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "N must be a multiple of 4%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(x) || !al16(w) || !al16(y) || (bias && !al16(bias)) || (residual && !al16(residual)))
    return lfail(DATR_LINEAR_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool wide = N > 256 && (N % 256 == 0 || N % 256 > 128);
  if (y_bf16) {
    // bf16 outputs leave through TMA stores of {64 columns x 32 rows} boxes; the only residual they take is the bf16 ReLU mask
    if (N % (wide ? 256 : 128) != 0) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "bf16 outputs need N to be a multiple of the tile width (128 / 256)%s");
    if (residual && !(relu == 3 && residual_bf16))
      return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "bf16 outputs take no residual except the bf16 ReLU mask of relu == 3%s");
    if (relu == 3 && !residual) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "relu == 3 needs the saved activation%s");
    if (relu != 0 && relu != 1 && relu != 3) return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "bf16 outputs support relu 0, 1 and 3%s");
  } else if (residual_bf16) {
    return lfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "a bf16 residual is the ReLU mask of a bf16 output only%s");
  }
  CUtensorMap mx, mw;
  if (int rc = make_map(&mx, x, M, K, BM, true)) return rc;
  if (int rc = make_map(&mw, w, N, K, wide ? 256 : 128, true)) return rc;
  const int flags = (y_bf16 ? 1 : 0) | (residual_bf16 ? 2 : 0);
  CUtensorMap my, mr;
  const CUtensorMap *pmy = nullptr, *pmr = nullptr;
  if (y_bf16) {
    if (int rc = make_map(&my, y, M, N, 32, true)) return rc;
    pmy = &my;
    if (residual) {
      if (int rc = make_map(&mr, residual, M, N, 32, true)) return rc;
      pmr = &mr;
    }
  }
  return wide ? launch<256, 3, true>(mx, mw, bias, static_cast<const float*>(residual), static_cast<float*>(y), M, N, K, relu, stream, flags, pmy, pmr)
              : launch<128, 5, true>(mx, mw, bias, static_cast<const float*>(residual), static_cast<float*>(y), M, N, K, relu, stream, flags, pmy, pmr);
}

const char* datr_linear_last_error(void) { return g_lin_err; }
uint64_t datr_linear_launch_count(void) { return g_lin_launches.load(std::memory_order_relaxed); }

}  // extern "C"
