// attn_fused.cu -- decoder self-attention as ONE tensor-core kernel (sm_100a): S = Q K^T, masked softmax, O = P V with
// the score tile living in tensor memory only.
//
// Reference: nn.MultiheadAttention of the decoder layer (models/dino/deformable_transformer.py:880-908) over the 900
// matching + <= 200 de-noising queries, 8 heads x 32 channels, boolean attention mask (dn_components.py:105-121).  The
// round-1 path (attention.py) kept the [N*H, T, T] score matrix in HBM between two library batched GEMMs and a softmax
// kernel; here a CTA owns 128 queries of one (image, head):
//   * Q tile (128 x 32) and 128-key K / V tiles arrive by TMA straight from the packed in-projection output
//     ([N, T, 2C] for q|k, [N, T, C] for v: row-strided 3-D tensor maps, no head-split copies);
//   * S = Q K^T: four tcgen05.mma.kind::tf32 (M 128, N 128, K 8), both operands K-major from shared memory, accumulator =
//     128 TMEM columns; the four softmax warps (thread = query row = TMEM lane) read it with tcgen05.ld;
//   * two sweeps over the keys: sweep 1 keeps the running row maximum and row sum (exp2 domain), sweep 2 recomputes S,
//     turns it into normalised probabilities, writes them back OVER the scores with tcgen05.st and the tensor core takes
//     them from tensor memory as the A operand of O += P V (V tile = MN-major B operand, 16 MMAs of K 8, N 32);
//   * the mask comes bit-packed (one uint4 per row and key tile; keys beyond T are packed as blocked);
//   * outputs: O [N, T, H*32] (what out_proj consumes -- no transpose copy), the log-sum-exp per row (for a fused
//     backward) and, on request, the probabilities [N*H, T, T] for the GEMM-based backward of attention.py.
// 80 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM overlap each other's softmax and MMA phases.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_attn.h"
#include "tcgen05_common.cuh"

namespace {

using namespace datr_tc;

thread_local char g_af_err[512] = "";
std::atomic<uint64_t> g_af_launches{0};

int ffail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_af_err, sizeof g_af_err, fmt, detail);
  return code;
}

constexpr int kD = 32;                    // channels per head
constexpr int kTile = 128;                // queries per CTA = keys per tile
constexpr int kTileBytes = kTile * kD * 4;
constexpr int kThreadsFwd = 192;          // warps 0-3 softmax, warp 4 TMA, warp 5 MMA
constexpr uint32_t kTmemCols = 256;       // S / P: columns [0, 128), O: [128, 160)
constexpr int kSmemFwd = 5 * kTileBytes + 1024 + 1024;   // Q + 2 K + 2 V + barriers + alignment slack

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// D[tmem] (+)= A[tmem] . B[smem]: the A operand (rows = TMEM lanes, K = consecutive 32-bit columns) is read from tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major 32-bit operand tile of {32 columns x 32k rows} boxes written by TMA with the 128-byte swizzle / 32-byte atoms
// (see wgrad_tf32.cu): rows of 128 bytes, swizzle atom = 4 rows, descriptor version 1, layout type SWIZZLE_128B_BASE32B.
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(4096 >> 4) << 16) | (uint64_t(512 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(1) << 61);
}

// D = fp32, A = B = TF32; S: both K-major, N = 128;  PV: A from tensor memory (K-major), B MN-major (bit 16), N = 32
constexpr uint32_t kIdescS = tf32_idesc<kTile>();
constexpr uint32_t kIdescPV = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | (uint32_t(kD >> 3) << 17) | (uint32_t(BM >> 4) << 24);

__global__ void __launch_bounds__(kThreadsFwd, 2)
attn_fwd_tf32(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
              const __grid_constant__ CUtensorMap map_v, const uint4* __restrict__ mask_bits, int mask_words,
              int H, int T, float scale_log2, float* __restrict__ out, float* __restrict__ lse, float* __restrict__ p_out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = smem;
  unsigned char* k_s = smem + kTileBytes;          // 2 stages
  unsigned char* v_s = smem + 3 * kTileBytes;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // MMA -> softmax: score tile complete
  uint64_t* s_done = bars + 10;   // softmax -> MMA: score tile consumed (sweep 1) / probabilities written (sweep 2)
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, n = blockIdx.z;
  const int nk = (T + kTile - 1) / kTile;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
    mbar_init(s_full, 1);
    mbar_init(s_done, 4);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + kTile;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(q_s, &map_q, h * kD, q0, n, q_full);
      for (int it = 0; it < 2 * nk; ++it) {                       // K tiles of both sweeps
        const int s = it & 1, j = it % nk;
        mbar_wait(k_empty + s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(k_full + s, kTileBytes);
        tma_load_3d(k_s + s * kTileBytes, &map_k, h * kD, j * kTile, n, k_full + s);
        if (it >= nk) {                                           // V tiles of sweep 2: four {32 x 32} boxes
          const int jj = it - nk, vs = jj & 1;
          mbar_wait(v_empty + vs, ((jj >> 1) & 1) ^ 1);
          mbar_expect_tx(v_full + vs, kTileBytes);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tma_load_3d(v_s + vs * kTileBytes + c * 4096, &map_v, h * kD, j * kTile + 32 * c, n, v_full + vs);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      const uint64_t qd = kmajor_sw128_desc(smem_u32(q_s));
      for (int it = 0; it < 2 * nk; ++it) {
        const int s = it & 1;
        mbar_wait(k_full + s, (it >> 1) & 1);
        tc_fence_after();
        const uint64_t kd = kmajor_sw128_desc(smem_u32(k_s + s * kTileBytes));
#pragma unroll
        for (int kk = 0; kk < kD / UMMA_K; ++kk) umma_tf32(tmem_s, qd + uint64_t(kk * 2), kd + uint64_t(kk * 2), kIdescS, kk != 0);
        umma_commit(k_empty + s);
        umma_commit(s_full);
        mbar_wait(s_done, it & 1);                                // scores read (sweep 1) / probabilities in place (sweep 2)
        tc_fence_after();
        if (it >= nk) {
          const int jj = it - nk, vs = jj & 1;
          mbar_wait(v_full + vs, (jj >> 1) & 1);
          tc_fence_after();
          const uint64_t vd = mnmajor_desc(smem_u32(v_s + vs * kTileBytes));
#pragma unroll
          for (int kk = 0; kk < kTile / UMMA_K; ++kk)             // 8 keys per MMA: +8 TMEM columns of P, +1024 bytes of V
            umma_tf32_ts(tmem_o, tmem_s + uint32_t(kk * UMMA_K), vd + uint64_t(kk * 64), kIdescPV, (jj | kk) != 0);
          umma_commit(v_empty + vs);
        }
      }
      umma_commit(o_full);
    }
  } else {
    // softmax warps: thread = query row q0 + 32 * warp + lane = TMEM lane
    const int row = q0 + warp * 32 + lane;
    const bool live = row < T;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const uint4* mrow = mask_bits + (size_t)(live ? row : 0) * (mask_words / 4);
    float m = -CUDART_INF_F, l = 0.f, inv_l = 0.f;
    float* prow = p_out ? p_out + ((size_t)(n * H + h) * T + (live ? row : 0)) * T : nullptr;
    for (int it = 0; it < 2 * nk; ++it) {
      const int j = it % nk;
      const bool second = it >= nk;
      const uint4 mw = __ldg(mrow + j);
      const uint32_t words[4] = {mw.x, mw.y, mw.z, mw.w};
      mbar_wait(s_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + uint32_t(c * 32), v);
        const uint32_t blocked = words[c];
        if (!second) {
          float cm = -CUDART_INF_F;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = ((blocked >> i) & 1u) ? -CUDART_INF_F : __uint_as_float(v[i]) * scale_log2;
            v[i] = __float_as_uint(x);
            cm = fmaxf(cm, x);
          }
          const float m_new = fmaxf(m, cm);
          if (m_new > -CUDART_INF_F) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sum += exp2f(__uint_as_float(v[i]) - m_new);
            l = l * exp2f(m - m_new) + sum;
            m = m_new;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = __uint_as_float(v[i]) * scale_log2 - m;
            const float p = ((blocked >> i) & 1u) ? 0.f : exp2f(x) * inv_l;
            v[i] = __float_as_uint(p);
          }
          tmem_st32(tmem_s + lane_addr + uint32_t(c * 32), v);
          if (prow != nullptr && live) {
            const int key0 = j * kTile + c * 32;
            if ((T & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (key0 + i < T) *reinterpret_cast<uint4*>(prow + key0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (key0 + i < T) prow[key0 + i] = __uint_as_float(v[i]);
            }
          }
        }
      }
      if (it == nk - 1) inv_l = l > 0.f ? 1.f / l : 0.f;         // a fully blocked row yields zeros (never the case in DINO)
      if (second) tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_done);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld32(tmem_o + lane_addr, o);
    if (live) {
      float* orow = out + ((size_t)n * T + row) * (size_t)(H * kD) + h * kD;
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(orow + i) = make_uint4(o[i], o[i + 1], o[i + 2], o[i + 3]);
      if (lse != nullptr) lse[((size_t)n * H + h) * T + row] = (m + log2f(l)) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}


// ---------------------------------------------------------------------------------------------------------------
// Backward.  With LSE_i = log sum_j exp(scale s_ij) from the forward and delta_i = <dO_i, O_i>:
//     P = exp(scale S - LSE),  dP = dO V^T,  dS = P o (dP - delta),  dQ = scale dS K,  dK = scale dS^T Q,  dV = P^T dO.
// Two launches of ONE kernel template, each recomputing the score tiles it needs in tensor memory (the products are
// cheap; nothing of size T x T ever reaches HBM):
//   kKV = false (CTA = 128 queries, loop over key tiles; rows / TMEM lanes = queries): S = Q K^T and dP = dO V^T, dS written
//     back over S, dQ += dS K with dS as the TMEM A operand and the K tile as MN-major B operand; also writes delta;
//   kKV = true (CTA = 128 keys, loop over query tiles; lanes = keys): S^T = K Q^T and dP^T = V dO^T, P^T over S^T and
//     dS^T over dP^T, dV += P^T dO and dK += dS^T Q with both A operands from tensor memory and dO / Q as MN-major B
//     operands.  Row statistics become column statistics here: LSE and delta of the tile's 128 queries are staged in
//     shared memory, the mask comes packed the other way round (bits along the queries).
// Warps 0-7: elementwise (warp w owns TMEM lanes 32 (w % 4) .. +31 and columns 64 (w / 4) .. +63 of a tile), warp 8: TMA,
// warp 9: MMA issue.  One CTA per SM (two 128-column score tiles + the accumulators need 320 TMEM columns).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kThreadsBwd = 320;
constexpr uint32_t kTmemColsBwd = 512;
struct BwdMaps { CUtensorMap r1, r2, c1, c2, mn1, mn2; };

template <bool kKV>
struct BwdSmem {
  static constexpr int kStage = (kKV ? 4 : 3) * kTileBytes;
  static constexpr int kStats = 2 * kTile * 8;
  static constexpr int kTotal = 2 * kTileBytes + 2 * kStage + kStats + 1024 + 1024;
};

template <bool kKV>
__global__ void __launch_bounds__(kThreadsBwd, 1)
attn_bwd_tf32(const __grid_constant__ BwdMaps maps, const uint4* __restrict__ mask_bits, int mask_words, int H, int T,
              float scale, const float* __restrict__ lse, float* __restrict__ delta, const float* __restrict__ dout,
              const float* __restrict__ outp, float* __restrict__ out1, long long out1_rs, float* __restrict__ out2,
              long long out2_rs) {
  using L = BwdSmem<kKV>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* r1_s = smem;
  unsigned char* r2_s = smem + kTileBytes;
  unsigned char* st_s = smem + 2 * kTileBytes;                               // 2 stages of {c1, c2, mn1[, mn2]}
  float2* stats = reinterpret_cast<float2*>(smem + 2 * kTileBytes + 2 * L::kStage);   // [2][128] {LSE * log2e, delta}
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kTileBytes + 2 * L::kStage + L::kStats);
  uint64_t* r_full = bars;
  uint64_t* st_full = bars + 1;    // [2]
  uint64_t* st_empty = bars + 3;   // [2]
  uint64_t* sd_full = bars + 5;    // MMA -> elementwise: score tiles complete
  uint64_t* e_done = bars + 6;     // elementwise -> MMA: A operands in tensor memory
  uint64_t* acc_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * kTile, h = blockIdx.y, n = blockIdx.z;
  const int nt = (T + kTile - 1) / kTile;
  constexpr float kLog2e = 1.4426950408889634f;

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.r1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.c1) : "memory");
    mbar_init(r_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(st_full + s, 1); mbar_init(st_empty + s, 1); }
    mbar_init(sd_full, 1);
    mbar_init(e_done, 8);
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(tmem_slot, kTmemColsBwd);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + kTile, tmem_a1 = tmem_base + 2 * kTile, tmem_a2 = tmem_a1 + kD;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(r_full, 2 * kTileBytes);
      tma_load_3d(r1_s, &maps.r1, h * kD, r0, n, r_full);
      tma_load_3d(r2_s, &maps.r2, h * kD, r0, n, r_full);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        unsigned char* st = st_s + s * L::kStage;
        mbar_wait(st_empty + s, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(st_full + s, L::kStage);
        tma_load_3d(st, &maps.c1, h * kD, j * kTile, n, st_full + s);
        tma_load_3d(st + kTileBytes, &maps.c2, h * kD, j * kTile, n, st_full + s);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tma_load_3d(st + 2 * kTileBytes + c * 4096, &maps.mn1, h * kD, j * kTile + 32 * c, n, st_full + s);
          if (kKV) tma_load_3d(st + 3 * kTileBytes + c * 4096, &maps.mn2, h * kD, j * kTile + 32 * c, n, st_full + s);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      mbar_wait(r_full, 0);
      const uint64_t r1d = kmajor_sw128_desc(smem_u32(r1_s)), r2d = kmajor_sw128_desc(smem_u32(r2_s));
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        const uint32_t st = smem_u32(st_s + s * L::kStage);
        mbar_wait(st_full + s, (j >> 1) & 1);
        tc_fence_after();
        const uint64_t c1d = kmajor_sw128_desc(st), c2d = kmajor_sw128_desc(st + kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kD / UMMA_K; ++kk) umma_tf32(tmem_s, r1d + uint64_t(kk * 2), c1d + uint64_t(kk * 2), kIdescS, kk != 0);
#pragma unroll
        for (int kk = 0; kk < kD / UMMA_K; ++kk) umma_tf32(tmem_dp, r2d + uint64_t(kk * 2), c2d + uint64_t(kk * 2), kIdescS, kk != 0);
        umma_commit(sd_full);
        mbar_wait(e_done, j & 1);
        tc_fence_after();
        const uint64_t m1d = mnmajor_desc(st + 2 * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kTile / UMMA_K; ++kk)
          umma_tf32_ts(tmem_a1, tmem_s + uint32_t(kk * UMMA_K), m1d + uint64_t(kk * 64), kIdescPV, (j | kk) != 0);
        if (kKV) {
          const uint64_t m2d = mnmajor_desc(st + 3 * kTileBytes);
#pragma unroll
          for (int kk = 0; kk < kTile / UMMA_K; ++kk)
            umma_tf32_ts(tmem_a2, tmem_dp + uint32_t(kk * UMMA_K), m2d + uint64_t(kk * 64), kIdescPV, (j | kk) != 0);
        }
        umma_commit(st_empty + s);
      }
      umma_commit(acc_full);
    }
  } else {
    // elementwise warps: TMEM lane = row of the CTA's tile, columns [col0, col0 + 64) of every score tile
    const int quarter = warp & 3, col0 = (warp >> 2) * 64;
    const int row = r0 + quarter * 32 + lane;
    const bool live = row < T;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint4* mrow = mask_bits + (size_t)(live ? row : 0) * (mask_words / 4);
    const int e = threadIdx.x;                                   // 0 .. 255
    const float sl2 = scale * kLog2e;
    float my_lse2 = 0.f, my_delta = 0.f;
    if (!kKV) {
      // rows are queries: LSE of the row, delta = <dO_row, O_row> over this head's channels (written out for the kKV pass)
      if (live) {
        const float4* a = reinterpret_cast<const float4*>(dout + ((size_t)n * T + row) * (size_t)(H * kD) + h * kD);
        const float4* b = reinterpret_cast<const float4*>(outp + ((size_t)n * T + row) * (size_t)(H * kD) + h * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 x = __ldg(a + i), y = __ldg(b + i);
          my_delta += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        }
        my_lse2 = __ldg(lse + ((size_t)n * H + h) * T + row) * kLog2e;
        if (col0 == 0) delta[((size_t)n * H + h) * T + row] = my_delta;
      }
    }
    for (int j = 0; j < nt; ++j) {
      if (kKV) {
        // columns are queries: stage {LSE * log2e, delta} of the tile's queries
        if (e < kTile) {
          const int q = j * kTile + e;
          float2 sd = make_float2(0.f, 0.f);
          if (q < T) sd = make_float2(__ldg(lse + ((size_t)n * H + h) * T + q) * kLog2e, __ldg(delta + ((size_t)n * H + h) * T + q));
          stats[(j & 1) * kTile + e] = sd;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const uint4 mw = __ldg(mrow + j);
      const uint32_t words[4] = {mw.x, mw.y, mw.z, mw.w};
      mbar_wait(sd_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cc = col0 + c * 32;
        uint32_t sv[32], dv[32];
        tmem_ld32(tmem_s + lane_addr + uint32_t(cc), sv);
        tmem_ld32(tmem_dp + lane_addr + uint32_t(cc), dv);
        const uint32_t blocked = words[cc >> 5];
        const float2* sc = stats + (j & 1) * kTile + cc;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float lse2 = my_lse2, dl = my_delta;
          if (kKV) { const float2 t = sc[i]; lse2 = t.x; dl = t.y; }
          const float p = ((blocked >> i) & 1u) ? 0.f : exp2f(__uint_as_float(sv[i]) * sl2 - lse2);
          const float ds = p * (__uint_as_float(dv[i]) - dl);
          if (kKV) { sv[i] = __float_as_uint(p); dv[i] = __float_as_uint(ds); }
          else sv[i] = __float_as_uint(ds);
        }
        tmem_st32(tmem_s + lane_addr + uint32_t(cc), sv);
        if (kKV) tmem_st32(tmem_dp + lane_addr + uint32_t(cc), dv);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(e_done);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // epilogue: warps 0-3 write accumulator 1 (dQ * scale, or dV), warps 4-7 accumulator 2 (dK * scale; kKV only)
    if (col0 == 0 || kKV) {
      uint32_t o[32];
      tmem_ld32((col0 == 0 ? tmem_a1 : tmem_a2) + lane_addr, o);
      if (live) {
        float* dst = col0 == 0 ? out1 + ((size_t)n * T + row) * (size_t)out1_rs + h * kD
                               : out2 + ((size_t)n * T + row) * (size_t)out2_rs + h * kD;
        const float f = (col0 == 0 && kKV) ? 1.f : scale;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(o[i]) * f, __uint_as_float(o[i + 1]) * f,
                                                            __uint_as_float(o[i + 2]) * f, __uint_as_float(o[i + 3]) * f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, kTmemColsBwd);
}

// blocked [T, T] bytes (or null) -> bits [T, words] (words = 4 * ceil(T / 128)), bit i of word w = key 32 w + i may NOT be
// attended; keys >= T are always blocked
// `bits_t` (optional, same shape): the transposed mask, bit i of word w of row r = query 32 w + i may not attend key r
__global__ void attn_pack_mask(const uint8_t* __restrict__ blocked, int T, int words, uint32_t* __restrict__ bits,
                               uint32_t* __restrict__ bits_t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * words) return;
  const int r = idx / words, w = idx % words;
  uint32_t b = 0, bt = 0;
  for (int i = 0; i < 32; ++i) {
    const int c = w * 32 + i;
    if (c >= T || (blocked != nullptr && blocked[(size_t)r * T + c])) b |= 1u << i;
    if (bits_t != nullptr && (c >= T || (blocked != nullptr && blocked[(size_t)c * T + r]))) bt |= 1u << i;
  }
  bits[idx] = b;
  if (bits_t != nullptr) bits_t[idx] = bt;
}

int encode3(CUtensorMap* map, const float* base, long long row_stride, int channels, int T, int N, int box_rows, bool mn_major) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return ffail(DATR_ATTN_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  const cuuint64_t gdim[3] = {cuuint64_t(channels), cuuint64_t(T), cuuint64_t(N)};
  const cuuint64_t gstr[2] = {cuuint64_t(row_stride) * 4, cuuint64_t(T) * cuuint64_t(row_stride) * 4};
  const cuuint32_t box[3] = {32, cuuint32_t(box_rows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_af_err, sizeof g_af_err, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return DATR_ATTN_ERR_CUDA;
  }
  return DATR_ATTN_OK;
}

}  // namespace

extern "C" {

int datr_attn_mask_words(int T) { return 4 * ((T + kTile - 1) / kTile); }

int datr_attn_pack_mask(const uint8_t* blocked, int T, uint32_t* bits, uint32_t* bits_t, void* stream_) {
  if (!bits || T <= 0) return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "bad argument%s");
  const int words = datr_attn_mask_words(T);
  const int total = T * words;
  attn_pack_mask<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(blocked, T, words, bits, bits_t);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ffail(DATR_ATTN_ERR_CUDA, "attn_pack_mask launch: %s", cudaGetErrorString(e));
  g_af_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_ATTN_OK;
}

int datr_attn_fused_forward(const float* q, long long q_row_stride, const float* k, long long k_row_stride, const float* v,
                            long long v_row_stride, const uint32_t* mask_bits, int N, int H, int T, float scale, float* out,
                            float* lse, float* p_out, void* stream_) {
  if (!q || !k || !v || !mask_bits || !out) return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (N <= 0 || H <= 0 || T <= 0 || T > 65535) return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "bad dimensions%s");
  const long long need = (long long)H * kD;
  if (q_row_stride < need || k_row_stride < need || v_row_stride < need || (q_row_stride & 3) || (k_row_stride & 3) ||
      (v_row_stride & 3))
    return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "row strides must cover H * 32 channels and be multiples of 4 elements%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(q) || !al16(k) || !al16(v) || !al16(out) || !al16(mask_bits) || (p_out && !al16(p_out)))
    return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "buffers must be 16-byte aligned%s");
  CUtensorMap mq, mk, mv;
  if (int rc = encode3(&mq, q, q_row_stride, H * kD, T, N, kTile, false)) return rc;
  if (int rc = encode3(&mk, k, k_row_stride, H * kD, T, N, kTile, false)) return rc;
  if (int rc = encode3(&mv, v, v_row_stride, H * kD, T, N, 32, true)) return rc;
  static std::atomic<uint64_t> opted{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(opted.load(std::memory_order_acquire) & bit)) {
    const cudaError_t e = cudaFuncSetAttribute(attn_fwd_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd);
    if (e != cudaSuccess) return ffail(DATR_ATTN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    opted.fetch_or(bit, std::memory_order_release);
  }
  const dim3 grid((T + kTile - 1) / kTile, H, N);
  attn_fwd_tf32<<<grid, kThreadsFwd, kSmemFwd, static_cast<cudaStream_t>(stream_)>>>(
      mq, mk, mv, reinterpret_cast<const uint4*>(mask_bits), datr_attn_mask_words(T), H, T, scale * 1.4426950408889634f, out,
      lse, p_out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ffail(DATR_ATTN_ERR_CUDA, "attn_fwd_tf32 launch: %s", cudaGetErrorString(e));
  g_af_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_ATTN_OK;
}

int datr_attn_fused_backward(const float* q, long long q_row_stride, const float* k, long long k_row_stride, const float* v,
                             long long v_row_stride, const uint32_t* mask_bits, const uint32_t* mask_bits_t, int N, int H,
                             int T, float scale, const float* out, const float* lse, const float* dout, float* delta,
                             float* dq, long long dq_row_stride, float* dk, long long dk_row_stride, float* dv,
                             long long dv_row_stride, void* stream_) {
  if (!q || !k || !v || !mask_bits || !mask_bits_t || !out || !lse || !dout || !delta || !dq || !dk || !dv)
    return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (N <= 0 || H <= 0 || T <= 0 || T > 65535) return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "bad dimensions%s");
  const long long need = (long long)H * kD;
  const long long strides[6] = {q_row_stride, k_row_stride, v_row_stride, dq_row_stride, dk_row_stride, dv_row_stride};
  for (long long st : strides)
    if (st < need || (st & 3)) return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "row strides must cover H * 32 channels and be multiples of 4 elements%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(q) || !al16(k) || !al16(v) || !al16(out) || !al16(dout) || !al16(dq) || !al16(dk) || !al16(dv) || !al16(mask_bits) ||
      !al16(mask_bits_t))
    return ffail(DATR_ATTN_ERR_BAD_ARGUMENT, "buffers must be 16-byte aligned%s");
  const int C = H * kD;
  CUtensorMap q_k, q_mn, k_k, k_mn, v_k, do_k, do_mn;
  if (int rc = encode3(&q_k, q, q_row_stride, C, T, N, kTile, false)) return rc;
  if (int rc = encode3(&q_mn, q, q_row_stride, C, T, N, 32, true)) return rc;
  if (int rc = encode3(&k_k, k, k_row_stride, C, T, N, kTile, false)) return rc;
  if (int rc = encode3(&k_mn, k, k_row_stride, C, T, N, 32, true)) return rc;
  if (int rc = encode3(&v_k, v, v_row_stride, C, T, N, kTile, false)) return rc;
  if (int rc = encode3(&do_k, dout, C, C, T, N, kTile, false)) return rc;
  if (int rc = encode3(&do_mn, dout, C, C, T, N, 32, true)) return rc;
  static std::atomic<uint64_t> opted{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(opted.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tf32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<false>::kTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_tf32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<true>::kTotal);
    if (e != cudaSuccess) return ffail(DATR_ATTN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    opted.fetch_or(bit, std::memory_order_release);
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const dim3 grid((T + kTile - 1) / kTile, H, N);
  const int words = datr_attn_mask_words(T);
  // pass 1 (rows = queries): dQ and delta;  rows Q, dO; columns K, V; MN-major K
  BwdMaps m1 = {q_k, do_k, k_k, v_k, k_mn, k_mn};
  attn_bwd_tf32<false><<<grid, kThreadsBwd, BwdSmem<false>::kTotal, stream>>>(
      m1, reinterpret_cast<const uint4*>(mask_bits), words, H, T, scale, lse, delta, dout, out, dq, dq_row_stride, nullptr, 0);
  // pass 2 (rows = keys): dV and dK;  rows K, V; columns Q, dO; MN-major dO and Q
  BwdMaps m2 = {k_k, v_k, q_k, do_k, do_mn, q_mn};
  attn_bwd_tf32<true><<<grid, kThreadsBwd, BwdSmem<true>::kTotal, stream>>>(
      m2, reinterpret_cast<const uint4*>(mask_bits_t), words, H, T, scale, lse, delta, dout, out, dv, dv_row_stride, dk, dk_row_stride);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ffail(DATR_ATTN_ERR_CUDA, "attn_bwd_tf32 launch: %s", cudaGetErrorString(e));
  g_af_launches.fetch_add(2, std::memory_order_relaxed);
  return DATR_ATTN_OK;
}

const char* datr_attn_fused_last_error(void) { return g_af_err; }
uint64_t datr_attn_fused_launch_count(void) { return g_af_launches.load(std::memory_order_relaxed); }

}  // extern "C"
