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

template <int BN, int STAGES>
struct Smem {
  static constexpr int kA = BM * BK * 4, kB = BN * BK * 4, kStage = kA + kB;
  static constexpr int kEpi = kEpiWarps * kStageTile;
  static constexpr int kBars = 1024;  // barriers + TMEM slot
  static constexpr int kTotal = STAGES * kStage + kEpi + kBars + 1024 /* alignment slack */;
};

// ---------------------------------------------------------------------------------------------------------------
// kernel: persistent, one CTA per SM, static round-robin over 128 x BN output tiles (tiles that share an X row
// block are adjacent in the order, so concurrently running CTAs hit the same X tile in L2).  The accumulator is
// double-buffered in tensor memory (2 x BN columns): the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
linear_tf32_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                   const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ y,
                   int M, int N, int K, int relu) {
  using L = Smem<BN, STAGES>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi = reinterpret_cast<float*>(smem + STAGES * L::kStage);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage + L::kEpi);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = K / BK;
  const int n_tiles = (N + BN - 1) / BN, m_tiles = (M + BM - 1) / BM;
  const int tiles = n_tiles * m_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_w) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, kEpiWarps); }
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
          tma_load_2d(a, &tma_x, kk * BK, m0, full + s);
          tma_load_2d(a + L::kA, &tma_w, kk * BK, n0, full + s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tf32_idesc<BN>();
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
          const uint64_t ad = kmajor_sw128_desc(a), bd = kmajor_sw128_desc(a + L::kA);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)   // +32 bytes along K inside the swizzle row = +2 in the address field
            umma_tf32(tmem_d, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0);
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
    float* tile_s = epi + (warp - 2) * (kStageTile / 4);
    const int tr = lane >> 3, tc = (lane & 7) * 4;           // transposed role: row tr + 4*j, columns tc .. tc+3
    constexpr int kCols = BN / 2;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN + half * kCols;
      const uint32_t as = ti & 1;
      const int row0 = m0 + lane_base + tr;
      float4 res[8];
      auto fetch_residual = [&](int c) {
        const int col = n0 + c + tc;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int row = row0 + 4 * j;
          res[j] = (residual && row < M && col + 4 <= N)
                       ? __ldg(reinterpret_cast<const float4*>(residual + (size_t)row * N + col))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
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
          if (bias) {
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
              *reinterpret_cast<float4*>(y + off) = t;
            } else if (vec) {
              t.x += res[j].x; t.y += res[j].y; t.z += res[j].z; t.w += res[j].w;
              if (relu == 2) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
              *reinterpret_cast<float4*>(y + off) = t;
            } else {
              const float ov[4] = {t.x, t.y, t.z, t.w};
              for (int e = 0; e < 4 && col + e < N; ++e) {
                const float r1 = residual ? __ldg(residual + off + e) : 0.f;
                const float o1 = relu == 3 ? (r1 > 0.f ? ov[e] : 0.f) : ov[e] + r1;
                y[off + e] = relu == 2 ? fmaxf(o1, 0.f) : o1;
              }
            }
          }
        }
        if (residual && c + 32 < kCols) fetch_residual(c + 32);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// [rows, cols] fp32 row-major matrix -> tensor map with a (box_rows x 32 columns) box, 128-byte swizzle, zero fill
int make_map(CUtensorMap* map, const float* base, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return lfail(DATR_LINEAR_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(cols) * 4};
  const cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_lin_err, sizeof g_lin_err, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return DATR_LINEAR_ERR_CUDA;
  }
  return DATR_LINEAR_OK;
}

template <int BN, int STAGES>
int launch(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, const float* residual, float* y, int M, int N,
           int K, int relu, cudaStream_t stream) {
  using L = Smem<BN, STAGES>;
  static std::atomic<uint64_t> opted{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(opted.load(std::memory_order_acquire) & bit)) {
    const cudaError_t e = cudaFuncSetAttribute(linear_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
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
  linear_tf32_kernel<BN, STAGES><<<grid, kThreads, L::kTotal, stream>>>(mx, mw, bias, residual, y, M, N, K, relu);
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

const char* datr_linear_last_error(void) { return g_lin_err; }
uint64_t datr_linear_launch_count(void) { return g_lin_launches.load(std::memory_order_relaxed); }

}  // extern "C"
