// wgrad_tf32.cu -- weight and bias gradients of a Linear layer on the 5th-generation tensor cores (sm_100a):
//
//     dW[N, K] = dz[M, N]^T . x[M, K]          db[N] = sum_m dz[m, n]
//
// for the nn.Linear layers of the DINO transformer and the 1x1 convolutions of the ResNet bottlenecks (reference
// autograd of torch.nn.functional.linear at models/dino/ops/modules/ms_deform_attn.py:94-125 and
// models/dino/deformable_transformer.py:784-805, :941-947), M = batch * tokens = 44 446 rows (or pixels).
//
//   * the contraction runs over the ROWS of both operands, so both are "MN-major" for the MMA: a TMA box of
//     {32 columns, 32 rows} lands as 32 rows of 128 bytes (128-byte swizzle with 32-byte atoms), and
//     tcgen05.mma.kind::tf32 reads it with the transposed-operand bits set in the instruction descriptor (leading
//     byte offset = distance between 32-column chunks, each MMA consumes 8 rows) -- neither operand is ever
//     transposed in memory;
//   * split-K: the M rows are cut into slabs; a CTA owns one 128 x BN tile of dW and one slab, accumulates in TMEM and
//     adds its partial tile to dW with vector reductions (dW / db are zero-filled by the library on the stream);
//   * db rides along as one extra MMA per 8 rows against a constant tile of ones (16 more TMEM columns);
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue; 4-stage mbarrier ring.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "datr_linear.h"
#include "tcgen05_common.cuh"

namespace {

using namespace datr_tc;

thread_local char g_wg_err[512] = "";
std::atomic<uint64_t> g_wg_launches{0};

int wfail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_wg_err, sizeof g_wg_err, fmt, detail);
  return code;
}

constexpr int kThreads = 192;
constexpr int kChunk = 32 * BK * 4;         // bytes of one {32 columns x 32 rows} box
constexpr int kOnesCols = 16;               // N of the bias-gradient MMA

constexpr int kChunkBf = 64 * 64 * 2;       // bf16 operands: bytes of one {64 columns x 64 rows} box

template <int BN, int STAGES, bool kBF16 = false>
struct Smem {
  static constexpr int kCh = kBF16 ? kChunkBf : kChunk;
  static constexpr int kA = kBF16 ? 2 * kChunkBf : 4 * kChunk, kB = kBF16 ? (BN / 64) * kChunkBf : (BN / 32) * kChunk;
  static constexpr int kStage = kA + kB;
  static constexpr int kOnes = kCh;
  static constexpr int kBars = 1024;
  static constexpr int kTotal = STAGES * kStage + kOnes + kBars + 1024;
};

// MN-major 16-bit operand: TMA boxes of {64 columns x 64 rows} with the plain 128-byte swizzle; rows (= the contraction
// index) are 128 bytes, the swizzle atom is 8 rows (1024 bytes = stride byte offset), 64-column chunks are kChunkBf bytes
// apart (leading byte offset); each K = 16 MMA consumes 16 rows = 2048 bytes.
__device__ __forceinline__ uint64_t mnmajor_bf16_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(kChunkBf >> 4) << 16) | (uint64_t(1024 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// D = fp32, A = B = bf16, both MN-major
__host__ __device__ constexpr uint32_t bf16_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}

// MN-major 32-bit operand: the tensor core transposes 32-bit elements in 32-byte units, so the tile uses the
// "128B swizzle with 32B atoms" (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA layout type SWIZZLE_128B_BASE32B = 1;
// the plain 128B swizzle silently yields zeros for transposed TF32 operands).  Rows are 128 bytes, the swizzle atom is
// 4 rows (512 bytes = stride byte offset between K atoms), 32-column chunks are `kChunk` bytes apart (leading byte
// offset); descriptor version 1.
__device__ __forceinline__ uint64_t mnmajor_sw128_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(kChunk >> 4) << 16) | (uint64_t(512 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(1) << 61);
}

// D = fp32, A = B = TF32, both MN-major (bits 15, 16), N = n, M = 128
__host__ __device__ constexpr uint32_t tf32_idesc_mn(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}

__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// 3x3 convolution mode (on != 0): the reduction runs over output pixels, taken as patches of kPW x kPH = 32 pixels of one
// image; `dz` is the output gradient [N, Ho, Wo, Cout], `x` the input [N, H, W, Cin], and dW is the channels_last weight
// [Cout, 3, 3, Cin] seen as a [Cout, 9 * Cin] matrix: a BN-wide column tile lies inside ONE filter tap (Cin % BN == 0),
// whose shifted input patch is fetched by a 4-D TMA box (zero fill = the convolution's padding, element stride = its
// stride), exactly like the forward kernel's A operand (conv3x3_tf32.cu).
struct ConvGeom { int on, Cin, stride, tiles_x, tiles_y; };
constexpr int kPW = 16, kPH = 2;

template <int BN, int STAGES, bool kBF16 = false>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tma_dz, const __grid_constant__ CUtensorMap tma_x,
                  float* __restrict__ dw, float* __restrict__ db, int M, int N, int K, int splits, int rows_per_split,
                  const ConvGeom cg) {
  using L = Smem<BN, STAGES, kBF16>;
  constexpr int BKr = kBF16 ? 64 : BK;       // rows (contraction steps) per pipeline stage
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* ones = reinterpret_cast<float*>(smem + STAGES * L::kStage);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage + L::kOnes);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  constexpr uint32_t kTmemCols = BN == 256 ? 512 : 256;      // BN accumulator columns + 16 for db, power of two

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_tiles = (K + BN - 1) / BN;
  const int tile = blockIdx.x / splits, split = blockIdx.x % splits;
  const int n0 = (tile / k_tiles) * BM, k0 = (tile % k_tiles) * BN;
  const int m_begin = split * rows_per_split;
  const int m_end = min(M, m_begin + rows_per_split);
  const int kblocks = m_end > m_begin ? (m_end - m_begin + BKr - 1) / BKr : 0;

  if constexpr (kBF16) {
    for (int i = threadIdx.x; i < kChunkBf / 4; i += kThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;   // bf16 1.0 pairs
  } else {
    for (int i = threadIdx.x; i < kChunk / 4; i += kThreads) ones[i] = 1.0f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_dz) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_x) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (kblocks > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int kb = 0; kb < kblocks; ++kb) {
          const int s = kb % STAGES;
          const int m = m_begin + kb * BKr;
          mbar_wait(empty + s, ((kb / STAGES) & 1) ^ 1);
          mbar_expect_tx(full + s, L::kStage);
          unsigned char* a = smem + s * L::kStage;
          if (cg.on) {
            const int pb = m / BK;                              // patch index: (image, patch row, patch column)
            const int per_img = cg.tiles_x * cg.tiles_y;
            const int img = pb / per_img, q = pb - img * per_img;
            const int y0 = (q / cg.tiles_x) * kPH, x0 = (q % cg.tiles_x) * kPW;
            const int tap = k0 / cg.Cin, ci0 = k0 - tap * cg.Cin;
            const int ix = cg.stride * x0 + tap % 3 - 1, iy = cg.stride * y0 + tap / 3 - 1;
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_load_4d(a + c * kChunk, &tma_dz, n0 + 32 * c, x0, y0, img, full + s);
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) tma_load_4d(a + L::kA + c * kChunk, &tma_x, ci0 + 32 * c, ix, iy, img, full + s);
            continue;
          }
          if constexpr (kBF16) {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d(a + c * kChunkBf, &tma_dz, n0 + 64 * c, m, full + s);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(a + L::kA + c * kChunkBf, &tma_x, k0 + 64 * c, m, full + s);
            continue;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_2d(a + c * kChunk, &tma_dz, n0 + 32 * c, m, full + s);
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_2d(a + L::kA + c * kChunk, &tma_x, k0 + 32 * c, m, full + s);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = kBF16 ? bf16_idesc_mn(BN) : tf32_idesc_mn(BN);
        constexpr uint32_t idesc1 = kBF16 ? bf16_idesc_mn(kOnesCols) : tf32_idesc_mn(kOnesCols);
        const uint64_t od = kBF16 ? mnmajor_bf16_desc(smem_u32(ones)) : mnmajor_sw128_desc(smem_u32(ones));
        for (int kb = 0; kb < kblocks; ++kb) {
          const int s = kb % STAGES;
          mbar_wait(full + s, (kb / STAGES) & 1);
          tc_fence_after();
          const uint32_t a = smem_u32(smem + s * L::kStage);
          if constexpr (kBF16) {
            const uint64_t ad = mnmajor_bf16_desc(a), bd = mnmajor_bf16_desc(a + L::kA);
#pragma unroll
            for (int j = 0; j < 4; ++j) {                     // next 16 rows = +2048 bytes = +128 in the address field
              umma_bf16(tmem_d, ad + uint64_t(j * 128), bd + uint64_t(j * 128), idesc, (kb | j) != 0);
              umma_bf16(tmem_d + BN, ad + uint64_t(j * 128), od, idesc1, (kb | j) != 0);
            }
          } else {
            const uint64_t ad = mnmajor_sw128_desc(a), bd = mnmajor_sw128_desc(a + L::kA);
#pragma unroll
            for (int j = 0; j < BK / UMMA_K; ++j) {           // next 8 rows = +1024 bytes = +64 in the address field
              umma_tf32(tmem_d, ad + uint64_t(j * 64), bd + uint64_t(j * 64), idesc, (kb | j) != 0);
              umma_tf32(tmem_d + BN, ad + uint64_t(j * 64), od, idesc1, (kb | j) != 0);
            }
          }
          umma_commit(empty + s);
        }
        umma_commit(acc_full);
      }
    } else {
      // accumulator rows = rows of dW.  tcgen05.ld.16x256b hands four neighbouring lanes one 32-byte sector of a dW row
      // (profiles/r02f_tmem_ld_layout_probe.txt), so the partial tile is added with 8-byte vector reductions that
      // always cover whole sectors (the 32x32b layout scattered 16-byte pieces over 32 rows per instruction).
      const int lane_base = (warp & 3) * 32;
      const int n = n0 + lane_base + lane;                    // row of dW whose bias gradient this thread adds
      const int rq = lane >> 2, cq = (lane & 3) * 2;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t t0 = tmem_d + (uint32_t(lane_base) << 16);
#pragma unroll 1
      for (int ch = 0; ch < 2 * (BN / 64); ++ch) {
        const int h = ch & 1, c0 = (ch >> 1) * 64;
        uint32_t v[32];
        tmem_ld_16x256b_x8(tmem_d + uint32_t(c0) + (uint32_t(lane_base + 16 * h) << 16), v);
        const int n_lo = n0 + lane_base + 16 * h + rq, n_hi = n_lo + 8;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int col = k0 + c0 + 8 * g + cq;
          if (col >= K) break;                                // K % 4 == 0 and col is even
          if (n_lo < N) red_add2(dw + (size_t)n_lo * K + col, __uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]));
          if (n_hi < N) red_add2(dw + (size_t)n_hi * K + col, __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
        }
      }
      if (db != nullptr && k0 == 0) {                         // one K tile per dW row block carries the bias gradient
        uint32_t v[32];
        tmem_ld32(t0 + uint32_t(BN), v);
        if (n < N) atomicAdd(db + n, __uint_as_float(v[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, kTmemCols);
}

int make_map(CUtensorMap* map, const float* base, int rows, int cols) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return wfail(DATR_LINEAR_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(cols) * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_wg_err, sizeof g_wg_err, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return DATR_LINEAR_ERR_CUDA;
  }
  return DATR_LINEAR_OK;
}

int make_map_bf16(CUtensorMap* map, const void* base, int rows, int cols) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return wfail(DATR_LINEAR_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(cols) * 2};
  const cuuint32_t box[2] = {64, 64};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_wg_err, sizeof g_wg_err, "cuTensorMapEncodeTiled (bf16) failed (CUresult %d)", int(r));
    return DATR_LINEAR_ERR_CUDA;
  }
  return DATR_LINEAR_OK;
}

template <int BN, int STAGES, bool kBF16 = false>
int launch(const CUtensorMap& mdz, const CUtensorMap& mx, float* dw, float* db, int M, int N, int K, cudaStream_t stream,
           const ConvGeom& cg = ConvGeom{0, 0, 0, 0, 0}) {
  using L = Smem<BN, STAGES, kBF16>;
  constexpr int BKr = kBF16 ? 64 : BK;
  static std::atomic<uint64_t> opted{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(opted.load(std::memory_order_acquire) & bit)) {
    const cudaError_t e = cudaFuncSetAttribute(wgrad_tf32_kernel<BN, STAGES, kBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return wfail(DATR_LINEAR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    opted.fetch_or(bit, std::memory_order_release);
  }
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const int tiles = ((N + BM - 1) / BM) * ((K + BN - 1) / BN);
  // one CTA per SM (shared memory): every wave pays the prologue and the reduction epilogue again.  Measured
  // (profiles/r02f_wgrad_waves.txt, M = 44 446): one wave is faster when dW has few row blocks (N <= 256: 55 -> 41 us at
  // 256 x 256, 173 -> 162 us at 256 x 2048), two waves when it has many (N = 2048: 136 vs 157 us).
  const int waves = N <= 256 ? 1 : 2;
  int splits = (waves * sms + tiles - 1) / tiles;
  // at least 128 rows per slab; 512 for the wide decoder-size gradients (M = 4 400, N x K = 2048 x 256: 37.6 -> 30.6 us,
  // 256 x 2048: 39.3 -> 28.7 us, tools/bench_wgrad_small.py); DATR_WGRAD_MIN_ROWS overrides (tuning hook)
  static const int forced_rows = getenv("DATR_WGRAD_MIN_ROWS") ? atoi(getenv("DATR_WGRAD_MIN_ROWS")) : 0;
  const int min_rows = forced_rows > 0 ? forced_rows : ((long long)N * K >= 2048LL * 256 ? 16 * BK : 4 * BK);
  const int max_splits = (M + min_rows - 1) / min_rows;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows_per_split = ((M + splits - 1) / splits + BKr - 1) / BKr * BKr;
  splits = (M + rows_per_split - 1) / rows_per_split;
  wgrad_tf32_kernel<BN, STAGES, kBF16><<<unsigned(tiles * splits), kThreads, L::kTotal, stream>>>(mdz, mx, dw, db, M, N, K, splits,
                                                                                                 rows_per_split, cg);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return wfail(DATR_LINEAR_ERR_CUDA, "wgrad_tf32_kernel launch: %s", cudaGetErrorString(e));
  g_wg_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_LINEAR_OK;
}

// dW / db are accumulated with reductions: zero-fill them on the stream -- with ONE memset when db directly follows dW in
// memory (the Python binding allocates them that way: ~150 launches fewer per training step)
int zero_outputs(float* dw, float* db, int N, int K, cudaStream_t stream) {
  const size_t nw = (size_t)N * K;
  cudaError_t e;
  if (db == dw + nw) {
    e = cudaMemsetAsync(dw, 0, sizeof(float) * (nw + (size_t)N), stream);
  } else {
    e = cudaMemsetAsync(dw, 0, sizeof(float) * nw, stream);
    if (e == cudaSuccess && db) e = cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, stream);
  }
  if (e != cudaSuccess) return wfail(DATR_LINEAR_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  return DATR_LINEAR_OK;
}

}  // namespace

extern "C" {

static int wgrad_tf32(const float* dz, const float* x, float* dw, float* db, int M, int N, int K, void* stream_, bool accumulate) {
  if (!dz || !x || !dw) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (M <= 0 || N <= 0 || K <= 0) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (N % 4 != 0 || K % 4 != 0) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "N and K must be multiples of 4%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(dz) || !al16(x) || !al16(dw)) return wfail(DATR_LINEAR_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!accumulate)
    if (int rc = zero_outputs(dw, db, N, K, stream)) return rc;
  CUtensorMap mdz, mx;
  if (int rc = make_map(&mdz, dz, M, N)) return rc;
  if (int rc = make_map(&mx, x, M, K)) return rc;
  return K > 128 ? launch<256, 4>(mdz, mx, dw, db, M, N, K, stream) : launch<128, 6>(mdz, mx, dw, db, M, N, K, stream);
}

int datr_linear_wgrad_tf32(const float* dz, const float* x, float* dw, float* db, int M, int N, int K, void* stream_) {
  return wgrad_tf32(dz, x, dw, db, M, N, K, stream_, false);
}

// dw += dz^T x, db += column sums of dz: the same kernel without the zero fill -- its partial tiles are reduced into dw / db
// with atomics anyway, so a gradient buffer that already holds earlier contributions (the step's flat .grad buffer) takes the
// weight gradient directly: no temporary, no zero fill, no `grad += new` pass.
int datr_linear_wgrad_tf32_acc(const float* dz, const float* x, float* dw, float* db, int M, int N, int K, void* stream_) {
  return wgrad_tf32(dz, x, dw, db, M, N, K, stream_, true);
}

// bf16 operands (dz [M, N], x [M, K] as bf16), fp32 accumulation, dw [N, K] / db [N] fp32.
static int wgrad_bf16(const void* dz, const void* x, float* dw, float* db, int M, int N, int K, void* stream_, bool accumulate);

int datr_linear_wgrad_bf16(const void* dz, const void* x, float* dw, float* db, int M, int N, int K, void* stream_) {
  return wgrad_bf16(dz, x, dw, db, M, N, K, stream_, false);
}

int datr_linear_wgrad_bf16_acc(const void* dz, const void* x, float* dw, float* db, int M, int N, int K, void* stream_) {
  return wgrad_bf16(dz, x, dw, db, M, N, K, stream_, true);
}

static int wgrad_bf16(const void* dz, const void* x, float* dw, float* db, int M, int N, int K, void* stream_, bool accumulate) {
  if (!dz || !x || !dw) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (M <= 0 || N <= 0 || K <= 0) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (N % 8 != 0 || K % 8 != 0) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "N and K must be multiples of 8 for bf16 operands%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(dz) || !al16(x) || !al16(dw)) return wfail(DATR_LINEAR_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!accumulate)
    if (int rc = zero_outputs(dw, db, N, K, stream)) return rc;
  CUtensorMap mdz, mx;
  if (int rc = make_map_bf16(&mdz, dz, M, N)) return rc;
  if (int rc = make_map_bf16(&mx, x, M, K)) return rc;
  return K > 128 ? launch<256, 4, true>(mdz, mx, dw, db, M, N, K, stream) : launch<128, 6, true>(mdz, mx, dw, db, M, N, K, stream);
}

// Weight (and bias) gradient of a 3x3 convolution, padding 1, stride 1 or 2, NHWC tensors:
//   dw[co, ky, kx, ci] = sum_{n, oy, ox} gz[n, oy, ox, co] * x[n, s*oy + ky - 1, s*ox + kx - 1, ci],   db[co] = sum gz[..., co]
// (the wgrad half of aten.convolution_backward for the ResNet bottleneck conv2 layers, reference backbone.py:97, the
// image-level discriminator, DA_utils.py:50-79, and the extra input projection, dino.py:118-123).
int datr_conv3x3_wgrad_nhwc_tf32(const float* gz, const float* x, float* dw, float* db, int N, int H, int W, int Cin,
                                 int Cout, int stride, void* stream_) {
  if (!gz || !x || !dw) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (stride != 1 && stride != 2) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "stride must be 1 or 2%s");
  if (Cin % 128 != 0 || Cout % 4 != 0) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "Cin %% 128 == 0 and Cout %% 4 == 0 required%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(gz) || !al16(x) || !al16(dw)) return wfail(DATR_LINEAR_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int K = 9 * Cin;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * K, stream);
  if (e == cudaSuccess && db) e = cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, stream);
  if (e != cudaSuccess) return wfail(DATR_LINEAR_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  ConvGeom cg;
  cg.on = 1; cg.Cin = Cin; cg.stride = stride;
  cg.tiles_x = (Wo + kPW - 1) / kPW; cg.tiles_y = (Ho + kPH - 1) / kPH;
  const long long patches = (long long)N * cg.tiles_x * cg.tiles_y;
  if (patches * BK > 0x7fffffffLL) return wfail(DATR_LINEAR_ERR_BAD_ARGUMENT, "problem too large%s");
  EncodeTiledFn enc = encode_fn();
  if (!enc) return wfail(DATR_LINEAR_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  CUtensorMap mdz, mx;
  auto encode4 = [&](CUtensorMap* map, const float* base, int C, int w_, int h_, int st) -> int {
    const cuuint64_t gdim[4] = {cuuint64_t(C), cuuint64_t(w_), cuuint64_t(h_), cuuint64_t(N)};
    const cuuint64_t gstr[3] = {cuuint64_t(C) * 4, cuuint64_t(w_) * C * 4, cuuint64_t(h_) * w_ * C * 4};
    const cuuint32_t box[4] = {32, cuuint32_t((kPW - 1) * st + 1), cuuint32_t((kPH - 1) * st + 1), 1};
    const cuuint32_t estr[4] = {1, cuuint32_t(st), cuuint32_t(st), 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(g_wg_err, sizeof g_wg_err, "cuTensorMapEncodeTiled (4-D) failed (CUresult %d)", int(r));
      return DATR_LINEAR_ERR_CUDA;
    }
    return DATR_LINEAR_OK;
  };
  if (int rc = encode4(&mdz, gz, Cout, Wo, Ho, 1)) return rc;
  if (int rc = encode4(&mx, x, Cin, W, H, stride)) return rc;
  // the kernel walks "rows" in units of BK: one patch of 32 pixels per k-block
  const int M = int(patches) * BK;
  return Cin % 256 == 0 ? launch<256, 4>(mdz, mx, dw, db, M, Cout, K, stream, cg) : launch<128, 6>(mdz, mx, dw, db, M, Cout, K, stream, cg);
}

const char* datr_linear_wgrad_last_error(void) { return g_wg_err; }
uint64_t datr_linear_wgrad_launch_count(void) { return g_wg_launches.load(std::memory_order_relaxed); }

}  // extern "C"
