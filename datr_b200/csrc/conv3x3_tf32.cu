// conv3x3_tf32.cu -- 3x3 convolution (padding 1, stride 1 or 2) + per-channel bias (+ReLU) on NHWC fp32 activations
// as an im2col-free implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
// This is the conv2 of every ResNet-50 bottleneck (reference models/dino/backbone.py:97 -> torchvision Bottleneck;
// 16 layers, 64..512 channels) followed by FrozenBatchNorm2d (backbone.py:62-72, folded into weight and bias by the
// caller) and ReLU.  Formulation: out[n, y, x, :] = sum over the 9 taps (dy, dx) and channel blocks of
//       in[n, s*y + dy - 1, s*x + dx - 1, c0:c0+32] . W[:, dy, dx, c0:c0+32]^T
//   * the GEMM M dimension is a patch of 8 x 16 output pixels of one image; for each (tap, channel block) ONE 4-D TMA
//     box load {32 channels, 16 pixels (element stride s), 8 rows (element stride s), 1 image} at the shifted
//     coordinate brings the A tile straight from the NHWC tensor -- out-of-image pixels (the zero padding) are zero-
//     filled by TMA, nothing is ever unfolded in memory.  The box lands as 128 rows of 128 bytes in the 128-byte
//     swizzle, i.e. exactly the K-major operand tile the GEMM kernel uses;
//   * B tiles come from the weight viewed as [Cout, 9*Cin] (the NHWC / channels_last weight layout) with a 2-D map;
//   * 9*Cin/32 k-blocks accumulate into one TMEM accumulator (tcgen05.mma.kind::tf32, M = 128, N = BN);
//   * same persistent / warp-specialised structure as linear_tf32.cu: TMA warp, MMA warp, 8 epilogue warps,
//     double-buffered accumulator; the epilogue adds the bias, applies ReLU and stores whole 128-byte pixel rows.
// The backward pass (dgrad / wgrad) stays with cuDNN (see datr_b200/conv.py).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_conv.h"
#include "tcgen05_common.cuh"

namespace {

using namespace datr_tc;

thread_local char g_conv_err[512] = "";
std::atomic<uint64_t> g_conv_launches{0};

int cfail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_conv_err, sizeof g_conv_err, fmt, detail);
  return code;
}

constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kEpiWarps = 8;
constexpr int TH = 8, TW = 16;  // output patch of one tile: 8 rows x 16 pixels = 128 GEMM rows

template <int BN, int STAGES>
struct Smem {
  static constexpr int kA = BM * BK * 4, kB = BN * BK * 4, kStage = kA + kB;
  static constexpr int kEpi = kEpiWarps * kStageTile;
  static constexpr int kBars = 1024;
  static constexpr int kTotal = STAGES * kStage + kEpi + kBars + 1024;
};

struct ConvShape {
  int N, H, W, Cin, Cout, Ho, Wo, stride, tiles_x, tiles_y, n_tiles;  // n_tiles = Cout tiles
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tf32_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                    const float* __restrict__ bias, float* __restrict__ y, const ConvShape cs, int relu) {
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
  const int cblocks = cs.Cin / BK;
  const int kblocks = 9 * cblocks;
  const int spatial = cs.tiles_x * cs.tiles_y;
  const int tiles = cs.N * spatial * cs.n_tiles;
  // tile -> (image, patch row, patch column, Cout tile); Cout tiles of one patch are adjacent (shared A in L2)
  auto decode = [&](int tile, int& n, int& y0, int& x0, int& n0) {
    n0 = (tile % cs.n_tiles) * BN;
    const int p = tile / cs.n_tiles;
    n = p / spatial;
    const int q = p % spatial;
    y0 = (q / cs.tiles_x) * TH;
    x0 = (q % cs.tiles_x) * TW;
  };

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
        int n, y0, x0, n0;
        decode(tile, n, y0, x0, n0);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const int tap = kb / cblocks, cb = kb % cblocks;
          const int dy = tap / 3, dx = tap % 3;
          mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(full + s, L::kStage);
          unsigned char* a = smem + s * L::kStage;
          // input pixel of output (y0 + i, x0 + j): (s*(y0+i) + dy - 1, s*(x0+j) + dx - 1); TMA zero-fills the border
          tma_load_4d(a, &tma_x, cb * BK, cs.stride * x0 + dx - 1, cs.stride * y0 + dy - 1, n, full + s);
          tma_load_2d(a + L::kA, &tma_w, tap * cs.Cin + cb * BK, n0, full + s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tf32_idesc<BN>();
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
        const uint32_t as = ti & 1;
        mbar_wait(acc_empty + as, ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          mbar_wait(full + s, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a = smem_u32(smem + s * L::kStage);
          const uint64_t ad = kmajor_sw128_desc(a), bd = kmajor_sw128_desc(a + L::kA);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_tf32(tmem_d, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0);
          umma_commit(empty + s);
        }
        umma_commit(acc_full + as);
      }
    }
  } else {
    const int lane_base = (warp & 3) * 32;                   // GEMM rows (= patch pixels) of this warp
    const int half = (warp - 2) >> 2;
    float* tile_s = epi + (warp - 2) * (kStageTile / 4);
    const int tr = lane >> 3, tc = (lane & 7) * 4;
    constexpr int kCols = BN / 2;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      int n, y0, x0, n0;
      decode(tile, n, y0, x0, n0);
      n0 += half * kCols;
      const uint32_t as = ti & 1;
      mbar_wait(acc_full + as, (ti >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN + half * kCols + (uint32_t(lane_base) << 16);
#pragma unroll 1
      for (int c = 0; c < kCols; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_d + uint32_t(c), v);
        if (c + 32 >= kCols) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + as);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(tile_s + lane * kStagePitch + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        float4 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = *reinterpret_cast<const float4*>(tile_s + (tr + 4 * j) * kStagePitch + tc);
        __syncwarp();
        const int col = n0 + c + tc;
        if (col + 4 <= cs.Cout) {
          const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = lane_base + tr + 4 * j;            // row of the 128-pixel patch
            const int oy = y0 + r / TW, ox = x0 + r % TW;
            if (oy < cs.Ho && ox < cs.Wo) {
              float4 t = o[j];
              t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w;
              if (relu == 1) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
              else if (relu == 2) {       // LeakyReLU(0.2) of the image-level domain discriminator (DA_utils.py:61-79)
                t.x = t.x > 0.f ? t.x : 0.2f * t.x; t.y = t.y > 0.f ? t.y : 0.2f * t.y;
                t.z = t.z > 0.f ? t.z : 0.2f * t.z; t.w = t.w > 0.f ? t.w : 0.2f * t.w;
              }
              *reinterpret_cast<float4*>(y + (((size_t)n * cs.Ho + oy) * cs.Wo + ox) * cs.Cout + col) = t;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

int encode(CUtensorMap* map, const float* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstride,
           const cuuint32_t* box, const cuuint32_t* estr) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return cfail(DATR_CONV_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable%s");
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, rank, const_cast<float*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_conv_err, sizeof g_conv_err, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return DATR_CONV_ERR_CUDA;
  }
  return DATR_CONV_OK;
}

template <int BN, int STAGES>
int launch(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, float* y, const ConvShape& cs, int relu,
           cudaStream_t stream) {
  using L = Smem<BN, STAGES>;
  static std::atomic<uint64_t> opted{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(opted.load(std::memory_order_acquire) & bit)) {
    const cudaError_t e = cudaFuncSetAttribute(conv3x3_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return cfail(DATR_CONV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    opted.fetch_or(bit, std::memory_order_release);
  }
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const long long tiles = (long long)cs.N * cs.tiles_x * cs.tiles_y * cs.n_tiles;
  const unsigned grid = unsigned(tiles < sms ? tiles : sms);
  conv3x3_tf32_kernel<BN, STAGES><<<grid, kThreads, L::kTotal, stream>>>(mx, mw, bias, y, cs, relu);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cfail(DATR_CONV_ERR_CUDA, "conv3x3_tf32_kernel launch: %s", cudaGetErrorString(e));
  g_conv_launches.fetch_add(1, std::memory_order_relaxed);
  return DATR_CONV_OK;
}

}  // namespace

extern "C" {

int datr_conv3x3_nhwc_tf32(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int Cin,
                           int Cout, int stride, int relu, void* stream_) {
  if (!x || !w || !y) return cfail(DATR_CONV_ERR_BAD_ARGUMENT, "null pointer argument%s");
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return cfail(DATR_CONV_ERR_BAD_ARGUMENT, "all dimensions must be positive%s");
  if (stride != 1 && stride != 2) return cfail(DATR_CONV_ERR_BAD_ARGUMENT, "stride must be 1 or 2%s");
  if (Cin % BK != 0 || Cout % 4 != 0) return cfail(DATR_CONV_ERR_BAD_ARGUMENT, "Cin %% 32 == 0 and Cout %% 4 == 0 required%s");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(x) || !al16(w) || !al16(y) || (bias && !al16(bias))) return cfail(DATR_CONV_ERR_ALIGNMENT, "buffers must be 16-byte aligned%s");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ConvShape cs;
  cs.N = N; cs.H = H; cs.W = W; cs.Cin = Cin; cs.Cout = Cout; cs.stride = stride;
  cs.Ho = (H + 2 - 3) / stride + 1;
  cs.Wo = (W + 2 - 3) / stride + 1;
  cs.tiles_x = (cs.Wo + TW - 1) / TW;
  cs.tiles_y = (cs.Ho + TH - 1) / TH;
  const int BN = Cout > 128 ? 256 : (Cout > 64 ? 128 : 64);
  cs.n_tiles = (Cout + BN - 1) / BN;

  CUtensorMap mx, mw;
  {  // activations [N, H, W, Cin]: box {32 channels, 16 pixels, 8 rows, 1 image} with element stride = conv stride
    const cuuint64_t gdim[4] = {cuuint64_t(Cin), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
    const cuuint64_t gstr[3] = {cuuint64_t(Cin) * 4, cuuint64_t(W) * Cin * 4, cuuint64_t(H) * W * Cin * 4};
    const cuuint32_t box[4] = {cuuint32_t(BK), cuuint32_t((TW - 1) * stride + 1), cuuint32_t((TH - 1) * stride + 1), 1};
    const cuuint32_t estr[4] = {1, cuuint32_t(stride), cuuint32_t(stride), 1};
    if (int rc = encode(&mx, x, 4, gdim, gstr, box, estr)) return rc;
  }
  {  // weights [Cout, 3, 3, Cin] seen as [Cout, 9*Cin]
    const cuuint64_t gdim[2] = {cuuint64_t(9) * Cin, cuuint64_t(Cout)};
    const cuuint64_t gstr[1] = {cuuint64_t(9) * Cin * 4};
    const cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(BN)};
    const cuuint32_t estr[2] = {1, 1};
    if (int rc = encode(&mw, w, 2, gdim, gstr, box, estr)) return rc;
  }
  switch (BN) {
    case 256: return launch<256, 3>(mx, mw, bias, y, cs, relu, stream);
    case 128: return launch<128, 5>(mx, mw, bias, y, cs, relu, stream);
    default:  return launch<64, 6>(mx, mw, bias, y, cs, relu, stream);
  }
}

const char* datr_conv_last_error(void) { return g_conv_err; }
uint64_t datr_conv_launch_count(void) { return g_conv_launches.load(std::memory_order_relaxed); }

}  // extern "C"
