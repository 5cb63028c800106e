// gemm_2cta_probe.cu -- standalone probe of a CTA-pair (cta_group::2) TF32 GEMM mainloop for sm_100a.
// NOT part of libdatr_b200.so and NOT yet run on a GPU (written after the round's GPU budget was spent): it is the
// starting point for item 1 of DESIGN.md section 8.  Build / run on the GPU box:
//    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Idatr_b200/csrc -o build/gemm_2cta_probe \
//         tools/probes/gemm_2cta_probe.cu -lcuda && timeout 60 build/gemm_2cta_probe
// Every mbarrier wait is bounded (datr_tc::mbar_wait traps after 2^26 polls), so a protocol error fails the launch
// instead of hanging the GPU.
//
// Y[M,N] = X[M,K] * W[N,K]^T, fp32 in HBM, TF32 products, fp32 accumulation in tensor memory.
// A cluster of two CTAs (one TPC) owns a 256 x 256 output tile:
//   * CTA r loads rows [128 r, 128 r + 128) of the X tile and rows [128 r, 128 r + 128) of the W tile (HALF of B) per
//     k-block -- 32 KB per CTA and k-block instead of the 48 KB of the single-CTA kernel (csrc/linear_tf32.cu), which is
//     what bounds that kernel (64 B/clk SM<-L2 port);
//   * both CTAs' TMA loads complete on the LEADER's full barrier (address with the peer bit cleared);
//   * the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256): rows 0-127 of the accumulator
//     land in the leader's tensor memory, rows 128-255 in the peer's; tcgen05.commit ... multicast::cluster frees
//     the stage in both CTAs and publishes the finished accumulator to both;
//   * each CTA's four epilogue warps read their own tensor memory and store their 128 rows.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tcgen05_common.cuh"

using namespace datr_tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int BN2 = 256;                 // N of the pair's tile; each CTA stages BN2 / 2 rows of W
constexpr int STAGES = 4;
constexpr int kA = BM * BK * 4;          // 16 KB: this CTA's 128 rows of X
constexpr int kBh = (BN2 / 2) * BK * 4;  // 16 KB: this CTA's half of the W tile
constexpr int kStage = kA + kBh;
constexpr int kThreads2 = 192;           // warp 0 TMA, warp 1 MMA (leader only) + TMEM alloc, warps 2-5 epilogue
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// TMA load whose completion bytes go to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on `bar` in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// instruction descriptor: D = fp32, A = B = TF32, K-major, N = 256, M = 256 (pair)
__host__ __device__ constexpr uint32_t tf32_idesc_pair() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN2 >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                 float* __restrict__ y, int M, int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* stage_epi = reinterpret_cast<float*>(smem + STAGES * kStage + 1024);      // 4 warps x 32 x 36 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;                          // which 256 x 256 tile
  const int n_tiles = (N + BN2 - 1) / BN2;
  const int m0 = (pair / n_tiles) * 256 + int(rank) * BM;    // this CTA's 128 rows of X / Y
  const int n0 = (pair % n_tiles) * BN2;
  const int kblocks = K / BK;

  if (warp == 0 && lane == 0) {
    // leader: full = 1 arrival (its own expect_tx) + the bytes of both CTAs' loads.  empty / acc_full get one multicast commit.
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc2(tmem_slot, BN2);                // issued by one warp of EACH CTA of the pair
  tc_fence_before();
  cluster_sync();                                            // barrier inits + allocation visible in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t s = kb % STAGES;
        mbar_wait(empty + s, ((kb / STAGES) & 1) ^ 1);       // own copy: released by the leader's multicast commit
        if (leader) mbar_expect_tx(full + s, 2 * kStage);    // bytes of both CTAs
        unsigned char* a = smem + s * kStage;
        tma_load_2d_pair(a, &tma_x, kb * BK, m0, full + s);
        tma_load_2d_pair(a + kA, &tma_w, kb * BK, n0 + int(rank) * (BN2 / 2), full + s);
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = tf32_idesc_pair();
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t s = kb % STAGES;
        mbar_wait(full + s, (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * kStage);
        const uint64_t ad = kmajor_sw128_desc(a), bd = kmajor_sw128_desc(a + kA);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_tf32_pair(tmem_base, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0);
        umma_commit_pair(empty + s);
      }
      umma_commit_pair(acc_full);
    }
  } else {
    // epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 of THIS CTA = rows m0 + 32 (w % 4) ..; all 256 columns
    const int lane_base = (warp & 3) * 32;
    float* tile_s = stage_epi + (warp - 2) * (kStageTile / 4);
    const int tr = lane >> 3, tc = (lane & 7) * 4;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int c = 0; c < BN2; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + uint32_t(c) + (uint32_t(lane_base) << 16), v);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(tile_s + lane * kStagePitch + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = m0 + lane_base + tr + 4 * j, col = n0 + c + tc;
        const float4 o = *reinterpret_cast<const float4*>(tile_s + (tr + 4 * j) * kStagePitch + tc);
        if (row < M && col + 4 <= N) *reinterpret_cast<float4*>(y + (size_t)row * N + col) = o;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  cluster_sync();                                            // both CTAs are done with tensor memory and the peer's barriers
  if (warp == 1) tmem_dealloc2(tmem_base, BN2);
}

static int make_map(CUtensorMap* map, const float* base, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return 1;
  const cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(cols) * 4};
  const cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

static float run(const float* x, const float* w, float* y, int M, int N, int K, int iters) {
  CUtensorMap mx, mw;
  if (make_map(&mx, x, M, K, BM) || make_map(&mw, w, N, K, BN2 / 2)) { printf("tensor map encode failed\n"); exit(1); }
  const int smem = STAGES * kStage + 1024 + 4 * kStageTile + 1024;
  CK(cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int pairs = ((M + 255) / 256) * ((N + BN2 - 1) / BN2);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  gemm_pair_kernel<<<2 * pairs, kThreads2, smem>>>(mx, mw, y, M, N, K);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) gemm_pair_kernel<<<2 * pairs, kThreads2, smem>>>(mx, mw, y, M, N, K);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms * 1e3f / iters;
}

int main() {
  {   // correctness at a small size against a double-precision host reference (TF32 bar: 1e-2 of the largest value)
    const int M = 512, N = 512, K = 256;
    std::vector<float> hx((size_t)M * K), hw((size_t)N * K), hy((size_t)M * N);
    uint32_t s = 12345u;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return float(int(s >> 9) % 2001 - 1000) / 1000.0f; };
    for (auto& v : hx) v = rnd();
    for (auto& v : hw) v = rnd();
    float *x, *w, *y;
    CK(cudaMalloc(&x, hx.size() * 4)); CK(cudaMalloc(&w, hw.size() * 4)); CK(cudaMalloc(&y, hy.size() * 4));
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(y, 0xff, hy.size() * 4));
    run(x, w, y, M, N, K, 1);
    CK(cudaMemcpy(hy.data(), y, hy.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, big = 0;
    for (int i = 0; i < M; i += 7)
      for (int j = 0; j < N; j += 5) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += double(hx[(size_t)i * K + k]) * double(hw[(size_t)j * K + k]);
        worst = fmax(worst, fabs(ref - double(hy[(size_t)i * N + j])));
        big = fmax(big, fabs(ref));
      }
    printf("correctness %dx%dx%d: max |err| %.3e of max |ref| %.3e -> %s\n", M, N, K, worst, big, worst < 1e-2 * big ? "OK" : "MISMATCH");
    CK(cudaFree(x)); CK(cudaFree(w)); CK(cudaFree(y));
  }
  {   // the encoder FFN linear1 shape: single-CTA kernel 115-119 us (profiles/r01l_bench_linear_and_wgrad.txt), HBM time ~63 us
    const int M = 44446, N = 2048, K = 256;
    float *x, *w, *y;
    CK(cudaMalloc(&x, (size_t)M * K * 4)); CK(cudaMalloc(&w, (size_t)N * K * 4)); CK(cudaMalloc(&y, (size_t)M * N * 4));
    CK(cudaMemset(x, 0, (size_t)M * K * 4)); CK(cudaMemset(w, 0, (size_t)N * K * 4));
    const float us = run(x, w, y, M, N, K, 20);
    printf("M=%d N=%d K=%d: %.1f us  (%.0f TFLOP/s, %.0f GB/s of compulsory traffic)\n", M, N, K, us,
           2.0 * M * N * K / us / 1e6, 4.0 * ((double)M * K + (double)N * K + (double)M * N) / us / 1e3);
  }
  return 0;
}
