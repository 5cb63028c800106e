// msda_tma_red_probe.cu -- standalone probe (NOT part of libdatr_b200.so): can the MSDeformAttn backward scatter leave
// the LSU?  Today every corner of every sample is one `red.global.add.v4.f32` line per quarter-warp; 22.8 M such lines per
// encoder call saturate the SM's L1TEX -> L2 request path (3.8 cycles per line), and the value gathers of
// the same kernel queue behind them in the same pipe.  The TMA unit has its own path to L2 and can reduce:
//   cp.reduce.async.bulk.tensor.5d.global.shared::cta.add   box = 32 channels x 1 head x 2 x 2 pixels  (512 bytes)
// i.e. ONE instruction per sample adds all four corner rows, with the bounds handling done by the tensor map.
//   probe A: red.global.add.v4.f32 per corner (today's scatter), no gathers
//   probe F: the four weighted rows staged in shared memory (4 x STS.128 per lane) + one TMA reduce per sample
//   probe H: today's backward pattern = LDG.128 gathers + red.global per corner
//   probe G: LDG.128 gathers + TMA reduce (do the two overlap?)
// Geometry as msda_tile_probes.cu: config-2 encoder, level 0 only (N=2, 8 heads, 100x167 queries, 16 samples per query
// with hash offsets |d| <= 4 px): 5.69 M samples, 22.8 M corner lines.  Build / run:
//    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/msda_tma_red_probe tools/probes/msda_tma_red_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int H = 100, W = 167, M = 8, NB = 2, S = H * W, TAPS = 16;

__host__ __device__ inline uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__host__ __device__ __forceinline__ void anchor(int y, int x, int t, int m, int& by, int& bx) {
  const uint32_t h = mix((uint32_t)((y * W + x) * TAPS + t) * 8u + (uint32_t)m);
  int ay = y + int(h % 9u) - 4, ax = x + int((h >> 8) % 9u) - 4;
  by = ay < 0 ? 0 : (ay > H - 2 ? H - 2 : ay);
  bx = ax < 0 ? 0 : (ax > W - 2 ? W - 2 : ax);
}

__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void tma_red_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// weights of the four corners of sample t (any deterministic positive numbers do)
__device__ __forceinline__ float4 corner_w(int t) { return make_float4(0.4f, 0.3f, 0.2f, 0.1f + 0.01f * t); }

template <bool kGather>
__global__ void __launch_bounds__(256, 4) probe_red(float* __restrict__ grad, const float* __restrict__ go,
                                                    const float* __restrict__ value, float* __restrict__ sink) {
  const int m = blockIdx.x % M, sub = threadIdx.x & 7;
  long long bq = (long long)(blockIdx.x / M) * 32 + (threadIdx.x >> 3);
  if (bq >= (long long)NB * S) return;
  const int b = int(bq / S), q = int(bq % S), y = q / W, x = q % W;
  const float4 g = *reinterpret_cast<const float4*>(go + (bq * M + m) * 32 + sub * 4);
  const long long off = ((long long)b * S * M + m) * 32 + sub * 4;
  float acc = 0.f;
  for (int t = 0; t < TAPS; ++t) {
    int by, bx; anchor(y, x, t, m, by, bx);
    const long long po = off + (long long)(by * W + bx) * (M * 32);
    const float4 w = corner_w(t);
    if (kGather) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(value + po)), c = __ldg(reinterpret_cast<const float4*>(value + po + M * 32));
      const float4 d = __ldg(reinterpret_cast<const float4*>(value + po + (long long)W * M * 32));
      const float4 e = __ldg(reinterpret_cast<const float4*>(value + po + (long long)(W + 1) * M * 32));
      acc += g.x * (a.x + c.x + d.x + e.x) + g.y * (a.y + c.y + d.y + e.y) + g.z * (a.z + c.z + d.z + e.z) + g.w * (a.w + c.w + d.w + e.w);
    }
    float* p = grad + po;
    red4(p, make_float4(w.x * g.x, w.x * g.y, w.x * g.z, w.x * g.w));
    red4(p + M * 32, make_float4(w.y * g.x, w.y * g.y, w.y * g.z, w.y * g.w));
    red4(p + (long long)W * M * 32, make_float4(w.z * g.x, w.z * g.y, w.z * g.z, w.z * g.w));
    red4(p + (long long)(W + 1) * M * 32, make_float4(w.w * g.x, w.w * g.y, w.w * g.z, w.w * g.w));
  }
  if (kGather) sink[(bq * M + m) * 8 + sub] = acc;
}

// TMA variant.  Per warp: kStages staging buffers of 4 rows x 512 bytes; lane sub == 0 of each row issues its row's box.
template <bool kGather, int kStages, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) probe_tma(const __grid_constant__ CUtensorMap map, const float* __restrict__ go,
                                                             const float* __restrict__ value, float* __restrict__ sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int m = blockIdx.x % M, sub = threadIdx.x & 7, r = threadIdx.x >> 3, warp = threadIdx.x >> 5;
  long long bq = (long long)(blockIdx.x / M) * 32 + r;
  const bool live = bq < (long long)NB * S;
  if (!live) bq = (long long)NB * S - 1;
  const int b = int(bq / S), q = int(bq % S), y = q / W, x = q % W;
  const float4 g = *reinterpret_cast<const float4*>(go + (bq * M + m) * 32 + sub * 4);
  const long long off = ((long long)b * S * M + m) * 32 + sub * 4;
  // staging: [warp][stage][row in warp][corner][32 floats]
  float* stage0 = reinterpret_cast<float*>(smem) + (size_t)warp * kStages * 4 * 128 + (r & 3) * 128 + sub * 4;
  float acc = 0.f;
#pragma unroll 1
  for (int t0 = 0; t0 < TAPS; t0 += kStages) {
    // the boxes issued kStages samples ago must have been read before their buffers are overwritten
    if (sub == 0) bulk_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kStages; ++j) {
      const int t = t0 + j;
      int by, bx; anchor(y, x, t, m, by, bx);
      const float4 w = corner_w(t);
      if (kGather) {
        const long long po = off + (long long)(by * W + bx) * (M * 32);
        const float4 a = __ldg(reinterpret_cast<const float4*>(value + po)), c = __ldg(reinterpret_cast<const float4*>(value + po + M * 32));
        const float4 d = __ldg(reinterpret_cast<const float4*>(value + po + (long long)W * M * 32));
        const float4 e = __ldg(reinterpret_cast<const float4*>(value + po + (long long)(W + 1) * M * 32));
        acc += g.x * (a.x + c.x + d.x + e.x) + g.y * (a.y + c.y + d.y + e.y) + g.z * (a.z + c.z + d.z + e.z) + g.w * (a.w + c.w + d.w + e.w);
      }
      float* st = stage0 + j * 4 * 128;
      *reinterpret_cast<float4*>(st) = make_float4(w.x * g.x, w.x * g.y, w.x * g.z, w.x * g.w);
      *reinterpret_cast<float4*>(st + 32) = make_float4(w.y * g.x, w.y * g.y, w.y * g.z, w.y * g.w);
      *reinterpret_cast<float4*>(st + 64) = make_float4(w.z * g.x, w.z * g.y, w.z * g.z, w.z * g.w);
      *reinterpret_cast<float4*>(st + 96) = make_float4(w.w * g.x, w.w * g.y, w.w * g.z, w.w * g.w);
    }
    fence_async_smem();
    __syncwarp();
    if (sub == 0 && live) {
#pragma unroll
      for (int j = 0; j < kStages; ++j) {
        int by, bx; anchor(y, x, t0 + j, m, by, bx);
        tma_red_5d(&map, stage0 + j * 4 * 128, 0, m, bx, by, b);
      }
      bulk_commit();
    }
  }
  if (sub == 0) bulk_wait_read<0>();
  if (kGather) sink[(bq * M + m) * 8 + sub] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
float time_it(F f, float* grad, size_t n, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaMemset(grad, 0, n * 4));
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best * 1e3f;
}

int main() {
  const size_t n = (size_t)NB * S * M * 32;
  float *value, *grad, *go, *sink, *ref;
  CK(cudaMalloc(&value, n * 4)); CK(cudaMalloc(&grad, n * 4)); CK(cudaMalloc(&go, n * 4)); CK(cudaMalloc(&sink, n)); CK(cudaMalloc(&ref, n * 4));
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = float(mix(uint32_t(i)) % 1000u) * 1e-3f - 0.5f;
  CK(cudaMemcpy(go, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(value, h.data(), n * 4, cudaMemcpyHostToDevice));

  void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  CUtensorMap map;
  const cuuint64_t gdim[5] = {32, M, W, H, NB};
  const cuuint64_t gstr[4] = {32 * 4, (cuuint64_t)M * 32 * 4, (cuuint64_t)W * M * 32 * 4, (cuuint64_t)S * M * 32 * 4};
  const cuuint32_t box[5] = {32, 1, 2, 2, 1}, estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, grad, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", int(cr)); return 1; }

  const unsigned ctas = unsigned(((long long)NB * S + 31) / 32 * M);
  const float tA = time_it([&] { probe_red<false><<<ctas, 256>>>(grad, go, value, sink); }, grad, n);
  CK(cudaMemcpy(ref, grad, n * 4, cudaMemcpyDeviceToDevice));
  printf("A  red.global.add.v4.f32 per corner, no gathers          %8.1f us\n", tA);
  const float tH = time_it([&] { probe_red<true><<<ctas, 256>>>(grad, go, value, sink); }, grad, n);
  printf("H  LDG.128 gathers + red.global per corner (today)       %8.1f us\n", tH);

  std::vector<float> hr(n), hg(n);
  CK(cudaMemcpy(hr.data(), ref, n * 4, cudaMemcpyDeviceToHost));
  auto check = [&](const char* tag) {
    CK(cudaMemcpy(hg.data(), grad, n * 4, cudaMemcpyDeviceToHost));
    double worst = 0, big = 0;
    for (size_t i = 0; i < n; ++i) { worst = fmax(worst, fabs(double(hg[i]) - hr[i])); big = fmax(big, fabs(double(hr[i]))); }
    printf("   %s max |diff| vs probe A = %.3g (max |ref| %.3g)\n", tag, worst, big);
  };
#define RUN_TMA(GATHER, STAGES, MINB, LABEL)                                                                  \
  {                                                                                                           \
    const size_t sm = (size_t)8 * STAGES * 4 * 512;                                                           \
    CK(cudaFuncSetAttribute(probe_tma<GATHER, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    const float t = time_it([&] { probe_tma<GATHER, STAGES, MINB><<<ctas, 256, sm>>>(map, go, value, sink); }, grad, n); \
    printf("%s stages %d, %d CTAs/SM                         %8.1f us\n", LABEL, STAGES, MINB, t);               \
    check(LABEL);                                                                                             \
  }
  RUN_TMA(false, 1, 4, "F  STS + TMA reduce per sample, no gathers,");
  RUN_TMA(false, 2, 4, "F  STS + TMA reduce per sample, no gathers,");
  RUN_TMA(false, 4, 4, "F  STS + TMA reduce per sample, no gathers,");
  RUN_TMA(false, 4, 2, "F  STS + TMA reduce per sample, no gathers,");
  RUN_TMA(true, 2, 4, "G  LDG.128 gathers + TMA reduce,            ");
  RUN_TMA(true, 4, 4, "G  LDG.128 gathers + TMA reduce,            ");
  RUN_TMA(true, 4, 2, "G  LDG.128 gathers + TMA reduce,            ");
  return 0;
}
