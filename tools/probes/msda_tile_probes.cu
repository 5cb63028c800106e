// msda_tile_probes.cu -- standalone micro-probes for the MSDeformAttn redesign costed in DESIGN.md section 7/8
// (NOT part of libdatr_b200.so; build and run on the GPU box:
//    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gpurun_out/msda_tile_probes tools/probes/msda_tile_probes.cu
//    gpurun_out/msda_tile_probes ).
//
// Question 1 (backward): the encoder backward spends its time pushing 22.8 M 128-byte `red.global.add.v4.f32` lines
// through the SM->L2 request path (447 us, profiles/r01b_msda_variants.txt).  Would accumulating a 16x16-query tile's
// gradient window (26x26 pixels x 32 channels = 86.5 KB) in shared memory with CAS-loop float2 adds
// (ATOMS.CAST.SPIN.64: sm_100a has no native fp32 shared-memory reduction) and flushing it once per tile be faster?
//   probe A: the red.global pattern of today's kernel            (one line per quarter-warp per corner)
//   probe B: the same contributions into a shared-memory window + one red.global per window line at the end
// Question 2 (forward): the forward is bound by the L1TEX multi-line replay rate, 2.1 cycles per 128-byte line.
//   probe C: LDG.128 gathers of 4 corner lines per sample from global memory (today)
//   probe D: the level window staged once per tile with cp.async, then LDS.128 gathers
// All probes use the config-2 encoder geometry of level 0 only (N=2, 8 heads, 100x167 queries sampling a 100x167 map,
// 16 samples per query = 4 "levels" x 4 points folded onto the same map, offsets |d| <= 4.5 px, hash-random) so that
// the line counts match one encoder call: 5.69 M samples, 22.8 M corner lines.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int H = 100, W = 167, M = 8, NB = 2, S = H * W, TQ = 16, TAPS = 16;
constexpr int HALO = 5, WIN = TQ + 2 * HALO;            // 26 x 26 window
constexpr int TILES_Y = (H + TQ - 1) / TQ, TILES_X = (W + TQ - 1) / TQ;

__host__ __device__ inline uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// anchor pixel of sample `t` of query (y, x): the query's own pixel plus a hash offset in [-4, 4]^2, clamped so that the
// 2x2 block stays inside the map (like the slot-table kernels)
__device__ __forceinline__ void anchor(int y, int x, int t, int m, int& by, int& bx) {
  const uint32_t h = mix((uint32_t)((y * W + x) * TAPS + t) * 8u + (uint32_t)m);
  by = min(max(y + int(h % 9u) - 4, 0), H - 2);
  bx = min(max(x + int((h >> 8) % 9u) - 4, 0), W - 2);
}

__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- probe A: today's pattern.  CTA = 32 consecutive queries of one head, 8 lanes per query, float4 per lane.
__global__ void __launch_bounds__(256, 4) probe_red_global(float* __restrict__ grad, const float* __restrict__ go) {
  const int m = blockIdx.x % M, sub = threadIdx.x & 7;
  const long long bq = (long long)(blockIdx.x / M) * 32 + (threadIdx.x >> 3);
  if (bq >= (long long)NB * S) return;
  const int b = int(bq / S), q = int(bq % S), y = q / W, x = q % W;
  const float4 g = *reinterpret_cast<const float4*>(go + (bq * M + m) * 32 + sub * 4);
  float* base = grad + ((long long)b * S * M + m) * 32 + sub * 4;
  for (int t = 0; t < TAPS; ++t) {
    int by, bx; anchor(y, x, t, m, by, bx);
    float* p = base + (long long)(by * W + bx) * (M * 32);
    red4(p, g); red4(p + M * 32, g); red4(p + (long long)W * M * 32, g); red4(p + (long long)(W + 1) * M * 32, g);
  }
}

// ---- probe B: CTA = one 16x16 query tile of one head; gradient window in shared memory, CAS-loop float2 adds.
__device__ __forceinline__ void smem_add2(float* p, float a, float b) {
  unsigned long long* u = reinterpret_cast<unsigned long long*>(p);
  unsigned long long old = *u, assumed;
  do {
    assumed = old;
    float2 v = *reinterpret_cast<float2*>(&assumed);
    v.x += a; v.y += b;
    old = atomicCAS(u, assumed, *reinterpret_cast<unsigned long long*>(&v));
  } while (old != assumed);
}

__global__ void __launch_bounds__(256, 2) probe_cas_window(float* __restrict__ grad, const float* __restrict__ go) {
  extern __shared__ __align__(16) float win[];                       // [WIN][WIN][32]
  const int m = blockIdx.x % M, tile = (blockIdx.x / M) % (TILES_Y * TILES_X), b = blockIdx.x / (M * TILES_Y * TILES_X);
  const int y0 = (tile / TILES_X) * TQ, x0 = (tile % TILES_X) * TQ, oy = y0 - HALO, ox = x0 - HALO;
  for (int i = threadIdx.x; i < WIN * WIN * 8; i += blockDim.x) reinterpret_cast<float4*>(win)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  for (int k = 0; k < TQ * TQ / 32; ++k) {
    const int qi = grp + 32 * k, y = y0 + qi / TQ, x = x0 + qi % TQ;
    if (y >= H || x >= W) continue;
    const long long bq = (long long)b * S + y * W + x;
    const float4 g = *reinterpret_cast<const float4*>(go + (bq * M + m) * 32 + sub * 4);
    for (int t = 0; t < TAPS; ++t) {
      int by, bx; anchor(y, x, t, m, by, bx);
      float* p = win + ((by - oy) * WIN + (bx - ox)) * 32 + sub * 4;      // |offset| <= 4 and HALO 5: always inside
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float* pc = p + ((c >> 1) * WIN + (c & 1)) * 32;
        smem_add2(pc, g.x, g.y); smem_add2(pc + 2, g.z, g.w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < WIN * WIN * 8; i += blockDim.x) {      // flush: one 128-byte line per 8 threads
    const int pix = i >> 3, wy = pix / WIN, wx = pix % WIN, gy = oy + wy, gx = ox + wx;
    if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
    red4(grad + (((long long)b * S + gy * W + gx) * M + m) * 32 + (i & 7) * 4, reinterpret_cast<float4*>(win)[i]);
  }
}

// ---- probe C: forward gathers from global memory (4 corner lines per sample, LDG.128 per lane).
__global__ void __launch_bounds__(256, 4) probe_gather_ldg(const float* __restrict__ value, float* __restrict__ out) {
  const int m = blockIdx.x % M, sub = threadIdx.x & 7;
  const long long bq = (long long)(blockIdx.x / M) * 32 + (threadIdx.x >> 3);
  if (bq >= (long long)NB * S) return;
  const int b = int(bq / S), q = int(bq % S), y = q / W, x = q % W;
  const float* base = value + ((long long)b * S * M + m) * 32 + sub * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < TAPS; ++t) {
    int by, bx; anchor(y, x, t, m, by, bx);
    const float* p = base + (long long)(by * W + bx) * (M * 32);
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), c = __ldg(reinterpret_cast<const float4*>(p + M * 32));
    const float4 d = __ldg(reinterpret_cast<const float4*>(p + (long long)W * M * 32)), e = __ldg(reinterpret_cast<const float4*>(p + (long long)(W + 1) * M * 32));
    acc.x += a.x + c.x + d.x + e.x; acc.y += a.y + c.y + d.y + e.y; acc.z += a.z + c.z + d.z + e.z; acc.w += a.w + c.w + d.w + e.w;
  }
  *reinterpret_cast<float4*>(out + (bq * M + m) * 32 + sub * 4) = acc;
}

// ---- probe D: the tile's window staged with cp.async (16 bytes per thread), then LDS.128 gathers.
__global__ void __launch_bounds__(256, 2) probe_gather_window(const float* __restrict__ value, float* __restrict__ out) {
  extern __shared__ __align__(16) float win[];
  const int m = blockIdx.x % M, tile = (blockIdx.x / M) % (TILES_Y * TILES_X), b = blockIdx.x / (M * TILES_Y * TILES_X);
  const int y0 = (tile / TILES_X) * TQ, x0 = (tile % TILES_X) * TQ, oy = y0 - HALO, ox = x0 - HALO;
  for (int i = threadIdx.x; i < WIN * WIN * 8; i += blockDim.x) {
    const int pix = i >> 3, gy = min(max(oy + pix / WIN, 0), H - 1), gx = min(max(ox + pix % WIN, 0), W - 1);
    const float* src = value + (((long long)b * S + gy * W + gx) * M + m) * 32 + (i & 7) * 4;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(win + i * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  for (int k = 0; k < TQ * TQ / 32; ++k) {
    const int qi = grp + 32 * k, y = y0 + qi / TQ, x = x0 + qi % TQ;
    if (y >= H || x >= W) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < TAPS; ++t) {
      int by, bx; anchor(y, x, t, m, by, bx);
      const float* p = win + ((by - oy) * WIN + (bx - ox)) * 32 + sub * 4;
      const float4 a = *reinterpret_cast<const float4*>(p), c = *reinterpret_cast<const float4*>(p + 32);
      const float4 d = *reinterpret_cast<const float4*>(p + WIN * 32), e = *reinterpret_cast<const float4*>(p + (WIN + 1) * 32);
      acc.x += a.x + c.x + d.x + e.x; acc.y += a.y + c.y + d.y + e.y; acc.z += a.z + c.z + d.z + e.z; acc.w += a.w + c.w + d.w + e.w;
    }
    const long long bq = (long long)b * S + y * W + x;
    *reinterpret_cast<float4*>(out + (bq * M + m) * 32 + sub * 4) = acc;
  }
}


// ---- probe E (written after A-D were measured; NOT yet run): per-tile pixel sort.  CTA = 16x8 query tile of one head.
// The tile's 2 048 samples are binned by anchor pixel with native integer shared atomics (count -> scan -> scatter of
// {query, 4 corner weights} records); then every 8-lane group OWNS window pixels and sums, in registers, the
// contributions of the four bins whose 2x2 block covers the pixel: one broadcast record read + one LDS.128 of the
// query's grad_out per contribution, no read-modify-write; one red.global per window line at the end.
constexpr int TQY = 8, WINX = TQ + 2 * HALO, WINY = TQY + 2 * HALO, NPIX = WINX * WINY;      // 26 x 18 = 468 pixels
constexpr int TILE_Q = TQ * TQY, TILE_S = TILE_Q * TAPS;                                       // 128 queries, 2 048 samples
constexpr int TILES_Y8 = (H + TQY - 1) / TQY;

__global__ void __launch_bounds__(256, 3) probe_pixel_sort(float* __restrict__ grad, const float* __restrict__ go) {
  extern __shared__ __align__(16) unsigned char raw[];
  float4* gos = reinterpret_cast<float4*>(raw);                         // [TILE_Q][8] grad_out of the tile's queries
  float4* rw = gos + TILE_Q * 8;                                        // [TILE_S] corner weights of a record
  int* rq = reinterpret_cast<int*>(rw + TILE_S);                        // [TILE_S] query of a record
  int* cnt = rq + TILE_S;                                               // [NPIX + 1] bin counts, then exclusive offsets
  const int m = blockIdx.x % M, tile = (blockIdx.x / M) % (TILES_Y8 * TILES_X), b = blockIdx.x / (M * TILES_Y8 * TILES_X);
  const int y0 = (tile / TILES_X) * TQY, x0 = (tile % TILES_X) * TQ, oy = y0 - HALO, ox = x0 - HALO;
  for (int i = threadIdx.x; i < TILE_Q * 8; i += blockDim.x) {
    const int qi = i >> 3, y = min(y0 + qi / TQ, H - 1), x = min(x0 + qi % TQ, W - 1);
    gos[i] = *reinterpret_cast<const float4*>(go + ((((long long)b * S + y * W + x) * M + m) * 32) + (i & 7) * 4);
  }
  for (int i = threadIdx.x; i <= NPIX; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  int bin[TILE_S / 256], rank[TILE_S / 256];
#pragma unroll
  for (int k = 0; k < TILE_S / 256; ++k) {                              // sample id -> (query, tap)
    const int sid = threadIdx.x + 256 * k, qi = sid / TAPS, y = y0 + qi / TQ, x = x0 + qi % TQ;
    bin[k] = -1;
    if (y < H && x < W) {
      int by, bx; anchor(y, x, sid % TAPS, m, by, bx);
      bin[k] = (by - oy) * WINX + (bx - ox);
      rank[k] = atomicAdd(cnt + bin[k], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {                                               // exclusive scan of NPIX counts by one warp
    int carry = 0;
    for (int base = 0; base < NPIX; base += 32) {
      const int i = base + threadIdx.x, v = i < NPIX ? cnt[i] : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if ((threadIdx.x & 31) >= d) inc += t; }
      if (i < NPIX) cnt[i] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (threadIdx.x == 0) cnt[NPIX] = carry;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < TILE_S / 256; ++k)
    if (bin[k] >= 0) {
      const int pos = cnt[bin[k]] + rank[k];
      rq[pos] = (threadIdx.x + 256 * k) / TAPS;
      rw[pos] = make_float4(1.f, 1.f, 1.f, 1.f);                        // (a real kernel stores attn * bilinear weights)
    }
  __syncthreads();
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  for (int pix = grp; pix < NPIX; pix += 32) {
    const int wy = pix / WINX, wx = pix % WINX;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool any = false;
#pragma unroll
    for (int c = 0; c < 4; ++c) {                                       // the bin whose corner c lands on this pixel
      const int ay = wy - (c >> 1), ax = wx - (c & 1);
      if (ay < 0 || ax < 0) continue;
      const int bb = ay * WINX + ax;
      for (int i = cnt[bb]; i < cnt[bb + 1]; ++i) {
        const float4 w4 = rw[i];
        const float w = c == 0 ? w4.x : c == 1 ? w4.y : c == 2 ? w4.z : w4.w;
        const float4 g = gos[rq[i] * 8 + sub];
        acc.x = fmaf(w, g.x, acc.x); acc.y = fmaf(w, g.y, acc.y); acc.z = fmaf(w, g.z, acc.z); acc.w = fmaf(w, g.w, acc.w);
        any = true;
      }
    }
    const int gy = oy + wy, gx = ox + wx;
    if (any && gy >= 0 && gy < H && gx >= 0 && gx < W)
      red4(grad + (((long long)b * S + gy * W + gx) * M + m) * 32 + sub * 4, acc);
  }
}

template <typename F>
float time_us(F launch, int iters = 20) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms * 1e3f / iters;
}

int main() {
  const size_t n = (size_t)NB * S * M * 32;
  float *value, *grad, *go, *out;
  CK(cudaMalloc(&value, n * 4)); CK(cudaMalloc(&grad, n * 4)); CK(cudaMalloc(&go, n * 4)); CK(cudaMalloc(&out, n * 4));
  CK(cudaMemset(value, 0, n * 4)); CK(cudaMemset(grad, 0, n * 4)); CK(cudaMemset(go, 0, n * 4));
  const int rows_ctas = ((NB * S + 31) / 32) * M, tile_ctas = NB * M * TILES_Y * TILES_X;
  const int smem = WIN * WIN * 32 * 4;
  CK(cudaFuncSetAttribute(probe_cas_window, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe_gather_window, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const double lines = (double)NB * S * M * TAPS * 4;
  printf("level-0-only geometry: %d samples, %.1f M corner lines, window %dx%d = %d bytes\n", NB * S * M * TAPS, lines / 1e6, WIN, WIN, smem);
  const float a = time_us([&] { probe_red_global<<<rows_ctas, 256>>>(grad, go); });
  const float b = time_us([&] { probe_cas_window<<<tile_ctas, 256, smem>>>(grad, go); });
  const float c = time_us([&] { probe_gather_ldg<<<rows_ctas, 256>>>(value, out); });
  const float d = time_us([&] { probe_gather_window<<<tile_ctas, 256, smem>>>(value, out); });
  const int smem_e = TILE_Q * 8 * 16 + TILE_S * 16 + TILE_S * 4 + (NPIX + 1) * 4;
  CK(cudaFuncSetAttribute(probe_pixel_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_e));
  const float e = time_us([&] { probe_pixel_sort<<<NB * M * TILES_Y8 * TILES_X, 256, smem_e>>>(grad, go); });
  CK(cudaGetLastError());
  printf("A red.global per corner line          %8.1f us  (%.2f cycles/line/SM at 1.965 GHz, 148 SMs)\n", a, a * 1965.0 * 148 / lines);
  printf("B shared-memory window (CAS) + flush  %8.1f us\n", b);
  printf("C LDG.128 gathers from global         %8.1f us  (%.2f cycles/line/SM)\n", c, c * 1965.0 * 148 / lines);
  printf("D cp.async window + LDS.128 gathers   %8.1f us\n", d);
  printf("E per-tile pixel sort + owned pixels  %8.1f us  (%d bytes of shared memory per CTA)\n", e, smem_e);
  return 0;
}
