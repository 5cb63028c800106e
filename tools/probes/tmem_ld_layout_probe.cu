// tmem_ld_layout_probe.cu -- which accumulator element does each thread receive from tcgen05.ld .16x256b / .16x128b?
// (The PTX ISA documents the fragment layouts with figures that are not available offline; this probe measures them.)
// One CTA, warps 0-3.  Warp w fills its 32 TMEM lanes with tcgen05.st.32x32b (lane = row, register j = column j,
// value = 1000 * row + column), then reads them back with the 16-lane shapes and prints (thread, register) -> (row, col).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tmem_ld_layout_probe tools/probes/tmem_ld_layout_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128) probe(float* out256, float* out128, float* out256x4) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (uint32_t(warp * 32) << 16);
  // fill: lane = row (32 * warp + lane), 32 columns
  {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(1000.f * float(warp * 32 + lane) + float(j));
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // .16x256b.x1 at lane offsets 0 and 16 of the warp's window: 4 registers each
  for (int h = 0; h < 2; ++h) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + (uint32_t(16 * h) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* o = out256 + ((warp * 2 + h) * 32 + lane) * 4;
    o[0] = __uint_as_float(r0); o[1] = __uint_as_float(r1); o[2] = __uint_as_float(r2); o[3] = __uint_as_float(r3);
  }
  // .16x128b.x1: 2 registers
  for (int h = 0; h < 2; ++h) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(base + (uint32_t(16 * h) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* o = out128 + ((warp * 2 + h) * 32 + lane) * 2;
    o[0] = __uint_as_float(r0); o[1] = __uint_as_float(r1);
  }
  // .16x256b.x4 (32 columns): 16 registers, lane offset 0 only
  {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(base) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out256x4[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(64) : "memory");
}

int main() {
  float *a, *b, *c;
  cudaMallocManaged(&a, 4 * 2 * 32 * 4 * 4); cudaMallocManaged(&b, 4 * 2 * 32 * 2 * 4); cudaMallocManaged(&c, 4 * 32 * 16 * 4);
  probe<<<1, 128>>>(a, b, c);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  for (int w = 0; w < 4; w += 3)
    for (int h = 0; h < 2; ++h) {
      printf("16x256b.x1 warp %d laneoff %d: thread -> (row,col) of r0 r1 r2 r3\n", w, 16 * h);
      for (int t = 0; t < 32; ++t) {
        const float* o = a + ((w * 2 + h) * 32 + t) * 4;
        printf("  t%02d:", t);
        for (int j = 0; j < 4; ++j) printf(" (%d,%d)", int(o[j]) / 1000, int(o[j]) % 1000);
        printf("\n");
      }
    }
  printf("16x128b.x1 warp 0 laneoff 0 / 16: thread -> (row,col) of r0 r1\n");
  for (int h = 0; h < 2; ++h)
    for (int t = 0; t < 32; ++t) {
      const float* o = b + ((0 * 2 + h) * 32 + t) * 2;
      printf("  h%d t%02d: (%d,%d) (%d,%d)\n", h, t, int(o[0]) / 1000, int(o[0]) % 1000, int(o[1]) / 1000, int(o[1]) % 1000);
    }
  printf("16x256b.x4 warp 1: thread -> (row,col) of r0..r15\n");
  for (int t = 0; t < 32; ++t) {
    printf("  t%02d:", t);
    for (int j = 0; j < 16; ++j) printf(" (%d,%d)", int(c[(1 * 32 + t) * 16 + j]) / 1000, int(c[(1 * 32 + t) * 16 + j]) % 1000);
    printf("\n");
  }
  return 0;
}
