// decoder_ops.cu -- small fused kernels of the DINO decoder layer loop (sm_100a), each replacing a chain of ATen launches
// of 1-15 us inside the captured decoder segment.
//
//   datr_sine_embed: gen_sineembed_for_position (reference models/dino/utils.py:gen_sineembed_for_position, called from
//     deformable_transformer.py:TransformerDecoder.forward once per layer on the [N, nq, 4] reference boxes):
//       out[r, blk * 128 + 2j + {0, 1}] = {sin, cos}(pos[r, c(blk)] * 2 pi / dim_t[2j + {0, 1}]),  c = (1, 0, 2, 3) = (y, x, w, h)
//     with dim_t = 10000 ** (2 * (i // 2) / 128) handed in as a table computed by torch itself, the same operation order
//     (multiply by 2 pi, divide by dim_t, sinf / cosf) and therefore the same bits as the 14 ATen kernels it replaces.
//     The boxes carry no gradient (they are detached between layers, deformable_transformer.py:742 of the reference).
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_decoder_ops.h"

namespace {

thread_local char g_do_err[256] = "";
std::atomic<uint64_t> g_do_launches{0};

__global__ void __launch_bounds__(256)
sine_embed_kernel(const float* __restrict__ pos, const float* __restrict__ dim_t, long long rows, int k, float two_pi,
                  float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread = one (row, block, feature pair)
  const long long total = rows * k * 64;
  if (idx >= total) return;
  const int j = int(idx % 64);
  const int blk = int((idx / 64) % k);
  const long long r = idx / (64LL * k);
  const int c = blk == 0 ? 1 : (blk == 1 ? 0 : blk);
  const float x = __ldg(pos + r * k + c) * two_pi;
  const float s = sinf(x / __ldg(dim_t + 2 * j));
  const float co = cosf(x / __ldg(dim_t + 2 * j + 1));
  reinterpret_cast<float2*>(out)[idx] = make_float2(s, co);
}

}  // namespace

extern "C" {

int datr_sine_embed(const float* pos, const float* dim_t, long long rows, int k, float* out, void* stream) {
  if (!pos || !dim_t || !out || rows <= 0 || (k != 2 && k != 4)) {
    snprintf(g_do_err, sizeof g_do_err, "datr_sine_embed: null pointer, rows <= 0 or k not in {2, 4}");
    return -1;
  }
  const long long total = rows * k * 64;
  const long long ctas = (total + 255) / 256;
  if (ctas > 0x7fffffffLL) { snprintf(g_do_err, sizeof g_do_err, "datr_sine_embed: problem too large"); return -1; }
  sine_embed_kernel<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(pos, dim_t, rows, k, 6.283185307179586f, out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_do_err, sizeof g_do_err, "sine_embed_kernel launch: %s", cudaGetErrorString(e)); return -3; }
  g_do_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

const char* datr_decoder_ops_last_error(void) { return g_do_err; }
uint64_t datr_decoder_ops_launch_count(void) { return g_do_launches.load(std::memory_order_relaxed); }

}  // extern "C"
