// decoder_ops.cu -- small fused kernels of the DINO decoder layer loop (sm_100a), each replacing a chain of ATen launches
// of 1-15 us inside the captured decoder segment.
//
//   datr_sine_embed: gen_sineembed_for_position (reference models/dino/utils.py:gen_sineembed_for_position, called from
//     deformable_transformer.py:TransformerDecoder.forward once per layer on the [N, nq, 4] reference boxes):
//       out[r, blk * 128 + 2j + {0, 1}] = {sin, cos}(pos[r, c(blk)] * 2 pi / dim_t[2j + {0, 1}]),  c = (1, 0, 2, 3) = (y, x, w, h)
//     with dim_t = 10000 ** (2 * (i // 2) / 128) handed in as a table computed by torch itself, the same operation order
//     (multiply by 2 pi, divide by dim_t, sinf / cosf) and therefore the same bits as the 14 ATen kernels it replaces.
//     The boxes carry no gradient (they are detached between layers, deformable_transformer.py:742 of the reference).
//   datr_pos_embed_hw, datr_bn_relu_maxpool_nhwc: two more gradient-free chains outside the decoder (position encoding of
//     the feature maps, tail of the frozen ResNet stem); see the comments at the kernels.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>
#include <cmath>
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

// PositionEmbeddingSineHW (reference models/dino/position_encoding.py:62-107) after its cumulative sums: for every pixel
//   out[r, 2j + {0,1}]          = {sin, cos}(y[r] / dim_t_h[2j + {0,1}])        j < feats / 2
//   out[r, feats + 2j + {0,1}]  = {sin, cos}(x[r] / dim_t_w[2j + {0,1}])
// ([N, H, W, 2 * feats]: the layout the token flattening wants).  Replaces ~20 ATen launches per feature level.
__global__ void __launch_bounds__(256)
pos_embed_hw_kernel(const float* __restrict__ y, const float* __restrict__ x, const float* __restrict__ dim_t_h,
                    const float* __restrict__ dim_t_w, long long rows, int feats, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread = one (pixel, axis, feature pair)
  if (idx >= rows * feats) return;
  const int half = feats / 2;
  const int j = int(idx % half);
  const int axis = int((idx / half) % 2);
  const long long r = idx / feats;
  const float c = __ldg((axis ? x : y) + r);
  const float* t = axis ? dim_t_w : dim_t_h;
  reinterpret_cast<float2*>(out)[idx] = make_float2(sinf(c / __ldg(t + 2 * j)), cosf(c / __ldg(t + 2 * j + 1)));
}

// ResNet stem tail: FrozenBatchNorm2d + ReLU + MaxPool2d(3, stride 2, padding 1) on the NHWC output of the 7x7 convolution
// (reference backbone.py / torchvision resnet.py: bn1 -> relu -> maxpool), one pass instead of three: y = max over the 3x3
// window of relu(x * scale + shift).  The stem is frozen (no gradient).  One thread = 4 channels of one output pixel.
__global__ void __launch_bounds__(256)
bn_relu_maxpool_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, int N, int H,
                       int W, int C, int Ho, int Wo, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4n = C / 4;
  const long long total = (long long)N * Ho * Wo * c4n;
  if (idx >= total) return;
  const int c4 = int(idx % c4n);
  const int wo = int((idx / c4n) % Wo);
  const int ho = int((idx / ((long long)c4n * Wo)) % Ho);
  const int n = int(idx / ((long long)c4n * Wo * Ho));
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4), sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = 2 * ho - 1 + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int xx = 2 * wo - 1 + dx;
      if (xx < 0 || xx >= W) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((long long)n * H + yy) * W + xx) * C) + c4);
      m.x = fmaxf(m.x, fmaxf(fmaf(v.x, sc.x, sh.x), 0.f)); m.y = fmaxf(m.y, fmaxf(fmaf(v.y, sc.y, sh.y), 0.f));
      m.z = fmaxf(m.z, fmaxf(fmaf(v.z, sc.z, sh.z), 0.f)); m.w = fmaxf(m.w, fmaxf(fmaf(v.w, sc.w, sh.w), 0.f));
    }
  }
  reinterpret_cast<float4*>(out)[idx] = m;
}

}  // namespace

extern "C" {

int datr_pos_embed_hw(const float* y, const float* x, const float* dim_t_h, const float* dim_t_w, long long rows, int feats,
                      float* out, void* stream) {
  if (!y || !x || !dim_t_h || !dim_t_w || !out || rows <= 0 || feats <= 0 || (feats & 1)) {
    snprintf(g_do_err, sizeof g_do_err, "datr_pos_embed_hw: null pointer, rows <= 0 or odd feature count");
    return -1;
  }
  const long long ctas = (rows * feats + 255) / 256;
  if (ctas > 0x7fffffffLL) { snprintf(g_do_err, sizeof g_do_err, "datr_pos_embed_hw: problem too large"); return -1; }
  pos_embed_hw_kernel<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, dim_t_h, dim_t_w, rows, feats, out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_do_err, sizeof g_do_err, "pos_embed_hw_kernel launch: %s", cudaGetErrorString(e)); return -3; }
  g_do_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int datr_bn_relu_maxpool_nhwc(const float* x, const float* scale, const float* shift, int N, int H, int W, int C, float* out,
                              void* stream) {
  if (!x || !scale || !shift || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) {
    snprintf(g_do_err, sizeof g_do_err, "datr_bn_relu_maxpool_nhwc: null pointer, non-positive size or C not a multiple of 4");
    return -1;
  }
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)N * Ho * Wo * (C / 4);
  const long long ctas = (total + 255) / 256;
  if (ctas > 0x7fffffffLL) { snprintf(g_do_err, sizeof g_do_err, "datr_bn_relu_maxpool_nhwc: problem too large"); return -1; }
  bn_relu_maxpool_kernel<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, scale, shift, N, H, W, C, Ho, Wo, out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_do_err, sizeof g_do_err, "bn_relu_maxpool_kernel launch: %s", cudaGetErrorString(e)); return -3; }
  g_do_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

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
