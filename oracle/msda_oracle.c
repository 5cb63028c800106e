/*
 * oracle/msda_oracle.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Plain-C CPU restatement of the reference's multi-scale deformable attention
 * arithmetic, written from the semantics of (paths relative to /root/reference):
 *   forward  : models/dino/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299
 *              (pixel coords loc*size-0.5, validity guard :288) + bilinear
 *              helper :33-84 (floor, 4 bounds-checked corners, hh*hw weights)
 *   backward : cuh:87-159 (grad_value scatter, grad_attn = top_grad*val,
 *              grad_loc = {W*grad_w, H*grad_h}*top_grad*attn) and the per-(l,p)
 *              reduction over channels of cuh:301-403.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  Parity pinning: checked against the reference's own
 * ms_deform_attn_core_pytorch (func.py:41-61) outputs and autograd gradients
 * stored under tests/golden/ (see tests/golden/make_golden.py).
 *
 * Layouts (all contiguous, row-major):
 *   value [N,S,M,D]  shapes int64 [L,2]=(H,W)  level_start int64 [L]
 *   loc   [N,Lq,M,L,P,2] (x,y)   attn [N,Lq,M,L,P]   out/grad_out [N,Lq,M,D]
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_ORACLE(T, SUF)                                                          \
  typedef struct {                                                                     \
    int ok;             /* sample inside (-1,H)x(-1,W) */                              \
    int64_t o[4];       /* element offsets of the 4 corners, -1 when out of map */     \
    T w[4];             /* bilinear weights, corner order (lo,lo)(lo,hi)(hi,lo)(hi,hi) */ \
    T lh, lw, hh, hw;                                                                  \
  } tap_##SUF;                                                                         \
                                                                                       \
  static tap_##SUF locate_##SUF(T lx, T ly, int64_t H, int64_t W, int64_t rowstride) { \
    tap_##SUF t;                                                                       \
    memset(&t, 0, sizeof t);                                                           \
    const T y = (T)(ly * (T)H - (T)0.5);                                               \
    const T x = (T)(lx * (T)W - (T)0.5);                                               \
    t.ok = (y > (T)-1 && x > (T)-1 && y < (T)H && x < (T)W);                           \
    if (!t.ok) return t;                                                               \
    const int64_t y0 = (int64_t)floor((double)y), x0 = (int64_t)floor((double)x);      \
    const int64_t y1 = y0 + 1, x1 = x0 + 1;                                            \
    t.lh = y - (T)y0; t.lw = x - (T)x0; t.hh = (T)1 - t.lh; t.hw = (T)1 - t.lw;        \
    t.w[0] = t.hh * t.hw; t.w[1] = t.hh * t.lw; t.w[2] = t.lh * t.hw; t.w[3] = t.lh * t.lw; \
    t.o[0] = (y0 >= 0 && x0 >= 0)         ? (y0 * W + x0) * rowstride : -1;            \
    t.o[1] = (y0 >= 0 && x1 <= W - 1)     ? (y0 * W + x1) * rowstride : -1;            \
    t.o[2] = (y1 <= H - 1 && x0 >= 0)     ? (y1 * W + x0) * rowstride : -1;            \
    t.o[3] = (y1 <= H - 1 && x1 <= W - 1) ? (y1 * W + x1) * rowstride : -1;            \
    return t;                                                                          \
  }                                                                                    \
                                                                                       \
  void msda_oracle_fwd_##SUF(const T* value, const int64_t* shapes,                    \
                             const int64_t* lvl_start, const T* loc, const T* attn,    \
                             int N, int S, int M, int D, int L, int Lq, int P, T* out) { \
    const int64_t rs = (int64_t)M * D;                                                 \
    _Pragma("omp parallel for collapse(2) schedule(static)")                           \
    for (int b = 0; b < N; ++b)                                                        \
      for (int q = 0; q < Lq; ++q)                                                     \
        for (int m = 0; m < M; ++m) {                                                  \
          const int64_t row = ((int64_t)b * Lq + q) * M + m;                           \
          T* o = out + row * D;                                                        \
          for (int c = 0; c < D; ++c) o[c] = 0;                                        \
          for (int l = 0; l < L; ++l) {                                                \
            const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];                    \
            const T* vbase = value + ((int64_t)b * S + lvl_start[l]) * rs + (int64_t)m * D; \
            for (int p = 0; p < P; ++p) {                                              \
              const int64_t k = (row * L + l) * P + p;                                 \
              const tap_##SUF t = locate_##SUF(loc[2 * k], loc[2 * k + 1], H, W, rs);  \
              if (!t.ok) continue;                                                     \
              const T a = attn[k];                                                     \
              for (int c = 0; c < D; ++c) {                                            \
                T v[4];                                                                \
                for (int i = 0; i < 4; ++i) v[i] = t.o[i] >= 0 ? vbase[t.o[i] + c] : (T)0; \
                o[c] += (t.w[0] * v[0] + t.w[1] * v[1] + t.w[2] * v[2] + t.w[3] * v[3]) * a; \
              }                                                                        \
            }                                                                          \
          }                                                                            \
        }                                                                              \
  }                                                                                    \
                                                                                       \
  /* grad_value/grad_loc/grad_attn are overwritten (zero-filled here first). */        \
  void msda_oracle_bwd_##SUF(const T* value, const int64_t* shapes,                    \
                             const int64_t* lvl_start, const T* loc, const T* attn,    \
                             const T* grad_out, int N, int S, int M, int D, int L,     \
                             int Lq, int P, T* grad_value, T* grad_loc, T* grad_attn) { \
    const int64_t rs = (int64_t)M * D;                                                 \
    memset(grad_value, 0, sizeof(T) * (size_t)N * S * M * D);                          \
    memset(grad_loc, 0, sizeof(T) * (size_t)N * Lq * M * L * P * 2);                   \
    memset(grad_attn, 0, sizeof(T) * (size_t)N * Lq * M * L * P);                      \
    /* (b,m) pairs never share a grad_value element -> race-free parallel units */     \
    _Pragma("omp parallel for collapse(2) schedule(static)")                           \
    for (int b = 0; b < N; ++b)                                                        \
      for (int m = 0; m < M; ++m)                                                      \
        for (int q = 0; q < Lq; ++q) {                                                 \
          const int64_t row = ((int64_t)b * Lq + q) * M + m;                           \
          const T* g = grad_out + row * D;                                             \
          for (int l = 0; l < L; ++l) {                                                \
            const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];                    \
            const int64_t base = ((int64_t)b * S + lvl_start[l]) * rs + (int64_t)m * D; \
            for (int p = 0; p < P; ++p) {                                              \
              const int64_t k = (row * L + l) * P + p;                                 \
              const tap_##SUF t = locate_##SUF(loc[2 * k], loc[2 * k + 1], H, W, rs);  \
              if (!t.ok) continue;                                                     \
              const T a = attn[k];                                                     \
              T ga = 0, gx = 0, gy = 0;                                                \
              for (int c = 0; c < D; ++c) {                                            \
                const T top = g[c], tv = top * a;                                      \
                T v[4];                                                                \
                for (int i = 0; i < 4; ++i) {                                          \
                  v[i] = 0;                                                            \
                  if (t.o[i] >= 0) {                                                   \
                    v[i] = value[base + t.o[i] + c];                                   \
                    grad_value[base + t.o[i] + c] += t.w[i] * tv;                      \
                  }                                                                    \
                }                                                                      \
                const T gh = -t.hw * v[0] - t.lw * v[1] + t.hw * v[2] + t.lw * v[3];   \
                const T gw = -t.hh * v[0] + t.hh * v[1] - t.lh * v[2] + t.lh * v[3];   \
                ga += top * (t.w[0] * v[0] + t.w[1] * v[1] + t.w[2] * v[2] + t.w[3] * v[3]); \
                gx += (T)W * gw * tv;                                                  \
                gy += (T)H * gh * tv;                                                  \
              }                                                                        \
              grad_attn[k] = ga; grad_loc[2 * k] = gx; grad_loc[2 * k + 1] = gy;       \
            }                                                                          \
          }                                                                            \
        }                                                                              \
  }

DEFINE_ORACLE(float, f32)
DEFINE_ORACLE(double, f64)
