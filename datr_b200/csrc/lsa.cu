// lsa.cu -- rectangular linear sum assignment (Hungarian matching) on the GPU, one CTA per problem (sm_100a).
//
// The reference matches predictions to ground-truth boxes with scipy.optimize.linear_sum_assignment on the host
// (models/dino/matcher.py:91, once per prediction set: 7 device->host syncs per step; round 1 of this repository batched
// them into ONE read-back).  That read-back was the last host synchronisation of the training step: the host could not
// enqueue ahead of it, and everything behind it (criterion, start of the backward) was exposed launch latency -- at N > 1
// GPUs also inter-rank skew.  The problems are tiny (900 queries x a few dozen boxes, 14 per step), so they are solved
// where the cost matrix already is.
//
// Algorithm = scipy's (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp, the modified Jonker-Volgenant shortest
// augmenting path method of D. F. Crouse, "On implementing 2D rectangular assignment algorithms", 2016), restated for one
// thread block: rows = ground-truth boxes (the short side; scipy transposes a tall matrix the same way), columns =
// queries; per row one Dijkstra-like search whose relaxation step and arg-min run over the columns in parallel.
// Arithmetic in fp64 on the fp32 cost matrix, in scipy's operation order, and scipy's tie rule (among equally short
// columns prefer an unassigned one, the LAST such in its `remaining` array order, else the FIRST) reproduced through the
// same swap-with-last bookkeeping of that array -- assignments are identical to the host solver's, ties included
// (tests/test_lsa_gpu.py compares them on random, tied and integer cost matrices).
// Output: per problem the matched (query, box) pairs sorted by query index, as scipy returns them, written as int64 into
// a flat buffer at host-chosen offsets ([queries..., boxes...] per problem).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "datr_lsa.h"

namespace {

thread_local char g_lsa_err[256] = "";
std::atomic<uint64_t> g_lsa_launches{0};

constexpr int kThreads = 256;

struct Cand { double val; int pos; int col; int free; };   // free: the column has no row yet

// scipy's scan: a later element replaces the best if strictly lower, or equal and unassigned  =>  among the minima: the
// unassigned one with the LARGEST position if any, else the one with the SMALLEST position
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
  if (a.val < b.val) return true;
  if (a.val > b.val) return false;
  if (a.free != b.free) return a.free > b.free;
  return a.free ? a.pos > b.pos : a.pos < b.pos;
}

__device__ __forceinline__ Cand shfl_down(const Cand& c, int d) {
  Cand r;
  r.val = __shfl_down_sync(0xffffffffu, c.val, d);
  r.pos = __shfl_down_sync(0xffffffffu, c.pos, d);
  r.col = __shfl_down_sync(0xffffffffu, c.col, d);
  r.free = __shfl_down_sync(0xffffffffu, c.free, d);
  return r;
}

// problems: int64 [P, 5] = {element offset of the problem's first cost entry, elements between consecutive queries,
//                           number of queries nq, number of boxes nt, offset of its output in `out`}
// cost entry of (query q, box t) = cost[offset + q * stride + t]
__global__ void __launch_bounds__(kThreads)
lsa_kernel(const float* __restrict__ cost, const int64_t* __restrict__ problems, int64_t* __restrict__ out, int max_nq,
           int max_nt) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int64_t* pr = problems + 5 * blockIdx.x;
  const float* c = cost + pr[0];
  const int64_t stride = pr[1];
  const int nc = int(pr[2]), nr = int(pr[3]);      // columns = queries, rows = boxes
  int64_t* o = out + pr[4];
  if (nr == 0) return;
  double* u = reinterpret_cast<double*>(smem);                 // [max_nt]
  double* v = u + max_nt;                                      // [max_nq]
  double* shortest = v + max_nq;                               // [max_nq]
  int* path = reinterpret_cast<int*>(shortest + max_nq);       // [max_nq]
  int* row4col = path + max_nq;                                // [max_nq]
  int* remaining = row4col + max_nq;                           // [max_nq]
  int* pos = remaining + max_nq;                               // [max_nq]
  int* col4row = pos + max_nq;                                 // [max_nt]
  unsigned char* SC = reinterpret_cast<unsigned char*>(col4row + max_nt);   // [max_nq]
  unsigned char* SR = SC + max_nq;                             // [max_nt]
  __shared__ Cand warp_best[kThreads / 32];
  __shared__ int s_i, s_sink, s_num_remaining;
  __shared__ double s_min;
  const int tid = threadIdx.x;

  for (int j = tid; j < nc; j += kThreads) { v[j] = 0.0; row4col[j] = -1; }
  for (int i = tid; i < nr; i += kThreads) { u[i] = 0.0; col4row[i] = -1; }
  __syncthreads();

  for (int cur = 0; cur < nr; ++cur) {
    for (int j = tid; j < nc; j += kThreads) {
      remaining[j] = nc - j - 1; pos[j] = nc - j - 1; SC[j] = 0; shortest[j] = CUDART_INF;
    }
    for (int i = tid; i < nr; i += kThreads) SR[i] = 0;
    if (tid == 0) { s_i = cur; s_sink = -1; s_num_remaining = nc; s_min = 0.0; }
    __syncthreads();
    while (true) {
      const int i = s_i;
      const double min_val = s_min, ui = u[i];
      if (tid == 0) SR[i] = 1;
      Cand best = {CUDART_INF, 0x7fffffff, -1, 0};
      for (int j = tid; j < nc; j += kThreads) {
        if (SC[j]) continue;
        const double r = ((min_val + double(c[(int64_t)j * stride + i])) - ui) - v[j];
        double sj = shortest[j];
        if (r < sj) { path[j] = i; shortest[j] = r; sj = r; }
        const Cand cand = {sj, pos[j], j, row4col[j] == -1 ? 1 : 0};
        if (best.col < 0 || better(cand, best)) best = cand;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const Cand other = shfl_down(best, d);
        if (other.col >= 0 && (best.col < 0 || better(other, best))) best = other;
      }
      if ((tid & 31) == 0) warp_best[tid >> 5] = best;
      __syncthreads();
      if (tid == 0) {
        Cand b = warp_best[0];
        for (int w = 1; w < kThreads / 32; ++w)
          if (warp_best[w].col >= 0 && (b.col < 0 || better(warp_best[w], b))) b = warp_best[w];
        // (an all-infinite row would be scipy's "infeasible"; the cost matrices of the matcher are finite)
        s_min = b.val;
        const int j = b.col;
        if (row4col[j] == -1) s_sink = j; else s_i = row4col[j];
        SC[j] = 1;
        const int idx = pos[j], last = remaining[--s_num_remaining];
        remaining[idx] = last; pos[last] = idx;
      }
      __syncthreads();
      if (s_sink != -1) break;
    }
    // dual variables
    const double min_val = s_min;
    if (tid == 0) u[cur] += min_val;
    for (int i = tid; i < nr; i += kThreads)
      if (SR[i] && i != cur) u[i] += min_val - shortest[col4row[i]];
    for (int j = tid; j < nc; j += kThreads)
      if (SC[j]) v[j] -= min_val - shortest[j];
    __syncthreads();
    // augment the previous solution along the path
    if (tid == 0) {
      int j = s_sink;
      while (true) {
        const int i = path[j];
        row4col[j] = i;
        const int t = col4row[i]; col4row[i] = j; j = t;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }
  // pairs sorted by query index (scipy: argsort of col4row after the transposition)
  for (int i = tid; i < nr; i += kThreads) {
    const int q = col4row[i];
    int rank = 0;
    for (int k = 0; k < nr; ++k) rank += col4row[k] < q;
    o[rank] = q;
    o[nr + rank] = i;
  }
}

size_t smem_bytes(int max_nq, int max_nt) {
  return size_t(max_nt) * 8 + size_t(max_nq) * 16 + size_t(max_nq) * 16 + size_t(max_nt) * 4 + size_t(max_nq) + size_t(max_nt) + 16;
}

}  // namespace

extern "C" {

int datr_lsa_solve(const float* cost, const int64_t* problems, int n_problems, int max_queries, int max_boxes, int64_t* out,
                   void* stream_) {
  if (!cost || !problems || !out || n_problems <= 0 || max_queries <= 0 || max_boxes < 0) {
    snprintf(g_lsa_err, sizeof g_lsa_err, "datr_lsa_solve: null pointer or non-positive size");
    return -1;
  }
  if (max_boxes > max_queries) {
    snprintf(g_lsa_err, sizeof g_lsa_err, "datr_lsa_solve: more boxes than queries (the short side must be the boxes)");
    return -4;
  }
  const size_t smem = smem_bytes(max_queries, max_boxes > 0 ? max_boxes : 1);
  if (smem > 200 * 1024) {
    snprintf(g_lsa_err, sizeof g_lsa_err, "datr_lsa_solve: problem too large for one thread block (%zu bytes of shared memory)", smem);
    return -4;
  }
  static std::atomic<size_t> opted{0};
  if (smem > 48 * 1024 && opted.load(std::memory_order_acquire) < smem) {
    const cudaError_t e = cudaFuncSetAttribute(lsa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { snprintf(g_lsa_err, sizeof g_lsa_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -3; }
    opted.store(200 * 1024, std::memory_order_release);
  }
  lsa_kernel<<<unsigned(n_problems), kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(cost, problems, out, max_queries,
                                                                                         max_boxes > 0 ? max_boxes : 1);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_lsa_err, sizeof g_lsa_err, "lsa_kernel launch: %s", cudaGetErrorString(e)); return -3; }
  g_lsa_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

const char* datr_lsa_last_error(void) { return g_lsa_err; }
uint64_t datr_lsa_launch_count(void) { return g_lsa_launches.load(std::memory_order_relaxed); }

}  // extern "C"
