// Deterministic merge of per-shard winner rows (shared by k_merge_topk and the fused peer exchange).
#pragma once
#include <climits>

#include "common.cuh"

namespace vzgp {

// rows [n_rows x width], row = [score, global index, features...]  ->  out [count x width]: larger score
// first, NaN as -inf, ties -> lower global index, rows with index < 0 last (the ordering of
// vectorized_base.py:575-598 applied to one pool that happens to live on several GPUs).  One CTA of 256
// threads; n_rows <= 2048, count <= 256.  Selection by repeated arg-max.  UNCACHED: read the rows with
// ld.volatile (they were written by peer GPUs over NVLink, bypass this SM's L1).
template <bool UNCACHED>
__device__ __forceinline__ double merge_load(const double* p) {
  if (UNCACHED) return __ldcv(p);
  return *p;
}

template <bool UNCACHED>
__device__ __forceinline__ void merge_topk_block(const double* rows, int n_rows, int width, int count, double* out) {
  __shared__ double sv[2048];
  __shared__ long long si[2048];
  __shared__ double rv[256];
  __shared__ long long ri[256];
  __shared__ int rr[256];
  const int tid = threadIdx.x;
  for (int r = tid; r < n_rows; r += 256) {
    double v = merge_load<UNCACHED>(rows + (size_t)r * width);
    const double gi = merge_load<UNCACHED>(rows + (size_t)r * width + 1);
    if (v != v) v = -INFINITY;
    sv[r] = v;
    si[r] = (gi < 0.0) ? LLONG_MAX : (long long)gi;
  }
  __syncthreads();
  for (int c = 0; c < count; ++c) {
    double bv = -INFINITY; long long bi = LLONG_MAX; int br = -1;
    for (int r = tid; r < n_rows; r += 256) {
      const long long i = si[r];
      if (i == LLONG_MAX - 1) continue;            // already taken
      const double v = sv[r];
      if (br < 0 || v > bv || (v == bv && i < bi)) { bv = v; bi = i; br = r; }
    }
    rv[tid] = bv; ri[tid] = bi; rr[tid] = br;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
      if (tid < s2) {
        const int o = tid + s2;
        const bool take = rr[o] >= 0 && (rr[tid] < 0 || rv[o] > rv[tid] || (rv[o] == rv[tid] && ri[o] < ri[tid]));
        if (take) { rv[tid] = rv[o]; ri[tid] = ri[o]; rr[tid] = rr[o]; }
      }
      __syncthreads();
    }
    const int win = rr[0];
    if (win >= 0) {
      for (int d = tid; d < width; d += 256) out[(size_t)c * width + d] = merge_load<UNCACHED>(rows + (size_t)win * width + d);
      if (tid == 0) si[win] = LLONG_MAX - 1;
    } else {
      for (int d = tid; d < width; d += 256) out[(size_t)c * width + d] = (d == 0) ? -INFINITY : (d == 1 ? -1.0 : 0.0);
    }
    __syncthreads();
  }
}

}  // namespace vzgp
