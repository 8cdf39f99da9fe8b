// 64x64 Cholesky + inverse in shared memory, shared by k_potf2_inv (linalg.cu: diagonal block of the
// blocked factorisation) and k_nll_grad_small (nll_small.cu: the whole model when N <= 64).
#pragma once
#include "common.cuh"

namespace vzgp {

#ifndef VZ_TSTAMP
#define VZ_TSTAMP(i) do {} while (0)
#endif
// The barrier between the phases of potf2_inv_64: the whole CTA by default; a CTA that runs it on a subset of
// its warps (the dataflow chain: 256 of 320 threads) defines its own before including this header.
#ifndef VZ_POTF2_SYNC
#define VZ_POTF2_SYNC() __syncthreads()
#endif

// Factor a (64x64, row stride 66, lower triangle + arbitrary diagonal-16-block upper parts, 16-blocks
// strictly above the diagonal ZERO) in place and write its inverse into x (must be ZERO on entry).
// t: [32][34] scratch, rd: [64] reciprocal diagonal, *s_bad is set to 1 on a non-positive / non-finite
// pivot (the factor then holds NaN).  256 threads, ends with a VZ_POTF2_SYNC().
//
// This is a pure latency chain (64 dependent pivots), organised around the critical path:
//  * four 16-column macro steps; the 16x16 diagonal piece is factored by ONE warp entirely in
//    registers (lane = row) in the square-root-free LDL^T form: the only long-latency operation on
//    the pivot-to-pivot chain is one reciprocal, and the column broadcasts (shuffles of the
//    UNSCALED column) overlap it.  The 16 columns are scaled by 1/sqrt(d_j) once, in parallel,
//    after the loop.
//  * the rows below are solved one thread per row (right-looking, 2 dependent ops per column),
//    the trailing part is updated by the whole CTA with 16-wide register tiles.
//  * the inverse: the four 16x16 diagonal blocks (one warp each, lane = column) and two
//    recursive-doubling levels  X21 = -B^-1 (C A^-1)  with fully unrolled dot products.
// 1/d to ~1 ulp without the IEEE division sequence (and its slow-path check): MUFU.RCP64H seed
// (`rcp.approx.ftz.f64`, ~20 bits) + two Newton steps.  This sits on the pivot-to-pivot chain, 64 times per
// block.  d is a Cholesky pivot: positive and far from the subnormal / overflow range when it matters.
__device__ __forceinline__ double rcp_newton(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

__device__ __forceinline__ void potf2_inv_64(double* a, double* x, double* t, double* rd, int* s_bad) {
  constexpr int LD = 66;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  VZ_TSTAMP(1);
#pragma unroll 1
  for (int mb = 0; mb < 4; ++mb) {
    const int c0 = 16 * mb;
    if (warp == 0) {
      // ---- 16x16 LDL^T in registers: lane l < 16 owns row c0 + l ----
      double v[16];
      const int row = c0 + (lane & 15);
#pragma unroll
      for (int k = 0; k < 16; k += 2) {   // lanes 16..31 only take part in the shuffles
        const double2 p = lane < 16 ? *reinterpret_cast<const double2*>(a + row * LD + c0 + k) : make_double2(0.0, 0.0);
        v[k] = p.x; v[k + 1] = p.y;
      }
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const double d = __shfl_sync(0xffffffffu, v[j], j);
        if (!(d > 0.0) || !isfinite(d)) bad = true;
        const double r = rcp_newton(d);
        const double wr = v[j] * r;
#pragma unroll
        for (int k = j + 1; k < 16; ++k) {
          const double wk = __shfl_sync(0xffffffffu, v[j], k);
          v[k] = fma(-wr, wk, v[k]);
        }
      }
      // lane j: d_j = v[j];  column scale 1/sqrt(d_j)
      double dj = 0.0;
#pragma unroll
      for (int j = 0; j < 16; ++j) dj = (lane == j) ? v[j] : dj;
      const double rs = bad ? nan("") : 1.0 / sqrt(dj);
      if (lane < 16) rd[c0 + lane] = rs;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const double rj = __shfl_sync(0xffffffffu, rs, j);
        v[j] = (j <= lane) ? v[j] * rj : 0.0;
      }
      if (lane < 16) {
#pragma unroll
        for (int k = 0; k < 16; k += 2)
          *reinterpret_cast<double2*>(a + row * LD + c0 + k) = make_double2(v[k], v[k + 1]);
      }
      if (bad && lane == 0) *s_bad = 1;
    }
    VZ_POTF2_SYNC();
    VZ_TSTAMP(2 + 3 * mb);
    const int rem = 48 - c0;                 // rows below this macro block
    if (tid < rem) {
      // ---- panel: row r solves x * L11^T = a[r, c0:c0+16], right-looking ----
      const int r = c0 + 16 + tid;
      double xr[16];
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const double2 p = *reinterpret_cast<const double2*>(a + r * LD + c0 + k);
        xr[k] = p.x; xr[k + 1] = p.y;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        xr[j] *= rd[c0 + j];
#pragma unroll
        for (int k = j + 1; k < 16; ++k) xr[k] = fma(-xr[j], a[(c0 + k) * LD + c0 + j], xr[k]);
      }
#pragma unroll
      for (int k = 0; k < 16; k += 2)
        *reinterpret_cast<double2*>(a + r * LD + c0 + k) = make_double2(xr[k], xr[k + 1]);
    }
    VZ_POTF2_SYNC();
    VZ_TSTAMP(3 + 3 * mb);
    // ---- trailing update: a[i][k] -= sum_j a[i][c0+j] * a[k][c0+j],  c0+16 <= k <= i ----
    // thread (ti, tk) owns the entries (ti + 16 p, tk + 16 q) of every 16x16 block (p, q), q <= p
    {
      const int ti = tid & 15, tk = tid >> 4;
      const int nb = rem >> 4;
      for (int p = 0; p < nb; ++p) {
        const int i = c0 + 16 + 16 * p + ti;
        double ai[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const double2 q2 = *reinterpret_cast<const double2*>(a + i * LD + c0 + j);
          ai[j] = q2.x; ai[j + 1] = q2.y;
        }
        for (int q = 0; q <= p; ++q) {
          const int k = c0 + 16 + 16 * q + tk;
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const double2 q2 = *reinterpret_cast<const double2*>(a + k * LD + c0 + j);
            s0 = fma(ai[j], q2.x, s0);
            s1 = fma(ai[j + 1], q2.y, s1);
          }
          if (k <= i) a[i * LD + k] -= s0 + s1;
        }
      }
    }
    VZ_POTF2_SYNC();
    VZ_TSTAMP(4 + 3 * mb);
  }
  VZ_TSTAMP(14);
  // ---- inverse, step A: the four 16x16 diagonal blocks, one warp each, lane = column ----
  if (warp < 4 && lane < 16) {
    const int b0 = 16 * warp, c = lane;
    double xc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) xc[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      xc[i] *= rd[b0 + i];
#pragma unroll
      for (int k = i + 1; k < 16; ++k) xc[k] = fma(-a[(b0 + k) * LD + b0 + i], xc[i], xc[k]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[(b0 + i) * LD + b0 + c] = xc[i];
  }
  VZ_POTF2_SYNC();
  VZ_TSTAMP(15);
  // ---- step B: 16-level doubling, two pairs; thread = one (row, col) of each pair ----
  {
    const int rr = tid >> 4, cc = tid & 15;
    double acc[2];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const int p0 = 32 * pr;
      double s0 = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) s0 = fma(a[(p0 + 16 + rr) * LD + p0 + k], x[(p0 + k) * LD + p0 + cc], s0);
      acc[pr] = s0;
    }
    t[rr * 34 + cc] = acc[0];
    t[(16 + rr) * 34 + cc] = acc[1];
    VZ_POTF2_SYNC();
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const int p0 = 32 * pr;
      double s0 = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) s0 = fma(x[(p0 + 16 + rr) * LD + p0 + 16 + k], t[(16 * pr + k) * 34 + cc], s0);
      acc[pr] = s0;
    }
    x[(16 + rr) * LD + cc] = -acc[0];
    x[(48 + rr) * LD + 32 + cc] = -acc[1];
  }
  VZ_POTF2_SYNC();
  VZ_TSTAMP(16);
  // ---- step C: 32-level doubling; thread = a 2x2 patch of the 32x32 block ----
  {
    const int r0 = (tid >> 4) * 2, q0 = (tid & 15) * 2;
    double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const double a0 = a[(32 + r0) * LD + k], a1 = a[(33 + r0) * LD + k];
      const double2 xv = *reinterpret_cast<const double2*>(x + k * LD + q0);
      s00 = fma(a0, xv.x, s00); s01 = fma(a0, xv.y, s01);
      s10 = fma(a1, xv.x, s10); s11 = fma(a1, xv.y, s11);
    }
    *reinterpret_cast<double2*>(t + r0 * 34 + q0) = make_double2(s00, s01);
    *reinterpret_cast<double2*>(t + (r0 + 1) * 34 + q0) = make_double2(s10, s11);
    VZ_POTF2_SYNC();
    s00 = s01 = s10 = s11 = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const double b0v = x[(32 + r0) * LD + 32 + k], b1v = x[(33 + r0) * LD + 32 + k];
      const double2 tv = *reinterpret_cast<const double2*>(t + k * 34 + q0);
      s00 = fma(b0v, tv.x, s00); s01 = fma(b0v, tv.y, s01);
      s10 = fma(b1v, tv.x, s10); s11 = fma(b1v, tv.y, s11);
    }
    *reinterpret_cast<double2*>(x + (32 + r0) * LD + q0) = make_double2(-s00, -s01);
    *reinterpret_cast<double2*>(x + (33 + r0) * LD + q0) = make_double2(-s10, -s11);
  }
  VZ_POTF2_SYNC();
  VZ_TSTAMP(17);
}

}  // namespace vzgp
