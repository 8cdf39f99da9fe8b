// 64 x 64 Cholesky + inverse for the chain CTA of the dataflow factorisation (dataflow.cu), organised
// around the critical path with look-ahead.
//
// potf2_inv_64 (potf2.cuh) runs its four 16-column macro steps as  diagonal LDL^T -> barrier -> panel solve
// -> barrier -> trailing update -> barrier  with the whole CTA, then the inverse: 28 k cycles per block, of
// which only the 64-pivot chain (4 x 2.3 k) is inherently sequential.  Here warp 0 owns that chain and does
// ONLY what the next diagonal block needs (the solve of the next 16 rows and the update of the next 16 x 16
// block, from registers), while warps 1-7 do the rest of each step - the lower panel rows, the other
// trailing blocks, the 16 x 16 diagonal inverses and the block recurrence of the inverse - in its shadow:
//
//   warp 0, step m:   LDL^T of block (m,m)  ->  [DIAG]  ->  wait [UPD] of step m-1
//                     ->  L[m+1][m] (16 rows)  ->  [ROW]  ->  A[m+1][m+1] -= L[m+1][m] L[m+1][m]^T (registers)
//   helpers, step m:  [DIAG] -> rows >= 16(m+2) of the panel, I_m = L[m][m]^-1
//                     -> [H] -> blocks (p,q), q >= m+2;  X[m][q] = -I_m W_mq
//                     -> [ROW] -> blocks (p,m+1), p >= m+2 -> [H] -> W_{m+1,q} = sum_k L[m+1][k] X[k][q] -> [UPD]
// (X = L^-1 by the row recurrence  X[p][q] = -I_p sum_{k=q}^{p-1} L[p][k] X[k][q];  everything of row 3 but
// a final substitution with L[3][3] is done before the last diagonal block is factored.)
// The diagonal blocks are factored in block-LDL^T form with 2 x 2 pivots (one reciprocal per two columns on
// the chain), the next diagonal block is updated on the tensor pipe (DMMA) in place.
// [..] are named barriers (ids 4-7).  Critical path: 4 x LDL^T + 3 x (row solve + block update) + I_3 + one
// block product = 16 k cycles.  Same contract as potf2_inv_64: a [64][66] holds the lower triangle (16-blocks
// strictly above the diagonal ZERO), x ZERO on entry, t [32][34] scratch, rd [64]; 256 threads; ends with
// VZ_POTF2_SYNC().
#pragma once
#include "device.cuh"
#include "potf2.cuh"

namespace vzgp {

#ifdef VZ_LA_TIMING
__device__ long long g_la_t[64];
#define VZ_LAT(i) do { if (lane == 0) g_la_t[i] = clock64(); } while (0)
#else
#define VZ_LAT(i) do {} while (0)
#endif

constexpr int kLaDiag = 4, kLaRow = 5, kLaUpd = 6, kLaH = 7;   // named barrier ids

__device__ __forceinline__ void la_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void la_sync(int id, int n) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(n) : "memory"); }

// One warp: a[r0+i][k0+k] -= sum_j a[r0+i][c0+j] * a[k0+k][c0+j],  i, k in [0,16).  Lane -> row i = lane / 2,
// eight columns k = 8 (lane & 1) + 0..7.
__device__ __forceinline__ void la_blk_update(double* a, int r0, int k0, int c0, int lane) {
  constexpr int LD = 66;
  const int i = lane >> 1, kh = (lane & 1) * 8;
  double ai[16];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const double2 q = *reinterpret_cast<const double2*>(a + (r0 + i) * LD + c0 + j);
    ai[j] = q.x; ai[j + 1] = q.y;
  }
  double s[8];
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const double* bk = a + (k0 + kh + kk) * LD + c0;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const double2 q = *reinterpret_cast<const double2*>(bk + j);
      s0 = fma(ai[j], q.x, s0);
      s1 = fma(ai[j + 1], q.y, s1);
    }
    s[kk] = s0 + s1;
  }
  double* dst = a + (r0 + i) * LD + k0 + kh;
#pragma unroll
  for (int kk = 0; kk < 8; kk += 2) {
    double2 v = *reinterpret_cast<double2*>(dst + kk);
    v.x -= s[kk]; v.y -= s[kk + 1];
    *reinterpret_cast<double2*>(dst + kk) = v;
  }
}

// One thread: row r solves x * L_mm^T = a[r, c0:c0+16] in place (right-looking), rd = 1 / diag(L_mm).
__device__ __forceinline__ void la_row_solve(double* a, int r, int c0, const double* rd, double (&xr)[16]) {
  constexpr int LD = 66;
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

// 1/d from the MUFU seed (~20 bits) and ONE cubic step  r0 (1 + e + e^2),  e = 1 - d r0  (3 dependent FMAs;
// rcp_newton in potf2.cuh spends 4 on two quadratic steps).  Error after the step ~ e^3 < 2^-58.
__device__ __forceinline__ double rcp_cubic(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  const double e2 = fma(e, e, e);
  return fma(r, e2, r);
}

// 16 lanes (c = column index of the right-hand side): solve L x = rhs for the lower-triangular block
// a[b0.., b0..] (rd = 1 / diag), result into xc.
__device__ __forceinline__ void la_col_solve16(const double* a, int b0, const double* rd, double (&xc)[16]) {
  constexpr int LD = 66;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    xc[i] *= rd[b0 + i];
#pragma unroll
    for (int k = i + 1; k < 16; ++k) xc[k] = fma(-a[(b0 + k) * LD + b0 + i], xc[i], xc[k]);
  }
}

// 16 lanes (c = column): x[b0.., b0..] = inverse of the lower-triangular block a[b0.., b0..].
__device__ __forceinline__ void la_tri_inv16(const double* a, double* x, int b0, const double* rd, int c) {
  constexpr int LD = 66;
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

// One warp: dst (16 x 16) = alpha * A * B (+ dst if ACC), all row-major blocks in shared memory.
template <bool ACC>
__device__ __forceinline__ void la_blk_mul(double* dst, int ldd, const double* A, int lda, const double* B, int ldb,
                                           double alpha, int lane) {
  const int i = lane >> 1, ch = (lane & 1) * 8;
  double ai[16];
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const double2 q = *reinterpret_cast<const double2*>(A + i * lda + k);
    ai[k] = q.x; ai[k + 1] = q.y;
  }
  double s[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) s[c] = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const double* bk = B + k * ldb + ch;
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      const double2 q = *reinterpret_cast<const double2*>(bk + c);
      s[c] = fma(ai[k], q.x, s[c]);
      s[c + 1] = fma(ai[k], q.y, s[c + 1]);
    }
  }
  double* d = dst + i * ldd + ch;
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    double2 v = ACC ? *reinterpret_cast<double2*>(d + c) : make_double2(0.0, 0.0);
    v.x = fma(alpha, s[c], v.x); v.y = fma(alpha, s[c + 1], v.y);
    *reinterpret_cast<double2*>(d + c) = v;
  }
}

__device__ __forceinline__ void potf2_inv_64_la(double* a, double* x, double* t, double* rd, int* s_bad) {
  constexpr int LD = 66, TL = 34;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto tblk = [&](int b) { return t + (b >> 1) * 16 * TL + (b & 1) * 16; };   // four 16 x 16 blocks in t [32][34]
  if (warp == 0) {
    // ================= the pivot chain =================
    double v[16];
    {
      const int row = lane & 15;
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const double2 p = lane < 16 ? *reinterpret_cast<const double2*>(a + row * LD + k) : make_double2(0.0, 0.0);
        v[k] = p.x; v[k + 1] = p.y;
      }
    }
#pragma unroll 1
    for (int mb = 0; mb < 4; ++mb) {
      const int c0 = 16 * mb;
      const int row = c0 + (lane & 15);
      // ---- 16 x 16 LDL^T in registers: lane l < 16 owns row c0 + l (lanes 16..31 only shuffle) ----
      VZ_LAT(8 * mb + 0);
      __syncwarp();                            // converged before the shuffle chain (a split warp serialises it)
      // Block LDL^T with 2 x 2 pivots D_k = [[pa, pb], [pb, pc]] (SPD, so no pivoting is needed): ONE reciprocal
      // (of det D_k) per TWO columns on the pivot-to-pivot chain.  Row i keeps [v0, v1] = [l0, l1] D_k; the
      // Cholesky columns follow at the end from D_k = C_k C_k^T:  L[i][p0] = v0 / s0,  L[i][p1] =
      // (v1 - v0 pb / pa) / s1  with  s0 = sqrt(pa), s1 = sqrt(det / pa)  (also right for the rows of the block).
      bool bad = false;
      double my_a = 1.0, my_b = 0.0, my_rdet = 1.0;    // lane k < 8 keeps the scalars of block k
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int p0 = 2 * k, p1 = p0 + 1;
        const double pa = __shfl_sync(0xffffffffu, v[p0], p0);
        const double pb = __shfl_sync(0xffffffffu, v[p0], p1);
        const double pc = __shfl_sync(0xffffffffu, v[p1], p1);
        const double det = fma(pa, pc, -(pb * pb));
        if (!(pa > 0.0) || !(det > 0.0) || !isfinite(det)) bad = true;
        const double r = rcp_cubic(det);
        const double u0 = fma(v[p0], pc, -(v[p1] * pb)), u1 = fma(v[p1], pa, -(v[p0] * pb));
        const double l0 = u0 * r, l1 = u1 * r;
#pragma unroll
        for (int m = p1 + 1; m < 16; ++m) {
          const double w0 = __shfl_sync(0xffffffffu, v[p0], m);
          const double w1 = __shfl_sync(0xffffffffu, v[p1], m);
          v[m] = fma(-l1, w1, fma(-l0, w0, v[m]));
        }
        if (lane == k) { my_a = pa; my_b = pb; my_rdet = r; }
      }
      // lane k < 8: 1 / s0, pb / pa, 1 / s1 of block k
      const double inv_a = 1.0 / my_a;
      double rs0 = sqrt(inv_a), boa = my_b * inv_a, rs1 = sqrt(my_a * my_rdet);
      if (bad) { rs0 = nan(""); rs1 = nan(""); }
      if (lane < 8) { rd[c0 + 2 * lane] = rs0; rd[c0 + 2 * lane + 1] = rs1; }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double q0 = __shfl_sync(0xffffffffu, rs0, k), qb = __shfl_sync(0xffffffffu, boa, k);
        const double q1 = __shfl_sync(0xffffffffu, rs1, k);
        const double v0 = v[2 * k], v1 = v[2 * k + 1];
        v[2 * k] = (2 * k <= lane) ? v0 * q0 : 0.0;
        v[2 * k + 1] = (2 * k + 1 <= lane) ? fma(-v0, qb, v1) * q1 : 0.0;
      }
      if (lane < 16) {
#pragma unroll
        for (int k = 0; k < 16; k += 2)
          *reinterpret_cast<double2*>(a + row * LD + c0 + k) = make_double2(v[k], v[k + 1]);
      }
      if (bad && lane == 0) *s_bad = 1;
      __syncwarp();
      VZ_LAT(8 * mb + 1);
      // Helpers finished step mb-1 (its updates of block row mb+1 included).  Waiting here, BEFORE the arrival
      // below, also guarantees that every helper has left the previous phase of the DIAG / ROW barriers.
      if (mb > 0) la_sync(kLaUpd, 256);
      VZ_LAT(8 * mb + 2);
      la_arrive(kLaDiag, 256);                 // L[mb][mb], rd published
      if (mb == 3) break;
      // ---- the next 16 rows of the panel, then the next diagonal block, straight into registers ----
      const int r = c0 + 16 + (lane & 15);
      {
        double xr[16];
        if (lane < 16) la_row_solve(a, r, c0, rd, xr);
      }
      __syncwarp();
      VZ_LAT(8 * mb + 3);
      la_arrive(kLaRow, 256);                  // L[mb+1][mb] published
      {
        // A[mb+1][mb+1] -= L[mb+1][mb] L[mb+1][mb]^T on the tensor pipe, in place: 16 x 16 x 16 = three 8 x 8
        // fragments of the lower triangle x 4 k-steps.  Lane (fr, fk) feeds rows 8f + fr, k = 4 ks + fk.
        const int fr = lane >> 2, fk = lane & 3;
        const double* lr = a + (c0 + 16 + fr) * LD + c0 + fk;
        double d00[2] = {0.0, 0.0}, d10[2] = {0.0, 0.0}, d11[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double f0 = lr[4 * ks], f1 = lr[8 * LD + 4 * ks];
          dmma_8x8x4(d00[0], d00[1], f0, f0);
          dmma_8x8x4(d10[0], d10[1], f1, f0);
          dmma_8x8x4(d11[0], d11[1], f1, f1);
        }
        double* blk = a + (c0 + 16 + fr) * LD + c0 + 16 + 2 * fk;
        double2 t0 = *reinterpret_cast<double2*>(blk), t1 = *reinterpret_cast<double2*>(blk + 8 * LD);
        double2 t2 = *reinterpret_cast<double2*>(blk + 8 * LD + 8);
        t0.x -= d00[0]; t0.y -= d00[1]; t1.x -= d10[0]; t1.y -= d10[1]; t2.x -= d11[0]; t2.y -= d11[1];
        *reinterpret_cast<double2*>(blk) = t0;
        *reinterpret_cast<double2*>(blk + 8 * LD) = t1;
        *reinterpret_cast<double2*>(blk + 8 * LD + 8) = t2;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 16; k += 2) {          // next block, one row per lane
        const double2 p = lane < 16 ? *reinterpret_cast<const double2*>(a + r * LD + c0 + 16 + k) : make_double2(0.0, 0.0);
        v[k] = p.x; v[k + 1] = p.y;
      }
    }
  } else {
    // ================= helpers: warps 1..7 =================
    const int hw = warp - 1;
#pragma unroll 1
    for (int mb = 0; mb < 4; ++mb) {
      const int c0 = 16 * mb;
      la_sync(kLaDiag, 256);
      if (hw == 0) {                            // panel rows below block mb+1
        const int r = c0 + 32 + lane;
        double xr[16];
        if (r < 64) la_row_solve(a, r, c0, rd, xr);
      }
      if (mb == 3) {
        // last block row of X: every one of its 64 columns is one forward substitution with L[3][3]
        // (right-hand side -W_3q, or the identity for the diagonal block) - no I_3, no block product
        if (hw >= 3 && lane < 16) {
          const int q = hw - 3;
          double xc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) xc[i] = q < 3 ? -tblk(q)[i * TL + lane] : (i == lane ? 1.0 : 0.0);
          la_col_solve16(a, 48, rd, xc);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[(48 + i) * LD + 16 * q + lane] = xc[i];
        }
        break;
      }
      if (hw == 6 && lane < 16) la_tri_inv16(a, x, c0, rd, lane);   // I_mb
      if (hw == 6) VZ_LAT(42 + mb);
      if (hw == 0) VZ_LAT(46 + mb);
      la_sync(kLaH, 224);
      if (hw == 3) VZ_LAT(50 + mb);
      // trailing blocks that do not involve block row mb+1
      if (mb == 0) {
        if (hw == 0) la_blk_update(a, 32, 32, 0, lane);
        if (hw == 1) la_blk_update(a, 48, 32, 0, lane);
        if (hw == 2) la_blk_update(a, 48, 48, 0, lane);
      } else if (mb == 1) {
        if (hw == 0) la_blk_update(a, 48, 48, 16, lane);
      }
      // X[mb][q] = -I_mb W_mb,q  (W computed at the end of the previous step)
      if (mb >= 1 && hw >= 3 && hw - 3 < mb)
        la_blk_mul<false>(x + c0 * LD + 16 * (hw - 3), LD, x + c0 * LD + c0, LD, tblk(hw - 3), TL, -1.0, lane);
      la_sync(kLaRow, 256);                     // L[mb+1][mb] is there
      if (mb == 0) {
        if (hw == 0) la_blk_update(a, 32, 16, 0, lane);
        if (hw == 1) la_blk_update(a, 48, 16, 0, lane);
      } else if (mb == 1) {
        if (hw == 0) la_blk_update(a, 48, 32, 16, lane);
      }
      if (mb >= 1) la_sync(kLaH, 224);          // X[mb][*] complete
      // W_{mb+1,q} = sum_{k=q}^{mb} L[mb+1][k] X[k][q],  q = 0..mb  (one warp per q)
      if (hw >= 3 && hw - 3 <= mb) {
        const int q = hw - 3, p = mb + 1;
        double* w = tblk(q);
        la_blk_mul<false>(w, TL, a + 16 * p * LD + 16 * q, LD, x + 16 * q * LD + 16 * q, LD, 1.0, lane);
        for (int k = q + 1; k <= mb; ++k)
          la_blk_mul<true>(w, TL, a + 16 * p * LD + 16 * k, LD, x + 16 * k * LD + 16 * q, LD, 1.0, lane);
      }
      if (hw == 3) VZ_LAT(54 + mb);
      la_arrive(kLaUpd, 256);
    }
  }
  VZ_LAT(40);
  VZ_POTF2_SYNC();
  VZ_LAT(41);
}

}  // namespace vzgp
