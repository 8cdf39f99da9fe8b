// Blocked Cholesky, L^-1, L^-T and K_y^-1 = L^-T L^-1 as ONE dataflow kernel.
//
// Replaces (reference, via TFP / XLA:CPU LAPACK potrf + triangular_solve): retrying_cholesky at
// vizier/_src/jax/models/tuned_gp_models.py:272-280, the factor inside GaussianProcess.log_prob
// (vizier/_src/jax/stochastic_process_model.py:940-966) and precompute_predictive (:968-997).
//
// Round 1 ran the factorisation as 3 launches per 64-column panel (k_potf2_inv / k_trsm_panel /
// k_syrk_potf2), then 2 log2(nb) launches of recursive doubling for L^-1 and one k_lauum, all on DFMA
// register tiles: 0.5 ms of launch-separated latency chain at N = 1000.  Here every 64 x 64 tile of the
// result is a TASK owned by one CTA; tasks hand tiles over through release/acquire flags in global memory,
// operands are staged by a producer warp with cp.async (64 x 16 boxes in the 128-byte-swizzle layout of the
// scoring kernels, 4-stage full/empty mbarrier ring) and multiplied on the FP64 tensor pipe (DMMA.8x8x4),
// so a tile GEMM starts the moment its inputs exist.  (The first version staged the boxes with TMA.  Tiles
// written a moment earlier by OTHER CTAs' generic-proxy stores were then occasionally read stale by the
// async proxy - wrong tiles in 3 of 3 runs at np = 2048, still 1 of 8 with release/acquire flags plus
// fence.proxy.async in every writing thread and in the reader - so cross-CTA hand-over stays in the generic
// proxy: cp.async.cg reads L2 like any other load ordered by the acquire.)
//
//   chain CTA (ticket 0), for j = 0 .. nb-1            the only sequential part (nb = np / 64)
//       L[j,j-1] = (A[j,j-1] - S(j,0)) * Linv_{j-1}^T   (Linv_{j-1} is still in shared memory)
//       D        =  A[j,j]   - S(j,1) - L[j,j-1] L[j,j-1]^T
//       L_jj, Linv_jj = potf2_inv_64(D)                 (potf2.cuh)
//   PART(j,w)   S(j,0) = sum_{k<=j-2} L[j,k] L[j-1,k]^T,  S(j,1) = sum_{k<=j-2} L[j,k] L[j,k]^T
//               (everything of the chain's update that does not depend on the previous panel)
//   TILE(i,j)   i >= j+2:  L[i,j] = (A[i,j] - sum_{k<j} L[i,k] L[j,k]^T) * Linv_jj^T      (left-looking)
//   LINV(i,j)   i > j:     Y[j,i] = -(sum_{k=j}^{i-1} Y[j,k] L[i,k]^T) * Linv_ii^T,  Y = L^-T; also stored
//               transposed as X[i,j] (X = L^-1).  Working with Y makes every product an "NT" GEMM whose
//               operands are both row-major [row][k] boxes.
//   KINV(i,j)   i >= j:    Kinv[i,j] = sum_{k>=i} Y[i,k] Y[j,k]^T
//
// Scheduling.  Tasks are sorted by the panel step at which their last input appears; every CTA takes a
// ticket (atomic counter) and runs the task of that rank.  A task only waits for tasks of lower rank and
// for chain steps that themselves only wait for lower ranks, and tickets are handed out in dispatch
// order, so the lowest unfinished task is always resident and can finish: no cooperative launch, no
// co-residency assumption, safe next to other kernels (concurrent ARD restarts on other streams).
// CTAs that arrive early accumulate the inputs that already exist and then follow the chain; the
// critical path is  nb x (flag hop + 2 tile GEMMs + potf2_inv_64).  Waits are bounded by a timeout that
// raises ctrl[1] instead of hanging the GPU.
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "async.cuh"
#include "device.cuh"
#include "launchers.h"

#ifdef VZ_DF_TIMING
namespace vzgp { __device__ long long g_df_t[64 * 8]; }
#define VZ_DFT(j, s) do { if (threadIdx.x == 0 && (j) < 64) g_df_t[(j) * 8 + (s)] = clock64(); } while (0)
#else
#define VZ_DFT(j, s) do {} while (0)
#endif
#define VZ_POTF2_SYNC() asm volatile("bar.sync 0, 256;\n" ::: "memory")   // the chain's 8 math warps
#include "potf2.cuh"
#include "potf2_la.cuh"

namespace vzgp {

constexpr int kDfThreads = 320;   // 8 math warps + producer / publisher warp + (chain only) prefetch warp
constexpr int kDfMath = 256;
constexpr int kDfStages = 4;
constexpr int kBoxD = 64 * 16;                 // one TMA box: 64 rows x 16 doubles (8 KB)
constexpr int kStageD = 2 * kBoxD;             // A box | B box
constexpr int kRingD = kDfStages * kStageD;    // 64 KB
constexpr int kTbufD = 4 * kBoxD;              // a whole 64 x 64 operand tile in box layout (32 KB)
constexpr int kChainLD = 66;                   // potf2_inv_64's row stride
constexpr int kChainD = 3 * 64 * kChainLD + 32 * 34;   // chain CTA: A_ | X_ | T_ | staged B' tile
constexpr size_t kDfWorkerBytes = sizeof(double) * (kRingD + kTbufD), kDfChainBytes = sizeof(double) * kChainD;
constexpr size_t kDfSmemBytes = 1024 + (kDfWorkerBytes > kDfChainBytes ? kDfWorkerBytes : kDfChainBytes) + 256;

enum { DF_PART = 0, DF_TILE = 1, DF_LINV = 2, DF_KINV = 3 };

struct DfArgs {
  double* L;        // [np x np] row-major: the shifted matrix on entry (lower tiles), the factor on exit
  double* X;
  double* Y;
  double* S;        // [nb][2][64*64] partial sums for the chain
  double* Kinv;     // lower tiles of K_y^-1, or nullptr
  int* ctrl;        // [0] tickets of the normal queue, [1] timeout marker, [2] tickets of the critical queue, [3] chain election
  int* flagL;       // [nb*nb] tile (i,j) of L final ((j,j): also Linv_jj in X and Y)
  int* flagY;       // [nb*nb] tile (j,i), j < i, of Y (and X[i,j]) final
  int* flagS;       // [nb*2]
  const int4* tasks;   // normal queue, in schedule order: (type, i, j, key)
  int ntasks;
  const int4* crit;    // critical queue (the PART tasks the chain waits for): (type, i, j, need)
  int ncrit;
  int np, nb;
  int* bad;         // raised on a non-positive / non-finite pivot
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long df_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Spin until *p != 0.  A producer that never shows up would be a bug (or a device fault upstream): give up
// after 4 s, mark the run (the host reports an error) and let every later wait fall through.
__device__ __forceinline__ void df_wait(const int* p, int* ctrl) {
  if (ld_acquire_gpu(p)) return;
  const unsigned long long t0 = df_timer_ns();
  while (!ld_acquire_gpu(p)) {
    __nanosleep(32);
    if (*reinterpret_cast<volatile int*>(ctrl + 1)) return;
    if (df_timer_ns() - t0 > 4000000000ull) { atomicExch(ctrl + 1, 1); return; }
  }
}

// acc[f][g][0..1]: rows wm*16 + f*8 + fr, columns wn*32 + g*8 + 2*fk + {0,1}   (DMMA D fragment)
struct DfFrag {
  double v[2][4][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) v[f][g][0] = v[f][g][1] = 0.0;
  }
};

// One 16-wide k slab from swizzled boxes: Abox/Bbox = [64 rows][16 doubles], 16-byte chunk c of row r at
// chunk c ^ (r & 7).  Lane (fr, fk) reads k = 8h + 2fk, 8h + 2fk + 1: the operands of two DMMA k-steps.
__device__ __forceinline__ void df_slab(DfFrag& acc, const double* Abox, const double* Bbox, int wm, int wn, int fr, int fk) {
  const double* Arow0 = Abox + (wm * 16 + fr) * 16;
  const double* Brow0 = Bbox + (wn * 32 + fr) * 16;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int co = (((h * 4 + fk) ^ fr) * 2);
    const double2 a0 = *reinterpret_cast<const double2*>(Arow0 + co);
    const double2 a1 = *reinterpret_cast<const double2*>(Arow0 + 8 * 16 + co);
    double2 b[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) b[g] = *reinterpret_cast<const double2*>(Brow0 + g * 8 * 16 + co);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      dmma_8x8x4(acc.v[0][g][0], acc.v[0][g][1], a0.x, b[g].x);
      dmma_8x8x4(acc.v[1][g][0], acc.v[1][g][1], a1.x, b[g].x);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      dmma_8x8x4(acc.v[0][g][0], acc.v[0][g][1], a0.y, b[g].y);
      dmma_8x8x4(acc.v[1][g][0], acc.v[1][g][1], a1.y, b[g].y);
    }
  }
}

// Producer warp: one 64 x 16 box (rows row0.., columns col0..col0+15 of a row-major [np x np] matrix) into
// the swizzled box layout: 16-byte chunk c of row r at chunk c ^ (r & 7).  512 chunks, 16 per lane.
__device__ __forceinline__ void df_copy_box(double* box, const double* __restrict__ G, int np, int row0, int col0, int lane) {
  const int c = lane & 7, r0 = lane >> 3;
  const double* src = G + (size_t)(row0 + r0) * np + col0 + c * 2;
  const unsigned dst0 = smem_u32(box);
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int r = r0 + 4 * t;
    const unsigned dst = dst0 + (unsigned)(r * 16 + ((c ^ (r & 7)) * 2)) * 8u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + (size_t)4 * t * np) : "memory");
  }
}
// All of this lane's cp.async so far -> one arrival on `bar` when they have landed.
__device__ __forceinline__ void df_copy_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// C = As * Bs^T over k = 0..63, both operands in padded shared memory [64][kChainLD] (chain CTA).
__device__ __forceinline__ void df_gemm_padded(DfFrag& acc, const double* As, const double* Bs, int wm, int wn, int fr, int fk) {
  constexpr int LD = kChainLD;
  const double* Ar = As + (wm * 16 + fr) * LD + 2 * fk;
  const double* Br = Bs + (wn * 32 + fr) * LD + 2 * fk;
#pragma unroll 2
  for (int kk = 0; kk < 64; kk += 8) {
    const double2 a0 = *reinterpret_cast<const double2*>(Ar + kk);
    const double2 a1 = *reinterpret_cast<const double2*>(Ar + 8 * LD + kk);
    double2 b[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) b[g] = *reinterpret_cast<const double2*>(Br + g * 8 * LD + kk);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      dmma_8x8x4(acc.v[0][g][0], acc.v[0][g][1], a0.x, b[g].x);
      dmma_8x8x4(acc.v[1][g][0], acc.v[1][g][1], a1.x, b[g].x);
      dmma_8x8x4(acc.v[0][g][0], acc.v[0][g][1], a0.y, b[g].y);
      dmma_8x8x4(acc.v[1][g][0], acc.v[1][g][1], a1.y, b[g].y);
    }
  }
}

// The same with Bs lower triangular (Bs[n][k] = 0 for k > n): the 8-wide k group q only reaches the column
// fragments whose last column is >= 8q, i.e. fragment g of warp column wn needs q <= 4 wn + g.  Both warps
// of a scheduler partition (same wm, wn = 0 / 1) together issue 36 of the 64 group-fragments.
__device__ __forceinline__ void df_gemm_padded_tri(DfFrag& acc, const double* As, const double* Bs, int wm, int wn, int fr, int fk) {
  constexpr int LD = kChainLD;
  const double* Ar = As + (wm * 16 + fr) * LD + 2 * fk;
  const double* Br = Bs + (wn * 32 + fr) * LD + 2 * fk;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (q > 4 * wn + 3) break;     // warp-uniform
    const double2 a0 = *reinterpret_cast<const double2*>(Ar + 8 * q);
    const double2 a1 = *reinterpret_cast<const double2*>(Ar + 8 * LD + 8 * q);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (q > 4 * wn + g) continue;
      const double2 b = *reinterpret_cast<const double2*>(Br + g * 8 * LD + 8 * q);
      dmma_8x8x4(acc.v[0][g][0], acc.v[0][g][1], a0.x, b.x);
      dmma_8x8x4(acc.v[1][g][0], acc.v[1][g][1], a1.x, b.x);
      dmma_8x8x4(acc.v[0][g][0], acc.v[0][g][1], a0.y, b.y);
      dmma_8x8x4(acc.v[1][g][0], acc.v[1][g][1], a1.y, b.y);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Chain CTA: the sequential part.  Warps 0-7 do the arithmetic (barrier 0 with 256 threads), warp 8 publishes
// (fence + release of the tiles the math warps stored, so that no math thread waits for a fence), warp 9
// prefetches the next panel's input tile while potf2_inv_64 runs.
//   PART tasks deliver  B'(j) = A[j,j-1] - sum_{k<=j-2} L[j,k] L[j-1,k]^T   and
//                       D'(j) = A[j,j]   - sum_{k<=j-2} L[j,k] L[j,k]^T      (flagS[j][0/1]),
//   step j:  L[j,j-1] = B'(j) Linv_{j-1}^T ;  D = D'(j) - L[j,j-1] L[j,j-1]^T ;  L_jj, Linv_jj = potf2_inv_64(D).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }

__device__ void df_chain(const DfArgs& a, double* sm, uint64_t* bars) {
  constexpr int LD = kChainLD;
  double* A_ = sm;                  // [64][66]
  double* X_ = sm + 64 * LD;        // [64][66]
  double* T_ = X_ + 64 * LD;        // [32][34]
  double* B_ = T_ + 32 * 34;        // [64][66] staged B'(j)
  __shared__ double rd[64];
  __shared__ int s_bad;
  uint64_t* b_full = bars;          // warp 9's cp.async of B'(j) landed (32 lane arrivals)
  uint64_t* b_empty = bars + 1;     // the 8 math warps are done reading B_ (T-GEMM of step j)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = a.np, nb = a.nb;
  auto math_sync = [&]() { named_sync(0, 256); };
  if (warp == 8) {
    // ---------------- publisher ----------------
    for (int j = 0; j < nb; ++j) {
      if (j > 0) {
        named_sync(2, 288);          // L[j,j-1] stored by the math warps
        if (lane == 0) { __threadfence(); st_release_gpu(a.flagL + j * nb + (j - 1), 1); }
      }
      named_sync(3, 288);            // potf2 done: L_jj in A_, Linv_jj in X_
      // The three diagonal tiles go out from here, so the math warps start the next panel at once.
      double* Ljj = a.L + (size_t)j * 64 * np + j * 64;
      double* Xjj = a.X + (size_t)j * 64 * np + j * 64;
      double* Yjj = a.Y + (size_t)j * 64 * np + j * 64;
      const int j2 = 2 * lane;
#pragma unroll 4
      for (int i = 0; i < 64; ++i) {
        *reinterpret_cast<double2*>(Ljj + (size_t)i * np + j2) = *reinterpret_cast<const double2*>(A_ + i * LD + j2);
        *reinterpret_cast<double2*>(Xjj + (size_t)i * np + j2) = *reinterpret_cast<const double2*>(X_ + i * LD + j2);
        *reinterpret_cast<double2*>(Yjj + (size_t)i * np + j2) = make_double2(X_[j2 * LD + i], X_[(j2 + 1) * LD + i]);   // transposed
      }
      if (lane == 0 && s_bad) *a.bad = 1;
      __syncwarp();
      if (j + 1 < nb) named_arrive(8, 288);   // A_ / X_ may be overwritten (the math warps wait for this before they do)
      if (lane == 0) { __threadfence(); st_release_gpu(a.flagL + j * nb + j, 1); }
    }
    return;
  }
  if (warp == 9) {
    // ---------------- prefetcher ----------------
    for (int j = 1; j < nb; ++j) {
      mbar_wait(b_empty, ((j - 1) & 1) ^ 1);         // B_ free (first use: passes at once)
      if (lane == 0) { df_wait(a.flagS + j * 2, a.ctrl); df_wait(a.flagS + j * 2 + 1, a.ctrl); }
      __syncwarp();
      const double* Bp = a.S + (size_t)(j * 2) * 4096;     // [64][64] row-major
      const unsigned dst0 = smem_u32(B_);
#pragma unroll 4
      for (int t = 0; t < 64; ++t) {                        // 2048 16-byte chunks, 64 per lane; chunk q: row q / 32
        const int q = lane + 32 * t, r = q >> 5, c2 = (q & 31) * 2;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst0 + (unsigned)(r * LD + c2) * 8u), "l"(Bp + r * 64 + c2) : "memory");
      }
      df_copy_arrive(b_full);
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    return;
  }
  // ---------------- math warps ----------------
  const int wm = warp & 3, wn = warp >> 2, fr = lane >> 2, fk = lane & 3;
  for (int j = 0; j < nb; ++j) {
    const double* Ljj = a.L + (size_t)j * 64 * np + j * 64;
    VZ_DFT(j, 0);
    if (j == 0) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = tid + 256 * u, i = e >> 5, j2 = (e & 31) * 2;
        const double2 v = *reinterpret_cast<const double2*>(Ljj + (size_t)i * np + j2);
        const bool upper_blk = (j2 >> 4) > (i >> 4);
        *reinterpret_cast<double2*>(A_ + i * LD + j2) = upper_blk ? make_double2(0.0, 0.0) : v;
        *reinterpret_cast<double2*>(X_ + i * LD + j2) = make_double2(0.0, 0.0);
      }
    } else {
      // ---- L[j,j-1] = B'(j) * Linv_{j-1}^T ; X_ still holds Linv_{j-1}, B_ the prefetched B'(j) ----
      mbar_wait(b_full, (j - 1) & 1);
      VZ_DFT(j, 1);
      double* Lsub = a.L + (size_t)j * 64 * np + (j - 1) * 64;
      const double* Dp = a.S + (size_t)(j * 2 + 1) * 4096;     // D'(j): final since flagS[j][1] (seen by warp 9)
      DfFrag acc;
      acc.zero();
      df_gemm_padded_tri(acc, B_, X_, wm, wn, fr, fk);          // Linv_{j-1} is lower triangular
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b_empty)) : "memory");
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
          *reinterpret_cast<double2*>(Lsub + (size_t)r * np + c) = make_double2(acc.v[f][g][0], acc.v[f][g][1]);
        }
      math_sync();                // every warp is done reading X_
      named_sync(8, 288);         // ... and so is the publisher (the diagonal tiles of step j-1 are out)
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
          *reinterpret_cast<double2*>(X_ + r * LD + c) = make_double2(acc.v[f][g][0], acc.v[f][g][1]);
        }
      named_arrive(2, 288);       // publisher: fence + release of L[j,j-1] (nobody here waits for it)
      // D'(j) in fragment layout; in flight during the syrk product
      DfFrag dp;
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
          const double2 v = __ldcg(reinterpret_cast<const double2*>(Dp + r * 64 + c));
          dp.v[f][g][0] = v.x; dp.v[f][g][1] = v.y;
        }
      math_sync();                // X_ (= L[j,j-1]) complete
      VZ_DFT(j, 2);
      // ---- D = D'(j) - L[j,j-1] L[j,j-1]^T ----
      acc.zero();
      df_gemm_padded(acc, X_, X_, wm, wn, fr, fk);
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
          const bool upper_blk = (c >> 4) > (r >> 4);
          *reinterpret_cast<double2*>(A_ + r * LD + c) =
              upper_blk ? make_double2(0.0, 0.0) : make_double2(dp.v[f][g][0] - acc.v[f][g][0], dp.v[f][g][1] - acc.v[f][g][1]);
        }
      math_sync();                // syrk reads of X_ finished
      for (int e = tid; e < 64 * LD; e += 256) X_[e] = 0.0;
    }
    if (tid == 0) s_bad = 0;
    math_sync();
    VZ_DFT(j, 3);
#ifdef VZ_DF_NO_LOOKAHEAD
    potf2_inv_64(A_, X_, T_, rd, &s_bad);
#else
    potf2_inv_64_la(A_, X_, T_, rd, &s_bad);
#endif
    VZ_DFT(j, 4);
    named_arrive(3, 288);         // publisher: stores L_jj, Linv_jj, Linv_jj^T from A_ / X_, fence, release
    VZ_DFT(j, 5);
  }
}

// ---------------------------------------------------------------------------------------------------
// Worker CTA: one tile task.
// ---------------------------------------------------------------------------------------------------
struct DfPlan {      // how a task's accumulation walks the k tiles (same on producer and consumers)
  int k0, k1;        // k tiles [k0, k1)
  int rowA, rowB;    // block rows of the A / B operand
  int finalB;        // block index of the diagonal inverse used by the final product, or -1
};
__device__ __forceinline__ DfPlan df_plan(int type, int i, int j, int nb) {
  DfPlan p;
  p.finalB = -1;
  if (type == DF_TILE) { p.k0 = 0; p.k1 = j; p.rowA = i; p.rowB = j; p.finalB = j; }
  else if (type == DF_PART) { p.k0 = 0; p.k1 = i - 1; p.rowA = i; p.rowB = j; }         // j = i-1 or i
  else if (type == DF_LINV) { p.k0 = j; p.k1 = i; p.rowA = j; p.rowB = i; p.finalB = i; }
  else { p.k0 = i; p.k1 = nb; p.rowA = i; p.rowB = j; }
  return p;
}

__device__ void df_worker(const DfArgs& a, const int4 task, double* sm, uint64_t* full_bar, uint64_t* empty_bar,
                          unsigned& slab) {
  const int type = task.x, ti = task.y, tj = task.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = a.np, nb = a.nb;
  double* ring = sm;
  double* Tbuf = sm + kRingD;
  const DfPlan p = df_plan(type, ti, tj, nb);
  if (warp > kDfMath / 32) return;     // the chain CTA's prefetch warp has no job in a worker
  if (warp == kDfMath / 32) {
    // ---------------- producer warp (cp.async) ----------------
    const double* GA = (type == DF_LINV || type == DF_KINV) ? a.Y : a.L;
    const double* GB = (type == DF_KINV) ? a.Y : a.L;
    for (int k = p.k0; k < p.k1; ++k) {
      if (lane == 0) {       // inputs of this k tile
        if (type == DF_TILE || type == DF_PART) {
          df_wait(a.flagL + p.rowA * nb + k, a.ctrl);
          df_wait(a.flagL + p.rowB * nb + k, a.ctrl);
        } else if (type == DF_LINV) {
          df_wait(k == tj ? a.flagL + tj * nb + tj : a.flagY + tj * nb + k, a.ctrl);
          df_wait(a.flagL + ti * nb + k, a.ctrl);
        } else {
          df_wait(k == ti ? a.flagL + ti * nb + ti : a.flagY + ti * nb + k, a.ctrl);
          df_wait(k == tj ? a.flagL + tj * nb + tj : a.flagY + tj * nb + k, a.ctrl);
        }
      }
      __syncwarp();          // lane 0's acquire, then the whole warp's loads
      for (int ks = 0; ks < 4; ++ks, ++slab) {
        const int stage = slab % kDfStages;
        mbar_wait(empty_bar + stage, ((slab / kDfStages) & 1) ^ 1);
        double* base = ring + stage * kStageD;
        df_copy_box(base, GA, np, p.rowA * 64, k * 64 + ks * 16, lane);
        df_copy_box(base + kBoxD, GB, np, p.rowB * 64, k * 64 + ks * 16, lane);
        df_copy_arrive(full_bar + stage);
      }
    }
    if (p.finalB >= 0) {
      if (lane == 0) df_wait(a.flagL + p.finalB * nb + p.finalB, a.ctrl);
      __syncwarp();
      for (int ks = 0; ks < 4; ++ks, ++slab) {
        const int stage = slab % kDfStages;
        mbar_wait(empty_bar + stage, ((slab / kDfStages) & 1) ^ 1);
        double* base = ring + stage * kStageD;
        df_copy_box(base + kBoxD, a.X, np, p.finalB * 64, p.finalB * 64 + ks * 16, lane);
        df_copy_arrive(full_bar + stage);
      }
    }
    return;
  }
  // ---------------- math warps ----------------
  const int wm = warp & 3, wn = warp >> 2, fr = lane >> 2, fk = lane & 3;
  auto math_sync = [&]() { asm volatile("bar.sync 1, %0;\n" ::"n"(kDfMath) : "memory"); };
  DfFrag acc;
  acc.zero();
  for (int k = p.k0; k < p.k1; ++k) {
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks, ++slab) {
      const int stage = slab % kDfStages;
      mbar_wait(full_bar + stage, (slab / kDfStages) & 1);
      const double* stg = ring + stage * kStageD;
      df_slab(acc, stg, stg + kBoxD, wm, wn, fr, fk);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(empty_bar + stage)) : "memory");
    }
  }
  if (p.finalB >= 0) {
    // T = (A - acc) or (-acc), as the A operand of the final product, in box layout.  TILE reads its own
    // tile of the matrix being factored here (nobody else writes it; it is overwritten in place below).
    const double* At = a.L + (size_t)ti * 64 * np + tj * 64;
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
        double2 v;
        if (type == DF_TILE) {
          const double2 o = *reinterpret_cast<const double2*>(At + (size_t)r * np + c);
          v = make_double2(o.x - acc.v[f][g][0], o.y - acc.v[f][g][1]);
        } else {
          v = make_double2(-acc.v[f][g][0], -acc.v[f][g][1]);
        }
        const int box = c >> 4, chunk = ((c & 15) >> 1) ^ (r & 7);
        *reinterpret_cast<double2*>(Tbuf + box * kBoxD + r * 16 + chunk * 2) = v;
      }
    math_sync();
    acc.zero();
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks, ++slab) {
      const int stage = slab % kDfStages;
      mbar_wait(full_bar + stage, (slab / kDfStages) & 1);
      const double* stg = ring + stage * kStageD;
      df_slab(acc, Tbuf + ks * kBoxD, stg + kBoxD, wm, wn, fr, fk);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(empty_bar + stage)) : "memory");
    }
  }
  // ---------------- epilogue ----------------
  double* out;
  int ldo = np;
  int* flag = nullptr;
  if (type == DF_TILE) { out = a.L + (size_t)ti * 64 * np + tj * 64; flag = a.flagL + ti * nb + tj; }
  else if (type == DF_PART) { out = a.S + (size_t)(ti * 2 + (tj == ti ? 1 : 0)) * 4096; ldo = 64; flag = a.flagS + ti * 2 + (tj == ti ? 1 : 0); }
  else if (type == DF_LINV) { out = a.Y + (size_t)tj * 64 * np + ti * 64; flag = a.flagY + tj * nb + ti; }
  else { out = a.Kinv + (size_t)ti * 64 * np + tj * 64; }
  if (type == DF_PART) {
    // B'(j) / D'(j) = (tile of the matrix being factored) - acc: the chain then loads one tile, not two
    const double* At = a.L + (size_t)ti * 64 * np + tj * 64;
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
        const double2 o = *reinterpret_cast<const double2*>(At + (size_t)r * np + c);
        acc.v[f][g][0] = o.x - acc.v[f][g][0];
        acc.v[f][g][1] = o.y - acc.v[f][g][1];
      }
  }
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
      *reinterpret_cast<double2*>(out + (size_t)r * ldo + c) = make_double2(acc.v[f][g][0], acc.v[f][g][1]);
    }
  if (type == DF_LINV) {   // X[i,j] = Y[j,i]^T
    double* xt = a.X + (size_t)ti * 64 * np + tj * 64;
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int r = wm * 16 + f * 8 + fr, c = wn * 32 + g * 8 + 2 * fk;
        xt[(size_t)c * np + r] = acc.v[f][g][0];
        xt[(size_t)(c + 1) * np + r] = acc.v[f][g][1];
      }
  }
  // (PART epilogue handled above.)  All warps are done with this task's tiles (Tbuf may be rewritten by the next task).  The barrier also
  // orders every thread's stores before thread 0's fence: the release below is cumulative.
  math_sync();
  if (flag != nullptr && tid == 0) {
    __threadfence();
    st_release_gpu(flag, 1);
  }
}

// Next task for a worker CTA (one thread calls it).  Two queues: the PART tasks feed the chain directly and
// jump the line - but a critical task is handed out only once every normal task it depends on has been
// handed out (crit.w = how many normal tickets that takes), so whoever holds a task only ever waits for
// tasks that are already held by running CTAs (or for the chain, which only waits for critical tasks whose
// last dependency is held by a CTA that does not wait for that chain step): no deadlock whatever the number
// of resident CTAs.  Returns type -1 when both queues are exhausted.
__device__ int4 df_next_task(const DfArgs& a) {
  for (;;) {
    const int c = *reinterpret_cast<volatile int*>(a.ctrl + 2);
    if (c < a.ncrit) {
      const int4 ct = a.crit[c];
      if (*reinterpret_cast<volatile int*>(a.ctrl) >= ct.w) {
        if (atomicCAS(a.ctrl + 2, c, c + 1) == c) return ct;
        continue;
      }
    }
    if (*reinterpret_cast<volatile int*>(a.ctrl) < a.ntasks) {
      const int t = atomicAdd(a.ctrl, 1);
      if (t < a.ntasks) return a.tasks[t];
    } else if (c >= a.ncrit) {
      return make_int4(-1, 0, 0, 0);
    }
    // normal queue exhausted, critical tasks left: their needs are met now, take them on the next turn
  }
}

__global__ void __launch_bounds__(kDfThreads, 2) k_chol_dataflow(const DfArgs a) {
  extern __shared__ double smem_raw[];
  double* sm = smem_raw + (((1024u - (static_cast<unsigned>(__cvta_generic_to_shared(smem_raw)) & 1023u)) & 1023u) >> 3);
  __shared__ int4 s_task;
  __shared__ int s_chain;
  __shared__ uint64_t bars[2 * kDfStages];
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_chain = atomicAdd(a.ctrl + 3, 1) == 0 ? 1 : 0;     // the first CTA to arrive runs the chain
    for (int s = 0; s < kDfStages; ++s) { mbar_init(bars + s, 32); mbar_init(bars + kDfStages + s, kDfMath / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (s_chain) {
    __shared__ uint64_t chain_bars[2];
    if (tid == 0) {
      mbar_init(chain_bars, 32);
      mbar_init(chain_bars + 1, kDfMath / 32);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    df_chain(a, sm, chain_bars);
    return;
  }
  // Persistent worker: one task at a time until both queues are exhausted.
  unsigned slab = 0;     // ring position, carried across tasks (same sequence on producer and consumers)
  for (;;) {
    if (tid == 0) s_task = df_next_task(a);
    __syncthreads();
    const int4 task = s_task;
    if (task.x < 0) break;
    df_worker(a, task, sm, bars, bars + kDfStages, slab);
    __syncthreads();                                   // everybody has read s_task and finished the task
  }
  if (tid >= kDfMath && tid < kDfMath + 32) asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
struct DfTaskHost { int key, type, i, j; };

static void build_tasks(int nb, bool want_kinv, std::vector<int4>* normal, std::vector<int4>* critical) {
  std::vector<DfTaskHost> t;
  // time of chain step j = 2j; a task's key = time after which its last input exists
  for (int j = 0; j < nb; ++j)
    for (int i = j + 2; i < nb; ++i) t.push_back({2 * j + 1, DF_TILE, i, j});
  for (int i = 1; i < nb; ++i)
    for (int j = i - 1; j >= 0; --j) t.push_back({2 * i + 1, DF_LINV, i, j});
  if (want_kinv)
    for (int i = nb - 1; i >= 0; --i)
      for (int j = 0; j <= i; ++j) t.push_back({2 * nb + 1, DF_KINV, i, j});
  std::stable_sort(t.begin(), t.end(), [](const DfTaskHost& x, const DfTaskHost& y) {
    if (x.key != y.key) return x.key < y.key;
    return x.type < y.type;           // TILE before LINV inside one step; then insertion order
  });
  normal->clear();
  for (const auto& e : t) normal->push_back(make_int4(e.type, e.i, e.j, e.key));
  // Critical queue: PART(j, w) in the order the chain consumes them.  Its last normal dependency is
  // TILE(j, j-2) (L[j, j-2]; the other inputs come earlier in the schedule or from the chain itself).
  critical->clear();
  for (int j = 1; j < nb; ++j) {
    int need = 0;
    if (j >= 2)
      for (size_t q = 0; q < normal->size(); ++q)
        if ((*normal)[q].x == DF_TILE && (*normal)[q].y == j && (*normal)[q].z == j - 2) { need = (int)q + 1; break; }
    critical->push_back(make_int4(DF_PART, j, j - 1, need));
    critical->push_back(make_int4(DF_PART, j, j, need));
  }
}

static bool df_enabled() {
  static const bool enabled = [] { const char* e = getenv("VZGP_DATAFLOW"); return !(e && e[0] == '0'); }();
  return enabled;
}

// Allocations and the task-list upload (not capturable): call before a stream capture that will contain
// chol_dataflow with the same (np, want_kinv).
int chol_dataflow_prepare(vzgp_handle* h, int np, bool want_kinv) {
  if (!df_enabled()) return 1;
  const int nb = np / 64;
  if (nb < 2 || nb > 256) return 1;
  const int w = want_kinv ? 1 : 0;
  if (h->df_nb[w] != nb) {
    std::vector<int4> tasks, crit;
    build_tasks(nb, want_kinv, &tasks, &crit);
    // one buffer: [normal queue | critical queue]
    VZ_TRY(h->df_tasks[w].reserve(sizeof(int4) * (tasks.size() + crit.size() + 1)));
    int4* d = h->df_tasks[w].as<int4>();
    VZ_CUDA(cudaMemcpyAsync(d, tasks.data(), sizeof(int4) * tasks.size(), cudaMemcpyHostToDevice, h->stream));
    VZ_CUDA(cudaMemcpyAsync(d + tasks.size(), crit.data(), sizeof(int4) * crit.size(), cudaMemcpyHostToDevice, h->stream));
    VZ_CUDA(cudaStreamSynchronize(h->stream));   // the vectors go out of scope
    h->df_ntasks[w] = (int)tasks.size();
    h->df_ncrit[w] = (int)crit.size();
    h->df_nb[w] = nb;
  }
  const size_t nflags = 8 + 2 * (size_t)nb * nb + 2 * (size_t)nb;
  VZ_TRY(h->df_flags.reserve(sizeof(int) * nflags));
  VZ_TRY(h->df_S.reserve(sizeof(double) * (size_t)nb * 2 * 4096));
  VZ_CUDA(cudaFuncSetAttribute(k_chol_dataflow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDfSmemBytes));
  return 0;
}

int chol_dataflow(vzgp_handle* h, double* L, double* Linv, double* LinvT, double* Kinv, int np, int* flag) {
  const int nb = np / 64;
  const bool want_kinv = Kinv != nullptr;
  {
    const int st = chol_dataflow_prepare(h, np, want_kinv);
    if (st != 0) return st;
  }
  const size_t nflags = 8 + 2 * (size_t)nb * nb + 2 * (size_t)nb;
  DfArgs a;
  a.L = L; a.X = Linv; a.Y = LinvT; a.S = h->df_S.as<double>(); a.Kinv = Kinv;
  int* f = h->df_flags.as<int>();
  a.ctrl = f; a.flagL = f + 8; a.flagY = a.flagL + nb * nb; a.flagS = a.flagY + nb * nb;
  a.tasks = h->df_tasks[want_kinv ? 1 : 0].as<int4>(); a.ntasks = h->df_ntasks[want_kinv ? 1 : 0];
  a.crit = a.tasks + a.ntasks; a.ncrit = h->df_ncrit[want_kinv ? 1 : 0];
  a.np = np; a.nb = nb; a.bad = flag;
  VZ_CUDA(cudaMemsetAsync(f, 0, sizeof(int) * nflags, h->stream));
  // Worker CTAs per launch.  A lone factorisation (the fit, a single evaluation) may take every slot
  // (2 CTAs per SM); the 4-5 concurrent evaluations of an ARD fit run one launch each on their own
  // streams and are given an equal share (vzgp_set_int(h, "dataflow_ctas", n), ard.py) so that all of
  // them are resident together.
  static const int env_cap = [] { const char* e = getenv("VZGP_DF_CTAS"); return e ? atoi(e) : 0; }();
  int cap = h->df_ctas > 0 ? h->df_ctas : (env_cap > 0 ? env_cap : 2 * h->sm_count - 8);
  if (cap < 8) cap = 8;
  const int workers = a.ntasks + a.ncrit < cap ? a.ntasks + a.ncrit : cap;
  k_chol_dataflow<<<1 + workers, kDfThreads, kDfSmemBytes, h->stream>>>(a);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}

// 1 if some wait inside the last dataflow launches timed out (synchronises the stream).
int chol_dataflow_timed_out(vzgp_handle* h, int* out) {
  *out = 0;
  if (!h->df_flags.ptr) return 0;
  VZ_CUDA(cudaMemcpyAsync(out, h->df_flags.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  VZ_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

}  // namespace vzgp

#ifdef VZ_DF_TIMING
// Debug builds only (make EXTRA=-DVZ_DF_TIMING): clock64 stamps of the chain CTA, [step][phase].
#ifdef VZ_LA_TIMING
extern "C" int vzgp_debug_la_timing(long long* out) {
  return cudaMemcpyFromSymbol(out, vzgp::g_la_t, sizeof(long long) * 64) == cudaSuccess ? 0 : -2;
}
#endif
extern "C" int vzgp_debug_df_timing(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, vzgp::g_df_t, sizeof(long long) * (n < 512 ? n : 512)) == cudaSuccess ? 0 : -2;
}
#endif
