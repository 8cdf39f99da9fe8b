// The W = K* . Linv^T contraction of the scoring path on the 5th-generation tensor cores (tcgen05, TMEM),
// by an Ozaki-style exact integer split of the fp64 operands.
//
// Same contract as k_score (score.cu): posterior mean / variance + UCB + trust region for a candidate pool;
// replaces BayesianScoringFunction.score_with_aux (acquisitions.py:177-207).  The FP64 tensor pipe (DMMA) tops
// out at ~37 TFLOP/s on this chip; tcgen05 has no f64 kind but multiplies 8-bit integers with exact 32-bit
// accumulation at ~100x that rate.  So both operands are written as fixed-point numbers of 7 base-256 digits
// relative to a power-of-two row scale, with BALANCED digits,
//     K*[i,k]   = 2^ea   * sum_{s=1..7} a_s[i,k] 2^(-8s)      a_s in [-128,127]
//     Linv[j,k] = 2^eb_j * sum_{t=1..7} b_t[j,k] 2^(-8t)      b_t in [-128,127]
// i.e. 55 bits below the row maximum - what fp64 carries for the entries that dominate the sum - and
//     W[i,j] = 2^(ea+eb_j) * sum_{g=2..8} 2^(-8g) G_g[i,j],   G_g = sum_{s+t=g} a_s b_t^T   (28 digit products)
// where every G_g is an EXACT int32 (|G_g| <= 7 * np * 128 * 128 < 2^31 for np <= 4096, the limit of this path).
// Products with s + t > 8 are dropped.  With balanced digits they have zero mean and sum to ~1e-15 of the
// row-scale product (unsigned digits would add a one-sided 3e-13: measured in tools/ozaki_emulation.py), the
// same order as the rounding of an fp64 accumulation.
//
// One persistent CTA per SM, 64 candidates per tile:
//   phase 1   (16 worker warps)  K* tile by 64-column steps as in k_score; mu and the L-inf distance on the
//             fly; every K* value is cut into its 7 digits and stored to a CTA-private scratch
//             [7 digit planes][64 candidates][np] (L2 resident).
//   phase 2   MMA D[128 x 64] (TMEM, s32) += A[128 x 32] B[64 x 32]^T with A = a digit plane of Linv rows
//             (j tile of 128), B = a digit plane of the candidates' K*; 7 accumulator groups (g) of 64 TMEM
//             columns.  A TMA producer thread streams, per 128-byte k chunk, the 7 K* planes (56 KB, double
//             buffered) and the 7 Linv planes (16 KB each through a 4-slot ring); ONE thread issues the 112
//             tcgen05.mma of the chunk and hands the buffers back with tcgen05.commit.
//   epilogue  (worker warps) tcgen05.ld the 7 groups, recombine in 64-bit integers, scale, square and add
//             into per-candidate sums; after the last j tile: variance, sigma, UCB, trust region.
#include <cuda.h>

#include <climits>
#include <cstring>

#include "launchers.h"
#include "score_small.cuh"
#include "tiles.cuh"

namespace vzgp {

namespace {

using GP1 = GemmCfg<64, 64, 16, 2, 4>;
constexpr int kWorkers = 512;                 // 16 worker warps: phase 1 + epilogue
constexpr int kI8Threads = kWorkers + 64;     // + TMA producer warp + MMA issuer warp
constexpr int kDigits = 7;
constexpr int kGroups = 7;                    // g = s + t - 2 in [0, 6]
constexpr int kJT = 128;                      // Linv rows per j tile = MMA M
constexpr int kKC = 128;                      // bytes (= k values) per chunk: one 128-byte swizzle atom per row
constexpr int kASlotBytes = kJT * kKC;        // 16 KB: one digit plane of a j tile x k chunk
constexpr int kASlots = 4;
constexpr int kBPlaneBytes = kTM * kKC;       // 8 KB
constexpr int kBBufBytes = kDigits * kBPlaneBytes;   // 56 KB: all digit planes of the candidates' k chunk
constexpr int kRingBytes = 2 * kBBufBytes + kASlots * kASlotBytes;   // 176 KB
constexpr int kTmemCols = 512;

struct I8Args {
  alignas(64) CUtensorMap mapK;   // K* digit scratch as u8 [grid*7*64 rows][np], box 64 x 128, SWIZZLE_128B
  alignas(64) CUtensorMap mapL;   // Linv digit planes as u8 [7][np][np], box 1 x 128 x 128, SWIZZLE_128B
  ScoreArgs s;                    // candidates, model, outputs (mapA / mapB / scratch unused)
  uint8_t* kdig;                  // [grid][7][64][np]
  const double* lscale;           // [np]  2^(ea + eb_j - 32)
  double kscale;                  // 2^(56 - ea)
};

// mbarrier wait that turns a protocol error into a trap (cudaErrorLaunchFailure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
  unsigned ok, spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0xfffu) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// K-major operand tile in the 128-byte-swizzle layout TMA writes: rows of 128 bytes, 8-row groups 1024 bytes
// apart (SBO), descriptor version 1, layout type 2 (SWIZZLE_128B).  The tile base is 1024-byte aligned; a
// 32-byte k step inside the atom adds 2 to the (16-byte unit) start address.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::i8 instruction descriptor: D = s32, A = B = s8, both K-major, M = 128, N = 64.
__host__ __device__ constexpr uint32_t umma_idesc_i8() {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTM >> 3) << 17) | ((uint32_t)(kJT >> 4) << 24);
}
// Power-of-two scale exponent e with |x| 2^-e <= 0.498 for |x| <= m: the top balanced digit stays in [-128, 127].
__host__ __device__ inline int balanced_scale_exp(double m) {
  int e = 0;
  const double f = frexp(m, &e);
  return f < 0.996 ? e + 1 : e + 2;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t addr, int32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

template <bool WITH_LINF>
__global__ void __launch_bounds__(kI8Threads, 1) k_score_i8(const __grid_constant__ I8Args ia) {
  const ScoreArgs& a = ia.s;
  extern __shared__ double smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem_raw) +
                  ((1024u - (static_cast<unsigned>(__cvta_generic_to_shared(smem_raw)) & 1023u)) & 1023u);
  constexpr int LD = kLD1;
  const int dc = a.kp.dc, dk = a.kp.dk, np = a.np;
  // phase 2 buffers; phase 1 staging aliases them (the phases do not overlap in time)
  uint8_t* bbuf = smem;                                  // [2][7][64][128]
  uint8_t* aring = smem + 2 * kBBufBytes;                // [kASlots][128][128]
  double* sa = reinterpret_cast<double*>(smem);          // [dc][LD]     candidates (transposed)
  double* sb = sa + dc * LD;                             // [2][dc][LD]  trials, double buffered
  double* s_alpha = reinterpret_cast<double*>(smem + kRingBytes);   // [2][64]
  double* s_mu = s_alpha + 128;                          // [64]
  double* s_linf = s_mu + 64;                            // [64]
  double* s_red = s_linf + 64;                           // [4][64] row sums of the four TMEM lane quarters
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 256);
  uint64_t* bfull = bars;              // [2]
  uint64_t* bempty = bars + 2;         // [2]
  uint64_t* afull = bars + 4;          // [kASlots]
  uint64_t* aempty = bars + 4 + kASlots;
  uint64_t* tfull = bars + 4 + 2 * kASlots;     // accumulators of a j tile complete
  uint64_t* tempty = tfull + 1;                 // ... drained by the 16 worker warps
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tempty + 1);
  int32_t* za = reinterpret_cast<int32_t*>(s_tmem + 2);  // [dk][LD]
  int32_t* zb = za + dk * LD;                            // [dk][LD]
  uint8_t* s_mask = reinterpret_cast<uint8_t*>(zb + dk * LD);  // [kMaxDc]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid / 16, tx = tid % 16;      // phase-1 mapping (32 x 16 threads), workers only
  const bool is_worker = warp < kWorkers / 32;
  const bool is_producer = warp == kWorkers / 32;
  const bool is_mma = warp == kWorkers / 32 + 1;
  uint8_t* kd = ia.kdig + (size_t)blockIdx.x * kDigits * kTM * np;
  if (tid < kMaxDc) s_mask[tid] = a.tr_mask[tid];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(bfull + i, 1); mbar_init(bempty + i, 1); }
    for (int i = 0; i < kASlots; ++i) { mbar_init(afull + i, 1); mbar_init(aempty + i, 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, kWorkers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (is_mma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  int clamped = 0;
  auto consumer_sync = [&]() { asm volatile("bar.sync 1, %0;\n" ::"n"(kWorkers) : "memory"); };

  const int ntiles = (a.M + kTM - 1) / kTM;
  const int njt = (np + kJT - 1) / kJT;
  const double* XTs = a.XT;
  const double* XTu = a.XT + (size_t)dc * np;
  unsigned b_n = 0, a_n = 0, jt_n = 0;   // B buffers / A slots / j tiles issued (producer), consumed (MMA thread, workers)

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * kTM;
    if (is_producer) {
      __syncthreads();   // phase 1 of this tile is complete: the digit scratch is visible, the staging smem is free
      if (lane == 0) {
        for (int jt = 0; jt < njt; ++jt) {
          for (int kc = 0; kc <= jt; ++kc) {
            const int buf = b_n & 1;
            mbar_wait_bounded(bempty + buf, ((b_n >> 1) & 1) ^ 1);
            mbar_expect_tx(bfull + buf, kBBufBytes);
            for (int s = 0; s < kDigits; ++s)
              tma_load_2d(bbuf + buf * kBBufBytes + s * kBPlaneBytes, &ia.mapK, kc * kKC,
                          ((int)blockIdx.x * kDigits + s) * kTM, bfull + buf);
            ++b_n;
            for (int t = 0; t < kDigits; ++t) {
              const int slot = a_n % kASlots;
              mbar_wait_bounded(aempty + slot, ((a_n / kASlots) & 1) ^ 1);
              mbar_expect_tx(afull + slot, kASlotBytes);
              tma_load_3d(aring + slot * kASlotBytes, &ia.mapL, kc * kKC, jt * kJT, t, afull + slot);
              ++a_n;
            }
          }
        }
      }
      __syncwarp();
      continue;
    }
    if (is_mma) {
      __syncthreads();
      if (lane == 0) {
        for (int jt = 0; jt < njt; ++jt) {
          mbar_wait_bounded(tempty, (jt_n & 1) ^ 1);   // the previous j tile's accumulators have been read
          tc_fence_after();
          unsigned touched = 0;
          for (int kc = 0; kc <= jt; ++kc) {
            const int buf = b_n & 1;
            mbar_wait_bounded(bfull + buf, (b_n >> 1) & 1);
            ++b_n;
            const int ksteps = (np - kc * kKC) >= kKC ? 4 : (np - kc * kKC + 31) / 32;   // the last chunk may be half
            const uint32_t bbase = smem_u32(bbuf + buf * kBBufBytes);
            for (int t = 0; t < kDigits; ++t) {
              const int slot = a_n % kASlots;
              mbar_wait_bounded(afull + slot, (a_n / kASlots) & 1);
              ++a_n;
              tc_fence_after();
              const uint64_t da = umma_desc_sw128(smem_u32(aring + slot * kASlotBytes));
              constexpr uint32_t idesc = umma_idesc_i8();
              for (int s = 0; s + t < kGroups; ++s) {     // digit pair (s+1, t+1): group g = s + t
                const int g = s + t;
                const uint64_t db = umma_desc_sw128(bbase + s * kBPlaneBytes);
                for (int kk = 0; kk < ksteps; ++kk) {
                  umma_i8(tmem + g * kTM, da + 2 * kk, db + 2 * kk, idesc, (touched >> g) & 1u);
                  touched |= 1u << g;
                }
              }
              umma_commit(aempty + slot);   // slot reusable once these MMAs have read it
            }
            umma_commit(bempty + buf);
          }
          umma_commit(tfull);
          ++jt_n;
        }
      }
      __syncwarp();
      continue;
    }
    // ================= worker warps =================
    consumer_sync();  // previous tile fully consumed by the workers (sa, s_mu, s_red)
    for (int e = tid; e < kTM * dc; e += kWorkers) {
      const int r = e / dc, d = e - r * dc;
      const int gr = m0 + r;
      const double v = gr < a.M ? __ldg(a.Xs + (size_t)gr * dc + d) : 0.0;
      sa[d * LD + r] = WITH_LINF ? v : v * a.kp.inv_ls_c[d];
    }
    if (dk > 0) stage_rows_T_i32(a.Zs, a.M, dk, m0, kTM, za, LD, kWorkers);

    // ---------------- phase 1: K* tile -> digits, mean, trust-region distance ----------------
    auto stage_trials = [&](int jb, int buf) {
      double* dst = sb + buf * dc * LD;
      const double* src = WITH_LINF ? XTu : XTs;
      for (int c = tid; c < dc * 32; c += kWorkers) {       // dc rows x 32 chunks of 16 B
        const int d = c >> 5, q = c & 31;
        cp_async16(dst + d * LD + q * 2, src + (size_t)d * np + jb * 64 + q * 2, true);
      }
      if (tid < 32) cp_async16(s_alpha + buf * 64 + tid * 2, a.alpha + jb * 64 + tid * 2, true);
    };
    double mu_part[2], lmin[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { mu_part[i] = 0.0; lmin[i] = INFINITY; }
    const int nj = np / 64;
    stage_trials(0, 0);
    cp_async_commit();
    for (int jb = 0; jb < nj; ++jb) {
      const int buf = jb & 1;
      if (jb + 1 < nj) stage_trials(jb + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      consumer_sync();
      if (dk > 0) {
        stage_rows_T_i32(a.Z, np, dk, jb * 64, 64, zb, LD, kWorkers);
        consumer_sync();
      }
      const double* sbj = sb + buf * dc * LD;
      const double* alj = s_alpha + buf * 64;
      double d2[2][4], lf[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { d2[i][j] = 0.0; lf[i][j] = 0.0; }
      for (int d = 0; d < dc; ++d) {
        const double2 av = *reinterpret_cast<const double2*>(sa + d * LD + GP1::row_of(ty, 0));
        const double2 b0 = *reinterpret_cast<const double2*>(sbj + d * LD + GP1::col_of(tx, 0));
        const double2 b1 = *reinterpret_cast<const double2*>(sbj + d * LD + GP1::col_of(tx, 2));
        const double aa[2] = {av.x, av.y}, bb[4] = {b0.x, b0.y, b1.x, b1.y};
        if (WITH_LINF) {
          const double w = a.kp.inv_ls2_c[d];
          const bool in_tr = s_mask[d] != 0;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double df = aa[i] - bb[j];
              d2[i][j] = fma(df * df, w, d2[i][j]);
              if (in_tr) lf[i][j] = fmax(lf[i][j], fabs(df));
            }
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double df = aa[i] - bb[j];
              d2[i][j] = fma(df, df, d2[i][j]);
            }
        }
      }
      for (int k = 0; k < dk; ++k) {
        const double w = a.kp.inv_ls2_k[k];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int avz = za[k * LD + GP1::row_of(ty, i)];
#pragma unroll
          for (int j = 0; j < 4; ++j) d2[i][j] += (avz != zb[k * LD + GP1::col_of(tx, j)]) ? w : 0.0;
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        long long q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cj = GP1::col_of(tx, j);
          const bool valid = (jb * 64 + cj) < a.n_valid;
          const double kv = valid ? matern52(d2[i][j], a.kp.sf2) : 0.0;
          mu_part[i] = fma(kv, alj[cj], mu_part[i]);
          if (WITH_LINF && (jb * 64 + cj) < a.tr_rows) lmin[i] = fmin(lmin[i], lf[i][j]);
          q[j] = __double2ll_rn(kv * ia.kscale);    // |q| < 2^55
        }
        uint8_t* row = kd + (size_t)GP1::row_of(ty, i) * np + jb * 64;
        // balanced base-256 digits, least significant first: digit = low byte (two's complement), carry (q + 128) >> 8
#pragma unroll
        for (int s = kDigits - 1; s >= 0; --s) {
          uint8_t* p = row + (size_t)s * kTM * np;
          *reinterpret_cast<uint16_t*>(p + GP1::col_of(tx, 0)) = (uint16_t)((q[0] & 255ll) | ((q[1] & 255ll) << 8));
          *reinterpret_cast<uint16_t*>(p + GP1::col_of(tx, 2)) = (uint16_t)((q[2] & 255ll) | ((q[3] & 255ll) << 8));
#pragma unroll
          for (int j = 0; j < 4; ++j) q[j] = (q[j] + 128) >> 8;
        }
      }
      consumer_sync();
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        mu_part[i] += __shfl_xor_sync(0xffffffffu, mu_part[i], o);
        if (WITH_LINF) lmin[i] = fmin(lmin[i], __shfl_xor_sync(0xffffffffu, lmin[i], o));
      }
      if (tx == 0) {
        s_mu[GP1::row_of(ty, i)] = mu_part[i];
        s_linf[GP1::row_of(ty, i)] = lmin[i];
      }
    }
    fence_proxy_async();  // generic-proxy writes (digit scratch, aliased smem) before async-proxy (TMA) accesses
    __syncthreads();      // this CTA's digit scratch is complete and visible

    // ---------------- epilogue of every j tile: TMEM -> sum_j W[i,j]^2 ----------------
    // warp w reads TMEM lanes 32 (w % 4) .. +31 (Linv rows j) and columns 16 (w / 4) .. +15 (candidates)
    const int quarter = warp & 3, cblock = warp >> 2;
    double rowsq[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) rowsq[c] = 0.0;
    for (int jt = 0; jt < njt; ++jt) {
      mbar_wait_bounded(tfull, jt_n & 1);
      ++jt_n;
      tc_fence_after();
      const int j = jt * kJT + quarter * 32 + lane;
      const double sc = j < np ? __ldg(ia.lscale + j) : 0.0;
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + cblock * 16;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        int32_t G[kGroups][4];
#pragma unroll
        for (int g = 0; g < kGroups; ++g) tmem_ld4(taddr + g * kTM + c4 * 4, G[g]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // sum_g G_g 2^(-8(g+2)) = 2^-32 (hi + lo 2^-32),  hi = G0 2^16 + G1 2^8 + G2,  lo = G3 2^24 + ... + G6
          const long long hi = ((long long)G[0][c] << 16) + ((long long)G[1][c] << 8) + (long long)G[2][c];
          const long long lo = ((long long)G[3][c] << 24) + ((long long)G[4][c] << 16) + ((long long)G[5][c] << 8) +
                               (long long)G[6][c];
          const double w = sc * fma((double)lo, 0x1p-32, (double)hi);
          rowsq[c4 * 4 + c] = fma(w, w, rowsq[c4 * 4 + c]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(tempty)) : "memory");
    }
    // sum over the 32 lanes (rows j) in a fixed order, then over the four lane quarters through smem
#pragma unroll
    for (int c = 0; c < 16; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rowsq[c] += __shfl_xor_sync(0xffffffffu, rowsq[c], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 16; ++c) s_red[quarter * 64 + cblock * 16 + c] = rowsq[c];
    }
    consumer_sync();
    if (tid < kTM) {
      const int r = tid, m = m0 + r;
      if (m < a.M) {
        const double rs = (s_red[r] + s_red[64 + r]) + (s_red[128 + r] + s_red[192 + r]);
        emit_score(a, m, rs, s_mu[r], s_linf[r], clamped);
      }
    }
  }
  if (clamped) atomicAdd(a.clamp_count, clamped);
  tc_fence_before();
  __syncthreads();
  if (is_mma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)kTmemCols));
}

// Linv (lower, fp64, [np x np]) -> 7 balanced base-256 digit planes relative to the row maximum, and the
// per-row scale 2^(ea + eb_j - 32) of the recombination.  One CTA per row.
__global__ void __launch_bounds__(128) k_slice_linv(const double* __restrict__ Linv, int np, int ea,
                                                    uint8_t* __restrict__ planes, double* __restrict__ lscale) {
  const int j = blockIdx.x, tid = threadIdx.x;
  __shared__ double s_max[4];
  const double* row = Linv + (size_t)j * np;
  double m = 0.0;
  for (int k = tid; k <= j; k += 128) m = fmax(m, fabs(row[k]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) s_max[tid >> 5] = m;
  __syncthreads();
  m = fmax(fmax(s_max[0], s_max[1]), fmax(s_max[2], s_max[3]));
  const int eb = (m > 0.0 && isfinite(m)) ? balanced_scale_exp(m) : 0;
  const double sc = ldexp(1.0, 56 - eb);
  for (int k = tid; k < np; k += 128) {
    const double x = (k <= j && isfinite(m)) ? row[k] : 0.0;
    long long q = __double2ll_rn(x * sc);
#pragma unroll
    for (int t = kDigits - 1; t >= 0; --t) {
      planes[((size_t)t * np + j) * np + k] = (uint8_t)(q & 255ll);
      q = (q + 128) >> 8;
    }
  }
  if (tid == 0) lscale[j] = ldexp(1.0, ea + eb - 32);
}

size_t score_i8_smem_bytes(int dk) {
  return 1024 + kRingBytes + sizeof(double) * (128 + 64 * 2 + 256) + sizeof(uint64_t) * (4 + 2 * kASlots + 2) + 16 +
         sizeof(int32_t) * dk * 2 * kLD1 + kMaxDc;
}

}  // namespace

bool score_i8_eligible(const vzgp_handle* h, int M) {
  if (h->kp.use_linear || h->np < kJT || h->np > 4096) return false;
  if ((size_t)3 * h->dc * kLD1 * sizeof(double) > (size_t)kRingBytes) return false;
  const int ntiles = (M + kTM - 1) / kTM;
  return ntiles >= h->sm_count;          // enough tiles for one CTA per SM (no column split on this path)
}

int launch_score_i8(vzgp_handle* h, const double* Xs, const int32_t* Zs, int M, const vzgp_acq* acq,
                    double* score, double* mu, double* sigma, double* linf) {
  const int np = h->np;
  const int ntiles = (M + kTM - 1) / kTM;
  const int grid = ntiles < h->sm_count ? ntiles : h->sm_count;
  const int ea = balanced_scale_exp(h->kp.sf2);   // K* <= sf2
  if (!h->i8_ready) {
    VZ_TRY(h->i8_planes.reserve((size_t)kDigits * np * np));
    VZ_TRY(h->i8_scale.reserve(sizeof(double) * np));
    k_slice_linv<<<np, 128, 0, h->stream>>>(h->Linv.as<double>(), np, ea, h->i8_planes.as<uint8_t>(), h->i8_scale.as<double>());
    VZ_CHECK_LAUNCH();
    h->launches++;
    h->i8_ready = true;
  }
  VZ_TRY(h->i8_kdig.reserve((size_t)grid * kDigits * kTM * np));
  I8Args ia;
  memset(&ia, 0, sizeof(ia));
  ScoreArgs& a = ia.s;
  a.Xs = Xs; a.Zs = Zs; a.M = M;
  a.XT = h->XT.as<double>(); a.Z = h->Z.as<int32_t>();
  a.np = np; a.n_valid = h->n_valid;
  a.Linv = h->Linv.as<double>(); a.ldi = np;
  a.alpha = h->alpha.as<double>();
  a.kp = h->kp; a.sn2 = h->sn2;
  a.coef = acq->ucb_coefficient;
  a.apply_tr = acq->use_trust_region ? 1 : 0;
  a.tr_rows = (acq->tr_rows > 0 && acq->tr_rows < h->n_valid) ? acq->tr_rows : h->n_valid;
  a.tr_strict = acq->tr_strict ? 1 : 0;
  a.radius = acq->trust_radius;
  for (int d = 0; d < kMaxDc; ++d)
    a.tr_mask[d] = (d < h->dc) ? (acq->tr_dim_mask ? (acq->tr_dim_mask[d] ? 1 : 0) : 1) : 0;
  a.nsplit = 1;
  a.mpad = ntiles * kTM;
  a.score = score; a.mu = mu; a.sigma = sigma; a.linf = linf;
  a.clamp_count = h->small.as<int>();
  ia.kdig = h->i8_kdig.as<uint8_t>();
  ia.lscale = h->i8_scale.as<double>();
  ia.kscale = ldexp(1.0, 56 - ea);
  {
    const uint64_t dims[2] = {(uint64_t)np, (uint64_t)grid * kDigits * kTM};
    const uint64_t strides[1] = {(uint64_t)np};
    const uint32_t box[2] = {(uint32_t)kKC, (uint32_t)kTM};
    VZ_TRY(make_tensor_map_u8(&ia.mapK, ia.kdig, 2, dims, strides, box));
  }
  {
    const uint64_t dims[3] = {(uint64_t)np, (uint64_t)np, (uint64_t)kDigits};
    const uint64_t strides[2] = {(uint64_t)np, (uint64_t)np * np};
    const uint32_t box[3] = {(uint32_t)kKC, (uint32_t)kJT, 1u};
    VZ_TRY(make_tensor_map_u8(&ia.mapL, h->i8_planes.as<uint8_t>(), 3, dims, strides, box));
  }
  const bool need_linf = (linf != nullptr) || (a.apply_tr && a.radius <= 0.5);
  const size_t sm = score_i8_smem_bytes(h->dk);
  if (sm > 227 * 1024) { set_error("k_score_i8 needs %zu bytes of shared memory", sm); return VZGP_ERR_UNSUPPORTED; }
  if (need_linf) {
    VZ_CUDA(cudaFuncSetAttribute(k_score_i8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_score_i8<true><<<grid, kI8Threads, sm, h->stream>>>(ia);
  } else {
    VZ_CUDA(cudaFuncSetAttribute(k_score_i8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_score_i8<false><<<grid, kI8Threads, sm, h->stream>>>(ia);
  }
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}

}  // namespace vzgp
