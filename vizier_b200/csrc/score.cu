// Fused posterior mean/variance + UCB + trust-region scoring over a candidate pool, the
// Philox candidate generator and top-k selection.
//
// Replaces (reference): BayesianScoringFunction.score_with_aux
// (vizier/_src/algorithms/designers/gp/acquisitions.py:177-207) = predict_with_aux
// (stochastic_process_model.py:800-868, TFP posterior_predictive) + UCB (:213-225) +
// _apply_trust_region (:152-174) + TrustRegion.min_linf_distance (:779-820); the candidate
// source of RandomVectorizedStrategy (random_vectorized_optimizer.py:78-100) and the top-k
// bookkeeping of vectorized_base.py:544-587.
//
// One persistent CTA per SM walks 64-candidate tiles:
//   phase 1  K* tile [64 x np] = Matern(x*, X) built 64 columns at a time from shared-memory
//            staged rows; mu = K* alpha and the L-inf trust-region distance are reduced on the
//            fly; the tile goes to a CTA-private scratch (L2 resident, never re-read by others).
//   phase 2  W = K* . Linv^T in 64 x 128 blocks on the FP64 tensor pipe (mma.sync m8n8k4 f64;
//            tcgen05 has no f64 kind), operands streamed by a TMA producer warp (cp.async.bulk.tensor,
//            128-byte swizzle) through a 4-stage ring with full/empty mbarriers, exploiting that
//            Linv is lower triangular (k <= j, all-zero fragments skipped); each block is squared
//            and row-summed in registers, W is never stored.
//   epilogue var = sf2 + sn2 - sum W^2 (clamped at 0), sigma, UCB, trust region, outputs.
#include <cuda.h>

#include <climits>
#include <cstring>

#include "launchers.h"
#include "score_small.cuh"
#include "topk_merge.cuh"
#include "tiles.cuh"

namespace vzgp {

using GP1 = GemmCfg<64, 64, 16, 2, 4>;  // phase-1 thread mapping: 512 threads, 2x4 outputs each
constexpr int kThreads = 512;        // consumer threads (16 math warps)
constexpr int kBlockThreads = 544;   // + one TMA producer warp

// Phase-2 tiling: 64 candidates x 128 output columns per pass, k-slabs of 32, 4-stage TMA ring.
// 16 warps as 4 (M) x 4 (N): each warp owns a 16 x 32 block = 2 x 4 DMMA tiles.
constexpr int kBN = 128;         // output columns per pass
constexpr int kBK = 32;          // k-slab = two TMA boxes of 16 doubles (one 128-byte swizzle atom per row)
constexpr int kStages = 4;
// One stage: A half0 | A half1 | B half0 | B half1, each a dense [rows][16] box written by TMA with
// the 128-byte swizzle (16-byte chunk c of row r lands at chunk c ^ (r & 7)).
constexpr int kAHalf = kTM * 16;                 // doubles
constexpr int kBHalf = kBN * 16;
constexpr int kStageDoubles = 2 * (kAHalf + kBHalf);   // 6144 doubles = 48 KB
constexpr unsigned kStageBytes = kStageDoubles * sizeof(double);

template <bool WITH_LINF>
__global__ void __launch_bounds__(kBlockThreads, 1) k_score(const __grid_constant__ ScoreArgs a) {
  extern __shared__ double smem_raw[];
  // the swizzled TMA boxes need a 1024-byte aligned base
  // (pointer arithmetic on the shared-space pointer so the compiler keeps LDS/STS addressing)
  double* smem = smem_raw + (((1024u - (static_cast<unsigned>(__cvta_generic_to_shared(smem_raw)) & 1023u)) & 1023u) >> 3);
  constexpr int LD = kLD1;
  const int dc = a.kp.dc, dk = a.kp.dk, np = a.np;
  double* ring = smem;                                   // [kStages][kStageDoubles]
  // Without the trust-region distance the features are pre-divided by the length scale (the
  // reference's FeatureScaled form, 2 flops per dimension).  With it, unscaled features are staged
  // so that |a-b| is exact, and the scaling is applied to the squared difference.
  // Phase 1 and phase 2 never overlap in time, so the phase-1 staging buffers alias the ring.
  double* sa = ring;                                     // [dc][LD]     candidates (transposed)
  double* sb = sa + dc * LD;                             // [2][dc][LD]  trials, double buffered
  double* s_alpha = ring + (kStages * kStageDoubles > 3 * kMaxDc * kLD1 ? kStages * kStageDoubles : 3 * kMaxDc * kLD1);  // [2][64]
  double* s_mu = s_alpha + 128;                          // [64]
  double* s_linf = s_mu + 64;                            // [64]
  double* s_rowsq = s_linf + 64;                         // [4][64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_rowsq + 256);  // [kStages] TMA bytes landed
  uint64_t* empty_bar = full_bar + kStages;                         // [kStages] all 16 math warps released the stage
  int32_t* za = reinterpret_cast<int32_t*>(full_bar + 8);  // [dk][LD]
  int32_t* zb = za + dk * LD;                            // [dk][LD]
  uint8_t* s_mask = reinterpret_cast<uint8_t*>(zb + dk * LD);  // [kMaxDc]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid / 16, tx = tid % 16;      // phase-1 mapping (32 x 16 threads)
  const int wm = warp & 3, wn = warp >> 2;     // phase-2 warp grid 4 (M) x 4 (N)
  const int fr = lane >> 2, fk = lane & 3;     // fragment row / k within a DMMA tile
  double* scr = a.scratch + (size_t)blockIdx.x * kTM * np;
  if (tid < kMaxDc) s_mask[tid] = a.tr_mask[tid];
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, kThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  int clamped = 0;
  unsigned slab_n = 0;   // slabs issued (producer) / consumed (math warps) since kernel start: ring position and phase
  const bool is_producer = warp == kThreads / 32;
  auto consumer_sync = [&]() { asm volatile("bar.sync 1, %0;\n" ::"n"(kThreads) : "memory"); };

  const int ntiles = (a.M + kTM - 1) / kTM;
  const int nblocks = (np + kBN - 1) / kBN;
  const int nsplit = a.nsplit;
  const double* XTs = a.XT;
  const double* XTu = a.XT + (size_t)dc * np;

  for (int work = blockIdx.x; work < ntiles * nsplit; work += gridDim.x) {
    const int tile = work / nsplit, split = work - tile * nsplit;
    const int m0 = tile * kTM;
    // Work split of this tile's phase 2 (identical for producer and consumers).
    const int nq = (nsplit == 1) ? nblocks : ((split == nblocks - 1 - split) ? 1 : 2);
    auto block_of = [&](int q) { return (nsplit == 1) ? q : (q == 0 ? split : nblocks - 1 - split); };
    auto slabs_in = [&](int jb) { int kend = (jb + 1) * kBN; if (kend > np) kend = np; return kend / kBK; };
    if (is_producer) {
      // ---- TMA producer warp: waits for phase 1 of this tile, then streams every slab of the
      // tile through the ring, gated only by the per-stage empty barriers.
      __syncthreads();
      if (lane == 0) {
        for (int q = 0; q < nq; ++q) {
          const int jb = block_of(q), nsl = slabs_in(jb);
          for (int ks = 0; ks < nsl; ++ks) {
            const int stage = slab_n % kStages;
            mbar_wait(empty_bar + stage, ((slab_n / kStages) & 1) ^ 1);
            double* base = ring + stage * kStageDoubles;
            const int k0 = ks * kBK;
            mbar_expect_tx(full_bar + stage, kStageBytes);
            tma_load_2d(base, &a.mapA, k0, (int)blockIdx.x * kTM, full_bar + stage);
            tma_load_2d(base + kAHalf, &a.mapA, k0 + 16, (int)blockIdx.x * kTM, full_bar + stage);
            tma_load_2d(base + 2 * kAHalf, &a.mapB, k0, jb * kBN, full_bar + stage);          // rows >= np: zero fill
            tma_load_2d(base + 2 * kAHalf + kBHalf, &a.mapB, k0 + 16, jb * kBN, full_bar + stage);
            ++slab_n;
          }
        }
      }
      __syncwarp();
      continue;
    }
    consumer_sync();  // previous tile fully consumed by the math warps (sa, s_mu, s_rowsq, ring)
    // candidate tile, transposed; scaled by 1/ls like the reference's FeatureScaled kernel
    for (int e = tid; e < kTM * dc; e += kThreads) {
      const int r = e / dc, d = e - r * dc;
      const int gr = m0 + r;
      const double v = gr < a.M ? __ldg(a.Xs + (size_t)gr * dc + d) : 0.0;
      sa[d * LD + r] = WITH_LINF ? v : v * a.kp.inv_ls_c[d];
    }
    if (dk > 0) stage_rows_T_i32(a.Zs, a.M, dk, m0, kTM, za, LD, kThreads);

    // ---------------- phase 1: K* tile, mean, trust-region distance ----------------
    // Trial rows arrive pre-transposed and pre-scaled (XT), 64 columns per step, through a
    // two-deep cp.async buffer so the copy of block jb+1 overlaps the math of block jb.
    auto stage_trials = [&](int jb, int buf) {
      double* dst = sb + buf * dc * LD;
      const double* src = WITH_LINF ? XTu : XTs;
      for (int c = tid; c < dc * 32; c += kThreads) {       // dc rows x 32 chunks of 16 B
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
      if (jb + 1 < nj) stage_trials(jb + 1, buf ^ 1);   // buffer buf^1 was released by the barrier below
      cp_async_commit();
      cp_async_wait<1>();
      consumer_sync();
      if (dk > 0) {  // categorical rows are rare: staged synchronously
        stage_rows_T_i32(a.Z, np, dk, jb * 64, 64, zb, LD, kThreads);
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
        double kv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cj = GP1::col_of(tx, j);
          const bool valid = (jb * 64 + cj) < a.n_valid;
          kv[j] = valid ? matern52(d2[i][j], a.kp.sf2) : 0.0;
          mu_part[i] = fma(kv[j], alj[cj], mu_part[i]);
          if (WITH_LINF && (jb * 64 + cj) < a.tr_rows) lmin[i] = fmin(lmin[i], lf[i][j]);
        }
        double* dst = scr + (size_t)GP1::row_of(ty, i) * np + jb * 64;
        *reinterpret_cast<double2*>(dst + GP1::col_of(tx, 0)) = make_double2(kv[0], kv[1]);
        *reinterpret_cast<double2*>(dst + GP1::col_of(tx, 2)) = make_double2(kv[2], kv[3]);
      }
      consumer_sync();  // everyone is done with buffer `buf` (and zb) before it is refilled
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
    fence_proxy_async();  // generic-proxy writes (scratch tile, aliased smem) before async-proxy (TMA) accesses
    __syncthreads();      // this CTA's scratch tile is complete and visible

    // ---------------- phase 2: row sums of (K* Linv^T)^2 on the DMMA pipe ----------------
    // Slab (jb, ks): A = scratch[0:64, ks*16 : +16], B = Linv[jb*128 : +128, ks*16 : +16];
    // block jb needs ks < min(np, (jb+1)*128)/16 because Linv is lower triangular.  The slab
    // stream is flattened over blocks so the TMA ring never drains between blocks.
    // With nsplit > 1 this CTA takes blocks {split, nblocks-1-split} (balanced triangular work).
    double rowsq[2] = {0.0, 0.0};
    for (int q = 0; q < nq; ++q) {
      const int jb = block_of(q);
      double acc[2][4][2];
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) { acc[f][g][0] = 0.0; acc[f][g][1] = 0.0; }
      const int nsl = slabs_in(jb);
      const int col_base = jb * kBN + wn * 32;   // first output column of this warp
      for (int ks = 0; ks < nsl; ++ks) {
        const int stage = slab_n % kStages;
        mbar_wait(full_bar + stage, (slab_n / kStages) & 1);   // TMA bytes of this slab have landed
        ++slab_n;
        // Fragment loads are 16 bytes: lane (fr, fk) takes k = 8h + 2fk and 8h + 2fk + 1 of each
        // 8-wide k group h, i.e. the operands of two DMMA k-steps (the k order inside a slab is
        // irrelevant as long as A and B agree).  Row r = ... + fr, so the swizzle XOR is fr.
        const double* stg = ring + stage * kStageDoubles;
        const double* Arow0 = stg + (wm * 16 + fr) * 16;            // + half*kAHalf + chunk*2
        const double* Brow0 = stg + 2 * kAHalf + (wn * 32 + fr) * 16;
        const int k0 = ks * kBK;
        // Linv[c, k] = 0 for k > c.  k0 and col_base are multiples of 32: slabs right of this
        // warp's 32 columns contribute nothing; in the diagonal slab (k0 == col_base) the k group
        // h only reaches column fragments g >= h (static pattern, no predication).
        if (k0 < col_base) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int co = ((((h & 1) * 4 + fk) ^ fr) * 2);   // swizzled 16-byte chunk, in doubles
            const double2 a0 = *reinterpret_cast<const double2*>(Arow0 + (h >> 1) * kAHalf + co);
            const double2 a1 = *reinterpret_cast<const double2*>(Arow0 + (h >> 1) * kAHalf + 8 * 16 + co);
            double2 b[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) b[g] = *reinterpret_cast<const double2*>(Brow0 + (h >> 1) * kBHalf + g * 8 * 16 + co);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              dmma_8x8x4(acc[0][g][0], acc[0][g][1], a0.x, b[g].x);
              dmma_8x8x4(acc[1][g][0], acc[1][g][1], a1.x, b[g].x);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              dmma_8x8x4(acc[0][g][0], acc[0][g][1], a0.y, b[g].y);
              dmma_8x8x4(acc[1][g][0], acc[1][g][1], a1.y, b[g].y);
            }
          }
        } else if (k0 == col_base) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int co = ((((h & 1) * 4 + fk) ^ fr) * 2);
            const double2 a0 = *reinterpret_cast<const double2*>(Arow0 + (h >> 1) * kAHalf + co);
            const double2 a1 = *reinterpret_cast<const double2*>(Arow0 + (h >> 1) * kAHalf + 8 * 16 + co);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (g < h) continue;  // compile-time: columns of fragment g lie left of k group h
              const double2 bg = *reinterpret_cast<const double2*>(Brow0 + (h >> 1) * kBHalf + g * 8 * 16 + co);
              dmma_8x8x4(acc[0][g][0], acc[0][g][1], a0.x, bg.x);
              dmma_8x8x4(acc[1][g][0], acc[1][g][1], a1.x, bg.x);
              dmma_8x8x4(acc[0][g][0], acc[0][g][1], a0.y, bg.y);
              dmma_8x8x4(acc[1][g][0], acc[1][g][1], a1.y, bg.y);
            }
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(empty_bar + stage)) : "memory");
      }
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          rowsq[f] = fma(acc[f][g][0], acc[f][g][0], rowsq[f]);
          rowsq[f] = fma(acc[f][g][1], acc[f][g][1], rowsq[f]);
        }
    }
    cp_async_wait<0>();
    // combine the 4 lanes that share a fragment row, then the four N-warps through smem
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      rowsq[f] += __shfl_xor_sync(0xffffffffu, rowsq[f], 1);
      rowsq[f] += __shfl_xor_sync(0xffffffffu, rowsq[f], 2);
      if (fk == 0) s_rowsq[wn * 64 + wm * 16 + f * 8 + fr] = rowsq[f];
    }
    consumer_sync();
    // ---------------- epilogue ----------------
    if (tid < kTM) {
      const int r = tid, m = m0 + r;
      if (m < a.M) {
        const double rs = (s_rowsq[r] + s_rowsq[64 + r]) + (s_rowsq[128 + r] + s_rowsq[192 + r]);
        if (nsplit == 1) {
          emit_score(a, m, rs, s_mu[r], s_linf[r], clamped);
        } else {
          a.part[(size_t)split * a.mpad + m] = rs;
          if (split == 0) {
            a.part[(size_t)nsplit * a.mpad + m] = s_mu[r];
            a.part[(size_t)(nsplit + 1) * a.mpad + m] = s_linf[r];
          }
        }
      }
    }
  }
  if (clamped) atomicAdd(a.clamp_count, clamped);
}

// nsplit > 1: sums the partial row sums in a fixed order and emits the scores.
__global__ void k_score_finalize(const ScoreArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  int clamped = 0;
  if (m < a.M) {
    double rs = 0.0;
    for (int s = 0; s < a.nsplit; ++s) rs += a.part[(size_t)s * a.mpad + m];
    emit_score(a, m, rs, a.part[(size_t)a.nsplit * a.mpad + m], a.part[(size_t)(a.nsplit + 1) * a.mpad + m], clamped);
  }
  if (clamped) atomicAdd(a.clamp_count, clamped);
}

// Stand-alone kernels of the small-pool path (device functions in score_small.cuh).
template <bool WITH_LINF>
__global__ void __launch_bounds__(kSmallThreads) k_cross_small(const ScoreArgs a) {
  extern __shared__ double smem_raw[];
  cross_small_block<WITH_LINF>(a, blockIdx.x, blockIdx.y >> 2, blockIdx.y & 3, smem_raw);
}
__global__ void __launch_bounds__(kSmallThreads) k_var_small(const __grid_constant__ ScoreArgs a) {
  extern __shared__ double smem_raw[];
  var_small_dispatch(a, blockIdx.x, blockIdx.y, smem_raw);
}
template <bool WITH_LINF>
__global__ void __launch_bounds__(256) k_small_finalize(const ScoreArgs a) {
  const int m = blockIdx.x * 32 + (threadIdx.x >> 3);
  int clamped = 0;
  small_finalize_8<WITH_LINF>(a, m, threadIdx.x & 7, m < a.M, clamped);
  if (clamped) atomicAdd(a.clamp_count, clamped);
}

size_t cross_small_smem_bytes(int dc, int dk) { return sizeof(double) * 2 * dc * kLD1 + sizeof(int32_t) * 2 * dk * kLD1; }
size_t var_small_smem_bytes() {
  const size_t cp = sizeof(double) * kVarStages * kVarStageDoubles, tma = 1024 + sizeof(double) * kVarStages * kVarTmaStageDoubles;
  return cp > tma ? cp : tma;
}

// ---- host side: TMA descriptors ------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  // function-local static with an initialiser: thread-safe (handles are driven from several host threads)
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    EncodeTiledFn f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      f = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
    return f;
  }();
  return fn;
}

// fp64 matrix [rows x cols] with row pitch ld (elements); box = box_rows x 16 doubles, 128-byte swizzle.
static int make_map(CUtensorMap* m, const double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    bool l2_promotion = true) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return VZGP_ERR_CUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * sizeof(double)};
  cuuint32_t box[2] = {16, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  l2_promotion ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return VZGP_ERR_CUDA; }
  return 0;
}

int make_tensor_map_f64(void* map, const double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  // no L2 promotion: the dataflow kernel loads tiles whose NEIGHBOURS are still being written by other CTAs
  static const bool promo = [] { const char* e = getenv("VZGP_DF_L2PROMO"); return e && e[0] == '1'; }();
  return make_map(static_cast<CUtensorMap*>(map), base, rows, cols, ld, box_rows, promo);
}

int make_tensor_map_u8(void* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                       const uint32_t* box, bool promote_256) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return VZGP_ERR_CUDA; }
  cuuint64_t gdim[3], gstride[2];
  cuuint32_t bx[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides[i];
  CUresult r = fn(static_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstride, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  promote_256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (u8) failed with CUresult %d", (int)r); return VZGP_ERR_CUDA; }
  return 0;
}

size_t score_smem_bytes(int dc, int dk, bool with_linf) {
  (void)with_linf; (void)dc;
  const size_t big = kStages * kStageDoubles > 3 * kMaxDc * kLD1 ? kStages * kStageDoubles : 3 * kMaxDc * kLD1;
  return 1024 + sizeof(double) * (big + 128 + 64 * 2 + 256 + 8) +
         sizeof(int32_t) * dk * 2 * kLD1 + kMaxDc;
}

// K* scratch of `bytes` bytes on the handle.  It is written and re-read by the same CTA tile after tile:
// pin it in L2 (persisting access-policy window) so its dirty lines are overwritten in place instead
// of being evicted to HBM.  Best effort: failures only cost DRAM write-backs.
static int ensure_scratch(vzgp_handle* h, size_t scratch_bytes) {
  if (scratch_bytes > h->scratch.bytes || h->scratch_window != h->scratch.ptr) {
    VZ_TRY(h->scratch.reserve(scratch_bytes));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
      size_t want = h->scratch.bytes;
      if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof(attr));
      size_t win = h->scratch.bytes;
      if (win > (size_t)prop.accessPolicyMaxWindowSize) win = (size_t)prop.accessPolicyMaxWindowSize;
      attr.accessPolicyWindow.base_ptr = h->scratch.ptr;
      attr.accessPolicyWindow.num_bytes = win;
      attr.accessPolicyWindow.hitRatio = win > 0 ? (float)((double)want / (double)win > 1.0 ? 1.0 : (double)want / (double)win) : 0.f;
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
    h->scratch_window = h->scratch.ptr;
  }
  return 0;
}

static void fill_score_args(vzgp_handle* h, const double* Xs, const int32_t* Zs, int M, const vzgp_acq* acq,
                            double* score, double* mu, double* sigma, double* linf, ScoreArgs* pa) {
  ScoreArgs& a = *pa;
  const int ntiles = (M + kTM - 1) / kTM;
  a.Xs = Xs; a.Zs = Zs; a.M = M;
  a.XT = h->XT.as<double>(); a.Z = h->Z.as<int32_t>();
  a.np = h->np; a.n_valid = h->n_valid;
  a.Linv = h->Linv.as<double>(); a.ldi = h->np;
  a.alpha = h->alpha.as<double>();
  a.kp = h->kp; a.sn2 = h->sn2;
  a.coef = acq->ucb_coefficient;
  a.apply_tr = acq->use_trust_region ? 1 : 0;
  a.tr_rows = (acq->tr_rows > 0 && acq->tr_rows < h->n_valid) ? acq->tr_rows : h->n_valid;
  a.tr_strict = acq->tr_strict ? 1 : 0;
  a.radius = acq->trust_radius;
  for (int d = 0; d < kMaxDc; ++d)
    a.tr_mask[d] = (d < h->dc) ? (acq->tr_dim_mask ? (acq->tr_dim_mask[d] ? 1 : 0) : 1) : 0;
  a.scratch = h->scratch.as<double>();
  a.mpad = ntiles * kTM;
  a.nsplit = 1;
  a.box_rows = kTM;
  static const int tma_small = [] { const char* e = getenv("VZGP_SMALL_TMA"); return e ? atoi(e) : 1; }();   // 0: cp.async variant
  a.use_tma = tma_small;
  a.part = a.part_rs = a.part_mu = a.part_linf = nullptr;
  a.score = score; a.mu = mu; a.sigma = sigma; a.linf = linf;
  a.clamp_count = h->small.as<int>();  // slot 0
}

int prepare_small_score(vzgp_handle* h, const double* Xs, const int32_t* Zs, int M, const vzgp_acq* acq,
                        double* score, double* mu, double* sigma, double* linf, ScoreArgs* a, bool* with_linf) {
  const int ntiles = (M + kTM - 1) / kTM;
  VZ_TRY(ensure_scratch(h, (size_t)ntiles * kTM * h->np * sizeof(double)));
  fill_score_args(h, Xs, Zs, M, acq, score, mu, sigma, linf, a);
  const int nvb = h->np / kVarCols, nmb = h->np / 64;
  VZ_TRY(h->Tws.reserve(sizeof(double) * (size_t)(nvb + 2 * nmb) * a->mpad));
  a->part_rs = h->Tws.as<double>();
  a->part_mu = a->part_rs + (size_t)nvb * a->mpad;
  a->part_linf = a->part_mu + (size_t)nmb * a->mpad;
  *with_linf = (linf != nullptr) || (a->apply_tr && a->radius <= 0.5);
  // TMA boxes of the W phase: K* rows of one tile x 16 doubles, and 8 rows of Linv x 16 doubles.
  a->box_rows = ntiles == 1 ? ((M + 7) / 8) * 8 : kTM;
  VZ_TRY(make_map(&a->mapA, a->scratch, (uint64_t)ntiles * kTM, (uint64_t)h->np, (uint64_t)h->np, (uint32_t)a->box_rows));
  VZ_TRY(make_map(&a->mapB, a->Linv, (uint64_t)h->np, (uint64_t)h->np, (uint64_t)h->np, kVarCols));
  return 0;
}

// ---------------------------------------------------------------------------
// General scoring path: explicit K* and W = K* Linv^T, then one warp per candidate.  Taken by models the
// fused kernels do not cover - today the `linear_coef` variant (tuned_gp_models.py:203-245), whose kernel
// is Matern + feature-scaled linear (so k(x*, x*) is not constant) and whose GP has a constant mean.
// Same results contract as k_score; ~3x the HBM/L2 traffic and DFMA instead of DMMA (k_gemm_nt_tri).
// ---------------------------------------------------------------------------
struct GeneralArgs {
  const double* Xs;      // [mp x dc] padded candidates of this chunk
  const double* Ks;      // [mp x np]
  const double* W;       // [mp x np]
  const double* X;       // [np x dc] trials
  const double* alpha;
  int mc, np, n_valid, dc;
  KernelParams kp;
  double sn2, mean_const, coef, radius;
  int apply_tr, tr_rows, tr_strict, want_linf;
  uint8_t tr_mask[kMaxDc];
  double* score; double* mu; double* sigma; double* linf;
  int* clamp_count;
};

__global__ void __launch_bounds__(256) k_general_finalize(GeneralArgs a) {
  const int m = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  if (m >= a.mc) return;
  const double* ks = a.Ks + (size_t)m * a.np;
  const double* w = a.W + (size_t)m * a.np;
  double mean = 0.0, rs = 0.0;
  for (int j = lane; j < a.n_valid; j += 32) mean = fma(ks[j], a.alpha[j], mean);
  for (int j = lane; j < a.np; j += 32) rs = fma(w[j], w[j], rs);
  double dist = INFINITY;
  if (a.want_linf) {
    for (int n = lane; n < a.tr_rows; n += 32) {
      double mx = 0.0;
      for (int d = 0; d < a.dc; ++d)
        if (a.tr_mask[d]) mx = fmax(mx, fabs(a.Xs[(size_t)m * a.dc + d] - a.X[(size_t)n * a.dc + d]));
      dist = fmin(dist, mx);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mean += __shfl_xor_sync(0xffffffffu, mean, o);
    rs += __shfl_xor_sync(0xffffffffu, rs, o);
    dist = fmin(dist, __shfl_xor_sync(0xffffffffu, dist, o));
  }
  if (lane != 0) return;
  double kss = a.kp.sf2;
  if (a.kp.use_linear) {
    double uu = 0.0;
    for (int d = 0; d < a.dc; ++d) {
      const double u = fma(a.Xs[(size_t)m * a.dc + d], a.kp.inv_ls_c[d], -a.kp.lin_b);
      uu = fma(u, u, uu);
    }
    kss = fma(a.kp.lin_a, uu, kss);
  }
  mean += a.mean_const;
  double var = kss - rs + a.sn2;
  if (var < 0.0) { var = 0.0; atomicAdd(a.clamp_count, 1); }
  const double sd = sqrt(var);
  double sc = fma(a.coef, sd, mean);
  if (a.apply_tr) {
    const bool inside = (a.tr_strict ? (dist < a.radius) : (dist <= a.radius)) || (a.radius > 0.5);
    sc = inside ? sc : (-1e4 - dist);
  }
  a.score[m] = sc;
  if (a.mu) a.mu[m] = mean;
  if (a.sigma) a.sigma[m] = sd;
  if (a.linf) a.linf[m] = dist;
}

static int launch_score_general(vzgp_handle* h, const double* Xs, const int32_t* Zs, int M, const vzgp_acq* acq,
                                double* score, double* mu, double* sigma, double* linf) {
  const int np = h->np, dc = h->dc, dk = h->dk;
  constexpr int kChunk = 4096;
  const size_t nks = (size_t)kChunk * np, nx = (size_t)kChunk * (dc > 0 ? dc : 1);
  VZ_TRY(h->gen.reserve(sizeof(double) * (2 * nks + nx) + sizeof(int32_t) * (size_t)kChunk * (dk > 0 ? dk : 1)));
  double* Ks = h->gen.as<double>();
  double* W = Ks + nks;
  double* Xp = W + nks;
  int32_t* Zp = reinterpret_cast<int32_t*>(Xp + nx);
  GeneralArgs a;
  a.Ks = Ks; a.W = W; a.Xs = Xp; a.X = h->X.as<double>(); a.alpha = h->alpha.as<double>();
  a.np = np; a.n_valid = h->n_valid; a.dc = dc; a.kp = h->kp; a.sn2 = h->sn2; a.mean_const = h->mean_const;
  a.coef = acq->ucb_coefficient; a.radius = acq->trust_radius;
  a.apply_tr = acq->use_trust_region ? 1 : 0;
  a.tr_rows = (acq->tr_rows > 0 && acq->tr_rows < h->n_valid) ? acq->tr_rows : h->n_valid;
  a.tr_strict = acq->tr_strict ? 1 : 0;
  a.want_linf = (linf != nullptr) || (a.apply_tr && a.radius <= 0.5);
  for (int d = 0; d < kMaxDc; ++d) a.tr_mask[d] = (d < dc) ? (acq->tr_dim_mask ? (acq->tr_dim_mask[d] ? 1 : 0) : 1) : 0;
  a.clamp_count = h->small.as<int>();
  for (int m0 = 0; m0 < M; m0 += kChunk) {
    const int mc = M - m0 < kChunk ? M - m0 : kChunk, mp = round_up(mc, 64);
    if (dc > 0) VZ_TRY(launch_pad_rows(h, Xs + (size_t)m0 * dc, mc, dc, mp, Xp));
    if (dk > 0) VZ_TRY(launch_pad_rows_i32(h, Zs + (size_t)m0 * dk, mc, dk, mp, Zp));
    VZ_TRY(launch_cross_kernel(h, Xp, Zp, mp, h->X.as<double>(), h->Z.as<int32_t>(), np, h->n_valid, h->kp, Ks, np));
    VZ_TRY(launch_gemm_nt_tri(h, Ks, np, mp, h->Linv.as<double>(), np, np, W, np));
    a.mc = mc;
    a.score = score + m0; a.mu = mu ? mu + m0 : nullptr; a.sigma = sigma ? sigma + m0 : nullptr;
    a.linf = linf ? linf + m0 : nullptr;
    k_general_finalize<<<(mc + 7) / 8, 256, 0, h->stream>>>(a);
    VZ_CHECK_LAUNCH();
    h->launches++;
  }
  return 0;
}

int launch_score(vzgp_handle* h, const double* Xs, const int32_t* Zs, int M, const vzgp_acq* acq,
                 double* score, double* mu, double* sigma, double* linf) {
  if (M <= 0) return 0;
  if (h->kp.use_linear) return launch_score_general(h, Xs, Zs, M, acq, score, mu, sigma, linf);
  const int ntiles = (M + kTM - 1) / kTM;
  const int nblocks = (h->np + kBN - 1) / kBN;
  // Pools of a few tiles (acquisition-optimiser batches) take the trial-axis decomposition.
  static const int small_tiles_max = [] {
    const char* e = getenv("VZGP_SMALL_TILES");   // tuning / test hook; 0 disables the small-pool path
    return e ? atoi(e) : 8;
  }();
  ScoreArgs a;
  if (ntiles <= small_tiles_max) {
    bool with_linf = false;
    VZ_TRY(prepare_small_score(h, Xs, Zs, M, acq, score, mu, sigma, linf, &a, &with_linf));
    const int nvb = h->np / kVarCols, nmb = h->np / 64;
    const size_t sm1 = cross_small_smem_bytes(h->dc, h->dk), sm2 = var_small_smem_bytes();
    const dim3 g1(nmb, ntiles * 4), g2(nvb, ntiles);
    if (with_linf) {
      VZ_CUDA(cudaFuncSetAttribute(k_cross_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
      k_cross_small<true><<<g1, kSmallThreads, sm1, h->stream>>>(a);
    } else {
      VZ_CUDA(cudaFuncSetAttribute(k_cross_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
      k_cross_small<false><<<g1, kSmallThreads, sm1, h->stream>>>(a);
    }
    VZ_CHECK_LAUNCH();
    VZ_CUDA(cudaFuncSetAttribute(k_var_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    k_var_small<<<g2, kSmallThreads, sm2, h->stream>>>(a);
    VZ_CHECK_LAUNCH();
    if (with_linf) k_small_finalize<true><<<(M + 31) / 32, 256, 0, h->stream>>>(a);
    else k_small_finalize<false><<<(M + 31) / 32, 256, 0, h->stream>>>(a);
    VZ_CHECK_LAUNCH();
    h->launches += 3;
    return 0;
  }
  {
    static const int env_i8 = [] { const char* e = getenv("VZGP_SCORE_I8"); return e ? atoi(e) : 1; }();
    const int want = h->score_i8 >= 0 ? h->score_i8 : env_i8;
    if (want && score_i8_eligible(h, M)) return launch_score_i8(h, Xs, Zs, M, acq, score, mu, sigma, linf);
  }
  // Medium pools cannot fill the GPU with one CTA per tile: share each tile's output column
  // blocks between nsplit CTAs (each recomputes the cheap K* tile).
  int nsplit = 1;
  if (ntiles * 2 <= h->sm_count && nblocks >= 2) nsplit = (nblocks + 1) / 2;
  const int nwork = ntiles * nsplit;
  const int grid = nwork < h->sm_count ? nwork : h->sm_count;
  VZ_TRY(ensure_scratch(h, (size_t)grid * kTM * h->np * sizeof(double)));
  fill_score_args(h, Xs, Zs, M, acq, score, mu, sigma, linf, &a);
  VZ_TRY(make_map(&a.mapA, a.scratch, (uint64_t)grid * kTM, (uint64_t)h->np, (uint64_t)h->np, kTM));
  VZ_TRY(make_map(&a.mapB, a.Linv, (uint64_t)h->np, (uint64_t)h->np, (uint64_t)h->np, kBN));
  a.nsplit = nsplit;
  if (nsplit > 1) {
    VZ_TRY(h->Tws.reserve(sizeof(double) * (size_t)(nsplit + 2) * a.mpad));
    a.part = h->Tws.as<double>();
  }
  const bool need_linf = (linf != nullptr) || (a.apply_tr && a.radius <= 0.5);
  const size_t sm = score_smem_bytes(h->dc, h->dk, need_linf);
  if (sm > 227 * 1024) {
    set_error("score kernel needs %zu bytes of shared memory (Dc=%d with trust-region distance)", sm, h->dc);
    return VZGP_ERR_UNSUPPORTED;
  }
  if (need_linf) {
    VZ_CUDA(cudaFuncSetAttribute(k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_score<true><<<grid, kBlockThreads, sm, h->stream>>>(a);
  } else {
    VZ_CUDA(cudaFuncSetAttribute(k_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_score<false><<<grid, kBlockThreads, sm, h->stream>>>(a);
  }
  VZ_CHECK_LAUNCH();
  h->launches++;
  if (nsplit > 1) {
    k_score_finalize<<<(M + 255) / 256, 256, 0, h->stream>>>(a);
    VZ_CHECK_LAUNCH();
    h->launches++;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// GP-UCB-PE acquisition from the pieces of two models (gp_ucb_pe.py:344-381, :434-492, :221-242).
// ---------------------------------------------------------------------------
struct PeCombine {
  int mode;
  double ucb, explore, penalty, threshold;
  int apply_tr;
  double radius;
};
__global__ void k_pe_combine(int M, PeCombine p, const double* __restrict__ mu, const double* __restrict__ sd,
                             const double* __restrict__ sd_all, const double* __restrict__ linf,
                             double* __restrict__ score) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double acq;
  if (p.mode == 0) {
    acq = fma(p.ucb, sd_all[m], mu[m]);
  } else {
    const double explore_ucb = fma(sd[m], p.explore, mu[m]);
    acq = sd_all[m] + p.penalty * fmin(explore_ucb - p.threshold, 0.0);
  }
  if (p.apply_tr) {
    const double dist = linf[m];
    const bool inside = (dist < p.radius) || (p.radius > 0.5);
    acq = inside ? acq : (-1e4 - dist);
  }
  score[m] = acq;
}

int launch_score_pe(vzgp_handle* hA, vzgp_handle* hB, const double* Xs, const int32_t* Zs, int M,
                    const vzgp_pe_params* pe, double* score, double* mu, double* sigma, double* sigma_all) {
  if (M <= 0) return 0;
  VZ_TRY(hA->pe_tmp.reserve(sizeof(double) * 6 * (size_t)M));
  double* t = hA->pe_tmp.as<double>();
  double* mu_a = mu ? mu : t;
  double* sd_a = sigma ? sigma : t + M;
  double* sd_b = sigma_all ? sigma_all : t + 2 * (size_t)M;
  double* linf_b = t + 3 * (size_t)M;
  double* dummy_a = t + 4 * (size_t)M;
  double* dummy_b = t + 5 * (size_t)M;
  vzgp_acq none;
  none.ucb_coefficient = 0.0; none.use_trust_region = 0; none.trust_radius = 1.0; none.tr_dim_mask = nullptr;
  none.tr_rows = 0; none.tr_strict = 0;
  VZ_TRY(launch_score(hA, Xs, Zs, M, &none, dummy_a, mu_a, sd_a, nullptr));
  vzgp_acq accb = none;
  accb.tr_dim_mask = pe->tr_dim_mask;
  accb.tr_rows = pe->tr_rows;
  const bool want_tr = pe->use_trust_region && pe->trust_radius <= 0.5;
  VZ_TRY(launch_score(hB, Xs, Zs, M, &accb, dummy_b, nullptr, sd_b, want_tr ? linf_b : nullptr));
  PeCombine p;
  p.mode = pe->mode; p.ucb = pe->ucb_coefficient; p.explore = pe->explore_coefficient;
  p.penalty = pe->penalty_coefficient; p.threshold = pe->threshold;
  p.apply_tr = want_tr ? 1 : 0; p.radius = pe->trust_radius;
  k_pe_combine<<<(M + 255) / 256, 256, 0, hA->stream>>>(M, p, mu_a, sd_a, sd_b, linf_b, score);
  VZ_CHECK_LAUNCH();
  hA->launches++;
  return 0;
}

// ---------------------------------------------------------------------------
// Stacked residual GPs of transfer learning (StackedResidualGP, gp/gp_models.py:91-140; combine_predictions_with_aux,
// gp/transfer_learning.py:62-152): level 0 is the first prior study's GP, every further level is trained on the
// residuals of the stack below it, the last level on the current study.  mean = sum of the level means; the stddevs
// are combined bottom-up by weighted geometric means  s <- sd_e^alpha_e * s^(1 - alpha_e)  (alpha_e from the degrees
// of freedom of the two levels, computed by the caller).  UCB and the trust region (distances to the TOP level's
// trials) on the combined prediction.
// ---------------------------------------------------------------------------
struct StackCombine {
  int E;
  double alpha[16];
  double coef;
  int apply_tr, tr_strict;
  double radius;
};
__global__ void k_stack_combine(int M, StackCombine p, const double* __restrict__ mu_e, const double* __restrict__ sd_e,
                                const double* __restrict__ linf, double* __restrict__ score, double* __restrict__ mu,
                                double* __restrict__ sigma) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double mean = mu_e[m], sd = sd_e[m];
  for (int e = 1; e < p.E; ++e) {
    mean += mu_e[(size_t)e * M + m];
    sd = pow(sd_e[(size_t)e * M + m], p.alpha[e]) * pow(sd, 1.0 - p.alpha[e]);
  }
  double sc = fma(p.coef, sd, mean);
  if (p.apply_tr) {
    const double dist = linf[m];
    const bool inside = (p.tr_strict ? (dist < p.radius) : (dist <= p.radius)) || (p.radius > 0.5);
    sc = inside ? sc : (-1e4 - dist);
  }
  score[m] = sc;
  if (mu) mu[m] = mean;
  if (sigma) sigma[m] = sd;
}

int launch_score_stack(vzgp_handle* const* hs, int E, const double* alphas, const double* Xs, const int32_t* Zs, int M,
                       const vzgp_acq* acq, double* score, double* mu, double* sigma, double* linf) {
  if (M <= 0) return 0;
  vzgp_handle* top = hs[E - 1];
  VZ_TRY(top->pe_tmp.reserve(sizeof(double) * (2 * (size_t)E + 2) * (size_t)M));
  double* t = top->pe_tmp.as<double>();
  double* mu_e = t;
  double* sd_e = t + (size_t)E * M;
  double* linf_buf = linf ? linf : t + 2 * (size_t)E * M;
  double* dummy = t + (2 * (size_t)E + 1) * M;
  const bool want_tr = acq->use_trust_region && acq->trust_radius <= 0.5;
  const bool want_linf = want_tr || linf != nullptr;
  vzgp_acq none;
  none.ucb_coefficient = 0.0; none.use_trust_region = 0; none.trust_radius = 1.0;
  none.tr_dim_mask = acq->tr_dim_mask; none.tr_rows = acq->tr_rows; none.tr_strict = 0;
  for (int e = 0; e < E; ++e)   // the trust region is measured against the trials of the top level (the current study)
    VZ_TRY(launch_score(hs[e], Xs, Zs, M, &none, dummy, mu_e + (size_t)e * M, sd_e + (size_t)e * M,
                        (e == E - 1 && want_linf) ? linf_buf : nullptr));
  StackCombine p;
  p.E = E; p.coef = acq->ucb_coefficient; p.apply_tr = want_tr ? 1 : 0; p.tr_strict = acq->tr_strict ? 1 : 0;
  p.radius = acq->trust_radius;
  for (int e = 0; e < 16; ++e) p.alpha[e] = e < E ? alphas[e] : 0.0;
  k_stack_combine<<<(M + 255) / 256, 256, 0, top->stream>>>(M, p, mu_e, sd_e, linf_buf, score, mu, sigma);
  VZ_CHECK_LAUNCH();
  top->launches++;
  return 0;
}

// ---------------------------------------------------------------------------
// Set-PE acquisition (SetPEScoreFunction, gp_ucb_pe.py:510-594): per set of q points
//   logdet(joint predictive covariance under model B)  +  penalty * sum_i min(mean_A + explore * stddev_A - threshold, 0)
//   [+ sum_i (dist_i > radius and radius <= 0.5) * (-1e4 - dist_i)      _apply_trust_region_to_set, :245-269]
// One warp per set: the q x q block of the [M x M] covariance is factored in shared memory (q <= 16); a pivot that is
// not positive gives -inf like the reference's NaN -> -inf rule (:495-507).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_set_pe_combine(int n_sets, int q, PeCombine p, const double* __restrict__ cov, int ldc,
                                                       const double* __restrict__ mu_a, const double* __restrict__ sd_a,
                                                       const double* __restrict__ linf, double* __restrict__ score,
                                                       double* __restrict__ sd_all) {
  __shared__ double c[16][17];
  const int s = blockIdx.x, lane = threadIdx.x;
  if (s >= n_sets) return;
  const int r0 = s * q;
  for (int e = lane; e < q * q; e += 32) c[e / q][e % q] = cov[(size_t)(r0 + e / q) * ldc + r0 + e % q];
  __syncwarp();
  if (sd_all && lane < q) sd_all[r0 + lane] = sqrt(fmax(c[lane][lane], 0.0));
  double logdet = 0.0;
  bool bad = false;
  for (int k = 0; k < q; ++k) {
    const double d = c[k][k];
    if (!(d > 0.0) || !isfinite(d)) { bad = true; break; }
    logdet += log(d);
    __syncwarp();
    const double inv = 1.0 / d;
    // right-looking LDL^T step on the trailing block: lane = row i > k
    for (int i = k + 1 + lane; i < q; i += 32) {
      const double lik = c[i][k] * inv;
      for (int j = k + 1; j <= i; ++j) c[i][j] -= lik * c[j][k];
    }
    __syncwarp();
    // keep the block symmetric for the next pivot column reads (c[j][k] with j > k is the lower part: already there)
  }
  double pen = 0.0, tr = 0.0;
  for (int i = lane; i < q; i += 32) {
    pen += fmin(mu_a[r0 + i] + p.explore * sd_a[r0 + i] - p.threshold, 0.0);
    if (p.apply_tr) {
      const double dist = linf[r0 + i];
      if (dist > p.radius) tr += -1e4 - dist;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pen += __shfl_xor_sync(0xffffffffu, pen, o);
    tr += __shfl_xor_sync(0xffffffffu, tr, o);
  }
  if (lane == 0) score[s] = (bad ? -INFINITY : logdet) + p.penalty * pen + tr;
}

int launch_set_pe_combine(vzgp_handle* h, int n_sets, int q, const vzgp_pe_params* pe, const double* cov, int ldc,
                          const double* mu_a, const double* sd_a, const double* linf, double* score, double* sd_all) {
  PeCombine p;
  p.mode = 1; p.ucb = pe->ucb_coefficient; p.explore = pe->explore_coefficient;
  p.penalty = pe->penalty_coefficient; p.threshold = pe->threshold;
  p.apply_tr = (pe->use_trust_region && pe->trust_radius <= 0.5 && linf != nullptr) ? 1 : 0;
  p.radius = pe->trust_radius;
  k_set_pe_combine<<<n_sets, 32, 0, h->stream>>>(n_sets, q, p, cov, ldc, mu_a, sd_a, linf, score, sd_all);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}

// ---------------------------------------------------------------------------
// Uniform ensemble of E models (UniformEnsemblePredictive, stochastic_process_model.py:846-868:
// equal-weight MixtureSameFamily): mean = avg mu_e, var = avg(sd_e^2 + mu_e^2) - mean^2.
// ---------------------------------------------------------------------------
struct EnsCombine {
  int E;
  double coef;
  int apply_tr, tr_strict;
  double radius;
};
__global__ void k_ensemble_combine(int M, EnsCombine p, const double* __restrict__ mu_e, const double* __restrict__ sd_e,
                                   const double* __restrict__ linf, double* __restrict__ score,
                                   double* __restrict__ mu, double* __restrict__ sigma, int* __restrict__ clamp_count) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double s1 = 0.0, s2 = 0.0;
  for (int e = 0; e < p.E; ++e) {
    const double a = mu_e[(size_t)e * M + m], b = sd_e[(size_t)e * M + m];
    s1 += a;
    s2 += fma(b, b, a * a);
  }
  const double mean = s1 / p.E;
  double var = s2 / p.E - mean * mean;
  if (var < 0.0) { var = 0.0; atomicAdd(clamp_count, 1); }
  const double sd = sqrt(var);
  double sc = fma(p.coef, sd, mean);
  if (p.apply_tr) {
    const double dist = linf[m];
    const bool inside = (p.tr_strict ? (dist < p.radius) : (dist <= p.radius)) || (p.radius > 0.5);
    sc = inside ? sc : (-1e4 - dist);
  }
  score[m] = sc;
  if (mu) mu[m] = mean;
  if (sigma) sigma[m] = sd;
}

int launch_score_ensemble(vzgp_handle* const* hs, int E, const double* Xs, const int32_t* Zs, int M,
                          const vzgp_acq* acq, double* score, double* mu, double* sigma, double* linf) {
  if (M <= 0) return 0;
  vzgp_handle* h0 = hs[0];
  VZ_TRY(h0->pe_tmp.reserve(sizeof(double) * (2 * (size_t)E + 2) * (size_t)M));
  double* t = h0->pe_tmp.as<double>();
  double* mu_e = t;
  double* sd_e = t + (size_t)E * M;
  double* linf_buf = linf ? linf : t + 2 * (size_t)E * M;
  double* dummy = t + (2 * (size_t)E + 1) * M;
  const bool want_tr = acq->use_trust_region && acq->trust_radius <= 0.5;
  const bool want_linf = want_tr || linf != nullptr;
  vzgp_acq none;
  none.ucb_coefficient = 0.0; none.use_trust_region = 0; none.trust_radius = 1.0;
  none.tr_dim_mask = acq->tr_dim_mask; none.tr_rows = acq->tr_rows; none.tr_strict = 0;
  for (int e = 0; e < E; ++e)   // all members share the trials: the distance is computed once
    VZ_TRY(launch_score(hs[e], Xs, Zs, M, &none, dummy, mu_e + (size_t)e * M, sd_e + (size_t)e * M,
                        (e == 0 && want_linf) ? linf_buf : nullptr));
  EnsCombine p;
  p.E = E; p.coef = acq->ucb_coefficient; p.apply_tr = want_tr ? 1 : 0; p.tr_strict = acq->tr_strict ? 1 : 0;
  p.radius = acq->trust_radius;
  k_ensemble_combine<<<(M + 255) / 256, 256, 0, h0->stream>>>(M, p, mu_e, sd_e, linf_buf, score, mu, sigma, h0->small.as<int>());
  VZ_CHECK_LAUNCH();
  h0->launches++;
  return 0;
}

// ---------------------------------------------------------------------------
// Philox candidate pool: X[m, d] = U(seed, STREAM_RANDOM_POOL, 0, (index_base+m)*dc + d)
// ---------------------------------------------------------------------------
__global__ void k_random_pool(double* __restrict__ X, int64_t total, int64_t elem_base,
                              uint64_t seed, uint32_t stream, uint32_t iteration) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < total; e += stride)
    X[e] = philox_uniform(seed, stream, iteration, (uint64_t)(elem_base + e));
}

int launch_random_fill(vzgp_handle* h, double* X, int64_t total, int64_t elem_base, uint64_t seed,
                       uint32_t stream, uint32_t iteration) {
  if (total <= 0) return 0;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
  k_random_pool<<<(unsigned)blocks, 256, 0, h->stream>>>(X, total, elem_base, seed, stream, iteration);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}

// ---------------------------------------------------------------------------
// Top-k by repeated arg-max (count is small: the number of suggestions).  Ordering: larger
// score first, ties -> lower index, NaN -> -inf.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
  return (v > bv) || (v == bv && i < bi);
}

// partial[b] = best of this block's slice, ignoring indices already in taken[0..ntaken)
__global__ void k_argmax_partial(const double* __restrict__ s, int64_t M,
                                 const long long* __restrict__ taken, int ntaken,
                                 ArgMax* __restrict__ partial) {
  __shared__ double sv[256];
  __shared__ long long si[256];
  double bv = -INFINITY;
  long long bi = LLONG_MAX;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M;
       e += (int64_t)gridDim.x * blockDim.x) {
    double v = s[e];
    if (isnan(v)) v = -INFINITY;
    bool skip = false;
    for (int t = 0; t < ntaken; ++t) skip |= (taken[t] == e);
    if (!skip && better(v, e, bv, bi)) { bv = v; bi = e; }
  }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o && better(sv[threadIdx.x + o], si[threadIdx.x + o], sv[threadIdx.x], si[threadIdx.x])) {
      sv[threadIdx.x] = sv[threadIdx.x + o];
      si[threadIdx.x] = si[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[blockIdx.x].v = sv[0]; partial[blockIdx.x].i = si[0]; }
}

__global__ void k_argmax_final(const ArgMax* __restrict__ partial, int nb, long long* __restrict__ taken,
                               double* __restrict__ vals, int slot) {
  __shared__ double sv[256];
  __shared__ long long si[256];
  double bv = -INFINITY;
  long long bi = LLONG_MAX;
  for (int e = threadIdx.x; e < nb; e += blockDim.x)
    if (better(partial[e].v, partial[e].i, bv, bi)) { bv = partial[e].v; bi = partial[e].i; }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o && better(sv[threadIdx.x + o], si[threadIdx.x + o], sv[threadIdx.x], si[threadIdx.x])) {
      sv[threadIdx.x] = sv[threadIdx.x + o];
      si[threadIdx.x] = si[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { taken[slot] = si[0]; vals[slot] = sv[0]; }
}

// Device-side top-k: results land in d_idx[count], d_val[count] (device).
int launch_topk_device(vzgp_handle* h, const double* score, int64_t M, int count, long long* d_idx,
                       double* d_val, ArgMax* d_partial, int nblocks) {
  for (int c = 0; c < count; ++c) {
    k_argmax_partial<<<nblocks, 256, 0, h->stream>>>(score, M, d_idx, c, d_partial);
    VZ_CHECK_LAUNCH();
    k_argmax_final<<<1, 256, 0, h->stream>>>(d_partial, nblocks, d_idx, d_val, c);
    VZ_CHECK_LAUNCH();
    h->launches += 2;
  }
  return 0;
}

// Gather rows: out[c, :] = X[idx[c], :]  (idx may be LLONG_MAX when fewer than count exist)
__global__ void k_gather_rows(const double* __restrict__ X, int dc, const long long* __restrict__ idx,
                              int count, int64_t M, double* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count * dc) return;
  int c = e / dc, d = e % dc;
  long long i = idx[c];
  out[e] = (i >= 0 && i < M) ? X[(size_t)i * dc + d] : 0.0;
}

int launch_gather_rows(vzgp_handle* h, const double* X, int dc, const long long* idx, int count,
                       int64_t M, double* out) {
  k_gather_rows<<<(count * dc + 255) / 256, 256, 0, h->stream>>>(X, dc, idx, count, M, out);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}

// Pack the local winners as rows [score, global index, Dc features] (fp64; indices < 2^53 exact).
__global__ void k_pack_topk(const double* __restrict__ X, int dc, const long long* __restrict__ idx,
                            const double* __restrict__ val, int count, int64_t M, int64_t index_base,
                            double* __restrict__ payload) {
  const int w = dc + 2;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count * w) return;
  const int c = e / w, d = e % w;
  const long long i = idx[c];
  const bool ok = (i >= 0 && i < M);
  double v;
  if (d == 0) v = ok ? val[c] : -INFINITY;
  else if (d == 1) v = ok ? (double)(index_base + i) : -1.0;
  else v = ok ? X[(size_t)i * dc + (d - 2)] : 0.0;
  payload[e] = v;
}

int launch_pack_topk(vzgp_handle* h, const double* X, int dc, const long long* idx, const double* val,
                     int count, int64_t M, int64_t index_base, double* payload) {
  const int n = count * (dc + 2);
  k_pack_topk<<<(n + 255) / 256, 256, 0, h->stream>>>(X, dc, idx, val, count, M, index_base, payload);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}

// Merge gathered winner rows: merge_topk_block (topk_merge.cuh), one CTA.
__global__ void __launch_bounds__(256) k_merge_topk(const double* __restrict__ rows, int n_rows, int width,
                                                    int count, double* __restrict__ out) {
  merge_topk_block<false>(rows, n_rows, width, count, out);
}

int launch_merge_topk(vzgp_handle* h, const double* rows, int n_rows, int width, int count, double* out) {
  k_merge_topk<<<1, 256, 0, h->stream>>>(rows, n_rows, width, count, out);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}


}  // namespace vzgp
