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
// One persistent CTA per SM walks 128-candidate tiles:
//   phase 1  K* tile [128 x np] = Matern(x*, X) built 64 columns at a time from shared-memory
//            staged rows; mu = K* alpha and the L-inf trust-region distance are reduced on the
//            fly; the tile goes to a CTA-private scratch (L2 resident, never re-read by others).
//   phase 2  W = K* . Linv^T by 128x64 register-tiled fp64 GEMM blocks, exploiting that Linv is
//            lower triangular (k <= j); each block is squared and row-summed in registers, W is
//            never stored.
//   epilogue var = sf2 + sn2 - sum W^2 (clamped at 0), sigma, UCB, trust region, outputs.
#include <climits>

#include "launchers.h"
#include "tiles.cuh"

namespace vzgp {

using GS = GemmCfg<128, 64, 16, 8, 4>;

struct ScoreArgs {
  const double* Xs;
  const int32_t* Zs;
  int M;
  const double* X;
  const int32_t* Z;
  int np;
  int n_valid;
  const double* Linv;
  int ldi;
  const double* alpha;
  KernelParams kp;
  double sn2;
  double coef;
  int apply_tr;     // trust region modifies the score
  double radius;
  uint8_t tr_mask[kMaxDc];
  double* scratch;  // [gridDim.x][128][np]
  double* score;
  double* mu;
  double* sigma;
  double* linf;
  int* clamp_count;
};

template <bool WITH_LINF>
__global__ void __launch_bounds__(256, 1) k_score(const ScoreArgs a) {
  extern __shared__ double smem[];
  constexpr int LDA = 130, LDB = 66;
  const int dc = a.kp.dc, dk = a.kp.dk, np = a.np;
  double* gemm_smem = smem;                              // GS::kSmemDoubles
  double* sa = gemm_smem + GS::kSmemDoubles;             // [dc][LDA]
  double* sb = sa + dc * LDA;                            // [dc][LDB]
  double* s_alpha = sb + dc * LDB;                       // [64]
  double* s_mu = s_alpha + 64;                           // [128]
  double* s_linf = s_mu + 128;                           // [128]
  int32_t* za = reinterpret_cast<int32_t*>(s_linf + 128);  // [dk][LDA]
  int32_t* zb = za + dk * LDA;                           // [dk][LDB]
  uint8_t* s_mask = reinterpret_cast<uint8_t*>(zb + dk * LDB);  // [kMaxDc]

  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  double* scr = a.scratch + (size_t)blockIdx.x * 128 * np;
  if (tid < kMaxDc) s_mask[tid] = a.tr_mask[tid];
  int clamped = 0;

  const int ntiles = (a.M + 127) / 128;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * 128;
    __syncthreads();  // previous tile's readers of sa / s_mu are done
    stage_rows_T(a.Xs, a.M, dc, m0, 128, sa, LDA);
    if (dk > 0) stage_rows_T_i32(a.Zs, a.M, dk, m0, 128, za, LDA);

    // ---------------- phase 1: K* tile, mean, trust-region distance ----------------
    double mu_part[8], lmin[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mu_part[i] = 0.0; lmin[i] = INFINITY; }
    for (int jb = 0; jb < np / 64; ++jb) {
      __syncthreads();  // sb / s_alpha free (and sa staged on the first pass)
      stage_rows_T(a.X, np, dc, jb * 64, 64, sb, LDB);
      if (dk > 0) stage_rows_T_i32(a.Z, np, dk, jb * 64, 64, zb, LDB);
      if (tid < 64) s_alpha[tid] = a.alpha[jb * 64 + tid];
      __syncthreads();
      double d2[8][4], lf[8][4];
      tile_d2<GS, 8, 4, WITH_LINF>(sa, LDA, sb, LDB, za, LDA, zb, LDB, a.kp, s_mask, ty, tx, d2,
                                   lf);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double kv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cj = GS::col_of(tx, j);
          const bool valid = (jb * 64 + cj) < a.n_valid;
          kv[j] = valid ? matern52(d2[i][j], a.kp.sf2) : 0.0;
          mu_part[i] = fma(kv[j], s_alpha[cj], mu_part[i]);
          if (WITH_LINF && valid) lmin[i] = fmin(lmin[i], lf[i][j]);
        }
        double* dst = scr + (size_t)GS::row_of(ty, i) * np + jb * 64;
        *reinterpret_cast<double2*>(dst + GS::col_of(tx, 0)) = make_double2(kv[0], kv[1]);
        *reinterpret_cast<double2*>(dst + GS::col_of(tx, 2)) = make_double2(kv[2], kv[3]);
      }
    }
    // reduce over the 16 lanes (tx) that share the same rows
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        mu_part[i] += __shfl_xor_sync(0xffffffffu, mu_part[i], o);
        if (WITH_LINF) lmin[i] = fmin(lmin[i], __shfl_xor_sync(0xffffffffu, lmin[i], o));
      }
      if (tx == 0) {
        s_mu[GS::row_of(ty, i)] = mu_part[i];
        s_linf[GS::row_of(ty, i)] = lmin[i];
      }
    }
    __syncthreads();  // scratch tile written by this CTA is visible to all its threads

    // ---------------- phase 2: row sums of (K* Linv^T)^2 ----------------
    double rowsq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rowsq[i] = 0.0;
    for (int jb = 0; jb < np / 64; ++jb) {
      double acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
      gemm_mainloop<128, 64, 16, 8, 4, false, false>(scr, np, 0, a.Linv, a.ldi, jb * 64, 0,
                                                     (jb + 1) * 64, acc, gemm_smem);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) rowsq[i] = fma(acc[i][j], acc[i][j], rowsq[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rowsq[i] += __shfl_xor_sync(0xffffffffu, rowsq[i], o);
    }
    // ---------------- epilogue ----------------
    if (tx == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = GS::row_of(ty, i);
        const int m = m0 + r;
        if (m >= a.M) continue;
        double var = a.kp.sf2 - rowsq[i] + a.sn2;
        if (var < 0.0) { var = 0.0; ++clamped; }
        const double sd = sqrt(var);
        const double mean = s_mu[r];
        double sc = fma(a.coef, sd, mean);
        const double dist = s_linf[r];
        if (a.apply_tr) {
          const bool inside = (dist <= a.radius) || (a.radius > 0.5);
          sc = inside ? sc : (-1e4 - dist);
        }
        a.score[m] = sc;
        if (a.mu) a.mu[m] = mean;
        if (a.sigma) a.sigma[m] = sd;
        if (a.linf) a.linf[m] = dist;
      }
    }
  }
  if (clamped) atomicAdd(a.clamp_count, clamped);
}

size_t score_smem_bytes(int dc, int dk) {
  return sizeof(double) * (GS::kSmemDoubles + dc * (130 + 66) + 64 + 128 + 128) +
         sizeof(int32_t) * dk * (130 + 66) + kMaxDc;
}

int launch_score(vzgp_handle* h, const double* Xs, const int32_t* Zs, int M, const vzgp_acq* acq,
                 double* score, double* mu, double* sigma, double* linf) {
  if (M <= 0) return 0;
  const int ntiles = (M + 127) / 128;
  const int grid = ntiles < h->sm_count ? ntiles : h->sm_count;
  VZ_TRY(h->scratch.reserve((size_t)grid * 128 * h->np * sizeof(double)));
  VZ_TRY(h->small.reserve(4096));
  ScoreArgs a;
  a.Xs = Xs; a.Zs = Zs; a.M = M;
  a.X = h->X.as<double>(); a.Z = h->Z.as<int32_t>();
  a.np = h->np; a.n_valid = h->n_valid;
  a.Linv = h->Linv.as<double>(); a.ldi = h->np;
  a.alpha = h->alpha.as<double>();
  a.kp = h->kp; a.sn2 = h->sn2;
  a.coef = acq->ucb_coefficient;
  a.apply_tr = acq->use_trust_region ? 1 : 0;
  a.radius = acq->trust_radius;
  for (int d = 0; d < kMaxDc; ++d)
    a.tr_mask[d] = (d < h->dc) ? (acq->tr_dim_mask ? (acq->tr_dim_mask[d] ? 1 : 0) : 1) : 0;
  a.scratch = h->scratch.as<double>();
  a.score = score; a.mu = mu; a.sigma = sigma; a.linf = linf;
  a.clamp_count = h->small.as<int>();  // slot 0
  const bool need_linf = (linf != nullptr) || (a.apply_tr && a.radius <= 0.5);
  const size_t sm = score_smem_bytes(h->dc, h->dk);
  if (need_linf) {
    VZ_CUDA(cudaFuncSetAttribute(k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_score<true><<<grid, 256, sm, h->stream>>>(a);
  } else {
    VZ_CUDA(cudaFuncSetAttribute(k_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_score<false><<<grid, 256, sm, h->stream>>>(a);
  }
  VZ_CHECK_LAUNCH();
  h->launches++;
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

}  // namespace vzgp
