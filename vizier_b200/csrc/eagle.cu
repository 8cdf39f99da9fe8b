// Vectorised Eagle/Firefly acquisition optimiser: device-resident population state and the
// suggest / update / trim / top-k steps (continuous features, n_parallel == 1).
//
// Replaces (reference): VectorizedEagleStrategy.init_state/_populate_pool_with_prior_trials
// (vizier/_src/algorithms/optimizers/eagle_strategy.py:527-713), suggest/_create_features/
// _create_random_perturbations (:720-952, :1013-1073), update/_update_pool_features_and_rewards/
// _trim_pool (:1075-1247) and VectorizedOptimizer._update_best_results
// (vizier/_src/algorithms/optimizers/vectorized_base.py:544-587).
// Randomness is Philox4x32-10 (see device.cuh); with n_parallel == 1 the reference's normalised
// Laplace perturbation is exactly +-1 per coordinate (eagle_strategy.py:1033-1044).
#include <climits>

#include "device.cuh"
#include "launchers.h"

namespace vzgp {


// ---------------------------------------------------------------------------
// init: pool <- Philox uniforms; rewards=-inf; perturbations=cfg.perturbation; best=-inf.
// ---------------------------------------------------------------------------
__global__ void k_eagle_init(EagleDev e) {
  const int64_t total = (int64_t)e.P * e.D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    e.pool[i] = philox_uniform(e.seed, kStreamInitPool, 0, (uint64_t)i);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.P; i += gridDim.x * blockDim.x) {
    e.rewards[i] = -INFINITY;
    e.pert[i] = e.cfg.perturbation;
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < e.count; c += blockDim.x) {
      e.best_r[c] = -INFINITY;
      e.best_id[c] = LLONG_MAX;
    }
    for (int c = threadIdx.x; c < e.count * e.D; c += blockDim.x) e.best_x[c] = 0.0;
    if (threadIdx.x == 0) { *e.best_reward = -INFINITY; *e.iter = 0; }
  }
}

// ---------------------------------------------------------------------------
// Prior-trial seeding (single CTA; the reference loop is sequential too).
// prior [n x D] in creation order, prior_r [n] their acquisition values.
// ord [n] int workspace, chosen_r [left] workspace.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_eagle_seed_priors(EagleDev e, const double* __restrict__ prior,
                                                           const double* __restrict__ prior_r,
                                                           int n, int* __restrict__ ord,
                                                           double* __restrict__ chosen_r) {
  __shared__ double sv[256];
  __shared__ int si[256];
  const int tid = threadIdx.x, D = e.D;
  const int n_random = (int)(e.P * (1.0 - e.cfg.prior_trials_pool_pct));
  const int left = e.P - n_random;
  // _mask_flip: valid entries newest first, then the -inf ones (eagle_strategy.py:472-496).
  if (tid == 0) {
    int w = 0;
    for (int i = n - 1; i >= 0; --i)
      if (!(isinf(prior_r[i]) && prior_r[i] < 0)) ord[w++] = i;
    for (int i = n - 1; i >= 0; --i)
      if (isinf(prior_r[i]) && prior_r[i] < 0) ord[w++] = i;
  }
  __syncthreads();
  const int chosen = n < left ? n : left;
  double* feat = e.pool + (size_t)n_random * D;  // chosen set lives directly in the pool
  // Save the random rows that chosen entries may have to fall back to: they are simply the
  // current pool contents, so only overwrite when the chosen reward is finite (done at the end).
  // Work on a staging copy in tmp area = e.batch is too small in general, so keep chosen features
  // in the pool and remember which stay random via chosen_r == -inf.
  for (int c = tid; c < chosen; c += 256) chosen_r[c] = prior_r[ord[c]];
  __syncthreads();
  // Stage chosen features into the pool rows, but keep the original random rows for -inf ones.
  for (int idx = tid; idx < chosen * D; idx += 256) {
    int c = idx / D, d = idx % D;
    if (!(isinf(chosen_r[c]) && chosen_r[c] < 0)) feat[(size_t)c * D + d] = prior[(size_t)ord[c] * D + d];
  }
  __syncthreads();
  for (int i = left; i < n; ++i) {
    const double* x = prior + (size_t)ord[i] * D;
    double bv = INFINITY;
    int bi = INT_MAX;
    for (int c = tid; c < left; c += 256) {
      // distance to chosen member c; members with -inf reward hold random rows in the pool but
      // the reference compares against the *prior* feature there; this only happens for padded
      // priors, which sort last and never enter the chosen set before valid ones run out.
      double s = 0.0;
      for (int d = 0; d < D; ++d) {
        double df = x[d] - feat[(size_t)c * D + d];
        s = fma(df, df, s);
      }
      if (s < bv || (s == bv && c < bi)) { bv = s; bi = c; }
    }
    sv[tid] = bv; si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) {
        if (sv[tid + o] < sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) {
          sv[tid] = sv[tid + o]; si[tid] = si[tid + o];
        }
      }
      __syncthreads();
    }
    const int ind = si[0];
    const double ri = prior_r[ord[i]];
    const bool repl = chosen_r[ind] < ri;
    __syncthreads();
    if (repl) {
      for (int d = tid; d < D; d += 256) feat[(size_t)ind * D + d] = x[d];
      if (tid == 0) chosen_r[ind] = ri;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// suggest: one warp per batch fly, 8 flies per CTA.  Dynamic smem: 8*P doubles (forces) +
// 8*D doubles (the flies' own features).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_eagle_suggest(EagleDev e) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int P = e.P, B = e.B, D = e.D;
  const int t = *e.iter;
  const int nb = P / B;
  const int start = (t % nb) * B;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  const int i = start + b;
  double* s_f = smem + (size_t)warp * P;
  double* s_x = smem + (size_t)8 * P + warp * D;
  for (int d = lane; d < D; d += 32) s_x[d] = e.pool[(size_t)i * D + d];
  __syncwarp();
  double* out = e.batch + (size_t)b * D;
  if (t < nb) {  // still initialising: return the pool features (projected)
    for (int d = lane; d < D; d += 32) out[d] = fmin(fmax(s_x[d], 0.0), 1.0);
    return;
  }
  const double ri = e.rewards[i];
  const double cexp = -e.cfg.visibility / (double)D * 10.0;
  int npull = 0, npush = 0;
  for (int j = lane; j < P; j += 32) {
    const double* pj = e.pool + (size_t)j * D;
    double d2 = 0.0;
    for (int d = 0; d < D; ++d) {
      double df = s_x[d] - pj[d];
      d2 = fma(df, df, d2);
    }
    const double rj = e.rewards[j];
    const double dir = rj - ri;
    const double sd = (dir >= 0.0) ? e.cfg.gravity : -e.cfg.negative_gravity;
    const double f = sd * exp(cexp * d2) * (isfinite(rj) ? 1.0 : 0.0);
    s_f[j] = f;
    npull += (f > 0.0);
    npush += (f < 0.0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    npull += __shfl_xor_sync(0xffffffffu, npull, o);
    npush += __shfl_xor_sync(0xffffffffu, npush, o);
  }
  __syncwarp();
  const double wpull = npull > 0 ? e.cfg.normalization_scale / (double)npull : 0.0;
  const double wpush = npush > 0 ? e.cfg.normalization_scale / (double)npush : 0.0;
  // lane handles dims lane and lane+32
  double acc0 = 0.0, acc1 = 0.0, ssum = 0.0;
  const int d0 = lane, d1 = lane + 32;
  for (int j = 0; j < P; ++j) {
    const double f = s_f[j];
    const double sc = f > 0.0 ? f * wpull : (f < 0.0 ? f * wpush : 0.0);
    ssum += sc;
    const double* pj = e.pool + (size_t)j * D;
    if (d0 < D) acc0 = fma(sc, pj[d0], acc0);
    if (d1 < D) acc1 = fma(sc, pj[d1], acc1);
  }
  const double pert = e.pert[i];
  if (d0 < D) {
    double u = philox_uniform(e.seed, kStreamPerturbSign, (uint32_t)t, (uint64_t)b * D + d0);
    double v = s_x[d0] + (acc0 - s_x[d0] * ssum) + (u >= 0.5 ? pert : -pert);
    out[d0] = fmin(fmax(v, 0.0), 1.0);
  }
  if (d1 < D) {
    double u = philox_uniform(e.seed, kStreamPerturbSign, (uint32_t)t, (uint64_t)b * D + d1);
    double v = s_x[d1] + (acc1 - s_x[d1] * ssum) + (u >= 0.5 ? pert : -pert);
    out[d1] = fmin(fmax(v, 0.0), 1.0);
  }
}

// ---------------------------------------------------------------------------
// update + trim + top-count bookkeeping: single CTA.  Dynamic smem: (B+count) doubles + ints.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool rank_better(double v, long long id, double bv, long long bid) {
  return (v > bv) || (v == bv && id < bid);
}

__global__ void __launch_bounds__(256) k_eagle_update(EagleDev e) {
  extern __shared__ double smem[];
  __shared__ double sv[256];
  __shared__ long long si[256];
  __shared__ int sp[256];
  const int tid = threadIdx.x;
  const int P = e.P, B = e.B, D = e.D, count = e.count;
  const int t = *e.iter;
  const int nb = P / B;
  const int start = (t % nb) * B;
  // new best reward
  double m = -INFINITY;
  for (int b = tid; b < B; b += 256) m = fmax(m, e.batch_r[b]);
  sv[tid] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) sv[tid] = fmax(sv[tid], sv[tid + o]);
    __syncthreads();
  }
  const double new_best = fmax(*e.best_reward, sv[0]);
  __syncthreads();

  // ---- top-count merge of (batch U best) into tmp, then copy back ----
  double* cv = smem;                                        // [B+count] ranking values
  unsigned char* used = reinterpret_cast<unsigned char*>(cv + B + count);  // [B+count]
  for (int q = tid; q < B + count; q += 256) {
    double v = q < B ? e.batch_r[q] : e.best_r[q - B];
    cv[q] = isnan(v) ? -INFINITY : v;
    used[q] = 0;
  }
  __syncthreads();
  for (int c = 0; c < count; ++c) {
    double bv = -INFINITY;
    long long bid = LLONG_MAX;
    int bp = -1;
    for (int q = tid; q < B + count; q += 256) {
      if (used[q]) continue;
      long long id = q < B ? (long long)t * B + q : e.best_id[q - B];
      if (bp < 0 || rank_better(cv[q], id, bv, bid)) { bv = cv[q]; bid = id; bp = q; }
    }
    sv[tid] = bv; si[tid] = bid; sp[tid] = bp;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o && sp[tid + o] >= 0 &&
          (sp[tid] < 0 || rank_better(sv[tid + o], si[tid + o], sv[tid], si[tid]))) {
        sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; sp[tid] = sp[tid + o];
      }
      __syncthreads();
    }
    const int q = sp[0];
    const long long qid = si[0];
    __syncthreads();
    if (q >= 0) {
      const double* src = q < B ? e.batch + (size_t)q * D : e.best_x + (size_t)(q - B) * D;
      for (int d = tid; d < D; d += 256) e.tmp_x[(size_t)c * D + d] = src[d];
      if (tid == 0) {
        e.tmp_r[c] = q < B ? e.batch_r[q] : e.best_r[q - B];
        e.tmp_id[c] = qid;
        used[q] = 1;
      }
    }
    __syncthreads();
  }
  for (int q = tid; q < count * D; q += 256) e.best_x[q] = e.tmp_x[q];
  for (int c = tid; c < count; c += 256) { e.best_r[c] = e.tmp_r[c]; e.best_id[c] = e.tmp_id[c]; }

  // ---- pool update ----
  for (int b = tid; b < B; b += 256) {
    const int i = start + b;
    const double rb = e.batch_r[b];
    double pert = e.pert[i];
    if (t < nb) {
      for (int d = 0; d < D; ++d) e.pool[(size_t)i * D + d] = e.batch[(size_t)b * D + d];
      e.rewards[i] = rb;
    } else {
      const double prev = e.rewards[i];
      const bool improve = rb > prev;
      double nr = improve ? rb : prev;
      if (!improve) pert *= e.cfg.penalize_factor;
      const bool trim = (pert < e.cfg.perturbation_lower_bound) && (nr != new_best);
      if (trim) {
        for (int d = 0; d < D; ++d)
          e.pool[(size_t)i * D + d] =
              philox_uniform(e.seed, kStreamTrim, (uint32_t)t, (uint64_t)b * D + d);
        pert = e.cfg.perturbation;
        nr = -INFINITY;
      } else if (improve) {
        for (int d = 0; d < D; ++d) e.pool[(size_t)i * D + d] = e.batch[(size_t)b * D + d];
      }
      e.rewards[i] = nr;
      e.pert[i] = pert;
    }
  }
  __syncthreads();
  if (tid == 0) { *e.best_reward = new_best; *e.iter = t + 1; }
}

// ---------------------------------------------------------------------------
int launch_eagle_init(vzgp_handle* h, const EagleDev& e) {
  k_eagle_init<<<64, 256, 0, h->stream>>>(e);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}
int launch_eagle_seed_priors(vzgp_handle* h, const EagleDev& e, const double* prior,
                             const double* prior_r, int n, int* ord, double* chosen_r) {
  k_eagle_seed_priors<<<1, 256, 0, h->stream>>>(e, prior, prior_r, n, ord, chosen_r);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}
size_t eagle_suggest_smem(const EagleDev& e) { return sizeof(double) * (size_t)8 * (e.P + e.D); }
size_t eagle_update_smem(const EagleDev& e) {
  return sizeof(double) * (size_t)(e.B + e.count) + (size_t)(e.B + e.count) + 16;
}
int eagle_prepare(const EagleDev& e) {
  VZ_CUDA(cudaFuncSetAttribute(k_eagle_suggest, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)eagle_suggest_smem(e)));
  VZ_CUDA(cudaFuncSetAttribute(k_eagle_update, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)eagle_update_smem(e)));
  return 0;
}
int launch_eagle_suggest(vzgp_handle* h, const EagleDev& e) {
  k_eagle_suggest<<<(e.B + 7) / 8, 256, eagle_suggest_smem(e), h->stream>>>(e);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}
int launch_eagle_update(vzgp_handle* h, const EagleDev& e) {
  k_eagle_update<<<1, 256, eagle_update_smem(e), h->stream>>>(e);
  VZ_CHECK_LAUNCH();
  h->launches++;
  return 0;
}

}  // namespace vzgp
