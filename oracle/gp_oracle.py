"""CPU oracle for the GP-Bandit hot path: kernel, Cholesky, loss, posterior, UCB, trust region.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``vizier_b200``) never
imports anything under ``oracle/`` and fails loudly if its CUDA library is
missing.

What it is: a NumPy/SciPy fp64 restatement of the arithmetic the reference
reaches through TensorFlow-Probability (JAX substrate) on the
``VizierGPBandit.suggest()`` path.  Each function cites the reference
file:line it follows (paths relative to ``/root/reference``).

PARITY STATUS ("parity unpinned" at the TFP boundary).  The arithmetic lives in
third-party packages that are NOT in ``/root/reference`` and not installable
here: ``tfp-nightly[jax]`` (unpinned, ``requirements-jax.txt:8``),
``jax/jaxlib>=0.4.34``, ``jaxopt>=0.8.3``, ``flax``, ``equinox``.  The
reference's own tests hold no golden values for K, L, alpha, mu, sigma, NLL or
grad-NLL.  This oracle is therefore pinned by:
  * the reference's known-answer tests that DO exist (UCB, TrustRegion
    distances/radii; see ``tests/test_oracle_reference_vectors.py``),
  * ``scipy.stats.multivariate_normal`` for the NLL,
  * 50-digit ``mpmath`` re-evaluation of kernel/Cholesky/posterior,
  * central finite differences for the analytic gradient.
Until a JAX/TFP box can dump golden vectors, GPU parity means "matches this
restatement", and every parity report says so.
"""

from __future__ import annotations

import dataclasses
import math
from typing import Optional, Sequence

import numpy as np
import scipy.linalg as sla
import scipy.optimize as sopt

SQRT5 = math.sqrt(5.0)

# Regulariser centres and hyper-parameter bounds:
# vizier/_src/jax/models/tuned_gp_models.py:147-159 (bounds, eps=1e-12),
# :167 (signal variance), :180/:192 (length scales), :269 (noise).
BOUNDARY_EPS = 1e-12
SIGNAL_VARIANCE_BOUNDS = (1e-3 - BOUNDARY_EPS, 10.0 + BOUNDARY_EPS)
LENGTH_SCALE_SQUARED_BOUNDS = (1e-2 - BOUNDARY_EPS, 1e2 + BOUNDARY_EPS)
NOISE_VARIANCE_BOUNDS = (1e-10 - BOUNDARY_EPS, 1.0 + BOUNDARY_EPS)
REG_CENTER_SIGNAL = 0.039
REG_CENTER_LENGTH = 0.5
REG_CENTER_NOISE = 0.0039
REG_WEIGHT = 0.01


@dataclasses.dataclass
class LinearParams:
  """The `linear_coef` variant (tuned_gp_models.py:203-245): the kernel gains
  tfpk.Linear(slope_amplitude = coef * slope, shift = coef * shift) on the length-scaled CONTINUOUS
  features [T: k(x, y) = slope_amplitude^2 * sum_d (x_d/l_d - shift)(y_d/l_d - shift), no bias term], and
  the GP a constant mean coef * mean_constant."""

  coef: float
  slope_amplitude: float      # in [1e-3, 10], regulariser 0.01 log(x/0.039)^2
  shift: float                # unbounded, regulariser 0.5 x^2
  mean_constant: float        # unbounded, regulariser 0.5 x^2


@dataclasses.dataclass
class GPParams:
  """theta of VizierGaussianProcess (tuned_gp_models.py:161-271)."""

  signal_variance: float
  continuous_length_scale_squared: np.ndarray  # [Dc]
  observation_noise_variance: float
  categorical_length_scale_squared: Optional[np.ndarray] = None  # [Dk]
  linear: Optional[LinearParams] = None

  def __post_init__(self):
    self.continuous_length_scale_squared = np.asarray(
        self.continuous_length_scale_squared, dtype=np.float64
    ).reshape(-1)
    if self.categorical_length_scale_squared is None:
      self.categorical_length_scale_squared = np.zeros((0,), np.float64)
    self.categorical_length_scale_squared = np.asarray(
        self.categorical_length_scale_squared, dtype=np.float64
    ).reshape(-1)

  # Flattening order = jaxopt's sorted-key pytree order:
  # categorical_ls2, continuous_ls2, observation_noise_variance, signal_variance
  def to_vector(self) -> np.ndarray:
    lin = [] if self.linear is None else [self.linear.shift, self.linear.slope_amplitude, self.linear.mean_constant]
    # with the linear part the sorted keys are: categorical_ls2, continuous_ls2, linear_shift,
    # linear_slope_amplitude, mean_fn, observation_noise_variance, signal_variance
    return np.concatenate([
        self.categorical_length_scale_squared,
        self.continuous_length_scale_squared,
        lin,
        [self.observation_noise_variance],
        [self.signal_variance],
    ])

  @classmethod
  def from_vector(cls, v: np.ndarray, dc: int, dk: int, linear_coef: Optional[float] = None) -> 'GPParams':
    v = np.asarray(v, dtype=np.float64)
    if linear_coef is not None:
      o = dk + dc
      return cls(
          categorical_length_scale_squared=v[:dk].copy(),
          continuous_length_scale_squared=v[dk : dk + dc].copy(),
          linear=LinearParams(float(linear_coef), slope_amplitude=float(v[o + 1]), shift=float(v[o]), mean_constant=float(v[o + 2])),
          observation_noise_variance=float(v[o + 3]),
          signal_variance=float(v[o + 4]),
      )
    return cls(
        categorical_length_scale_squared=v[:dk].copy(),
        continuous_length_scale_squared=v[dk : dk + dc].copy(),
        observation_noise_variance=float(v[dk + dc]),
        signal_variance=float(v[dk + dc + 1]),
    )


def param_bounds(dc: int, dk: int) -> tuple[np.ndarray, np.ndarray]:
  """Box bounds in to_vector() order (tuned_gp_models.py:147-159)."""
  lo = np.concatenate([
      np.full(dk, LENGTH_SCALE_SQUARED_BOUNDS[0]),
      np.full(dc, LENGTH_SCALE_SQUARED_BOUNDS[0]),
      [NOISE_VARIANCE_BOUNDS[0]],
      [SIGNAL_VARIANCE_BOUNDS[0]],
  ])
  hi = np.concatenate([
      np.full(dk, LENGTH_SCALE_SQUARED_BOUNDS[1]),
      np.full(dc, LENGTH_SCALE_SQUARED_BOUNDS[1]),
      [NOISE_VARIANCE_BOUNDS[1]],
      [SIGNAL_VARIANCE_BOUNDS[1]],
  ])
  return lo, hi


def param_bounds_linear(dc: int, dk: int) -> tuple[np.ndarray, np.ndarray]:
  """Bounds in to_vector() order with the linear part: shift and mean are unbounded (no constraint in
  tuned_gp_models.py:216-245 -> +-inf in jaxopt_wrappers._get_bounds), the slope has the amplitude bounds."""
  lo, hi = param_bounds(dc, dk)
  o = dk + dc
  lo = np.concatenate([lo[:o], [-np.inf, SIGNAL_VARIANCE_BOUNDS[0], -np.inf], lo[o:]])
  hi = np.concatenate([hi[:o], [np.inf, SIGNAL_VARIANCE_BOUNDS[1], np.inf], hi[o:]])
  return lo, hi


def log_uniform_init(rng: np.random.Generator, dc: int, dk: int) -> np.ndarray:
  """theta_0 ~ exp(U*log(hi/lo)+log lo) (tuned_gp_models.py:42-63)."""
  lo, hi = param_bounds(dc, dk)
  u = rng.uniform(size=lo.shape)
  return np.exp(u * np.log(hi / lo) + np.log(lo))


# ----------------------------------------------------------------------------
# Kernel (tuned_gp_models.py:170-201, mask_features.py:46-53)
# ----------------------------------------------------------------------------
def scaled_sq_dist(
    x1: np.ndarray,
    x2: np.ndarray,
    ls2: np.ndarray,
    z1: Optional[np.ndarray] = None,
    z2: Optional[np.ndarray] = None,
    ls2_cat: Optional[np.ndarray] = None,
    cont_dim_valid: Optional[np.ndarray] = None,
    cat_dim_valid: Optional[np.ndarray] = None,
) -> np.ndarray:
  """d2[i,j] = sum_d (x1_id-x2_jd)^2/ls2_d + sum_k 1[z1_ik!=z2_jk]/ls2cat_k.

  Masked (padded) dimensions are zeroed in both arguments before the kernel
  (mask_features.py:46-53) so they contribute nothing.
  """
  x1 = np.asarray(x1, np.float64)
  x2 = np.asarray(x2, np.float64)
  ls2 = np.asarray(ls2, np.float64)
  d2 = np.zeros((x1.shape[0], x2.shape[0]), np.float64)
  dc = x1.shape[1]
  diff = np.empty_like(d2)
  for d in range(dc):
    if cont_dim_valid is not None and not cont_dim_valid[d]:
      continue
    np.subtract(x1[:, d][:, None], x2[:, d][None, :], out=diff)
    np.multiply(diff, diff, out=diff)
    diff /= ls2[d]
    d2 += diff
  if z1 is not None and z1.shape[1] > 0:
    for k in range(z1.shape[1]):
      if cat_dim_valid is not None and not cat_dim_valid[k]:
        continue
      d2 += (z1[:, k][:, None] != z2[:, k][None, :]).astype(np.float64) / ls2_cat[k]
  return d2


def matern52_from_d2(d2: np.ndarray, signal_variance: float) -> np.ndarray:
  """k = sf2*(1+s+s^2/3)*exp(-s), s=sqrt(5*d2); TFP form exp(log1p(s+s^2/3)-s)."""
  s = np.sqrt(d2)
  s *= SQRT5
  poly = s * s
  poly /= 3.0
  poly += s
  np.log1p(poly, out=poly)
  poly -= s
  np.exp(poly, out=poly)
  poly *= signal_variance
  return poly


def kernel(
    params: GPParams, x1, x2, z1=None, z2=None, cont_dim_valid=None, cat_dim_valid=None
) -> np.ndarray:
  d2 = scaled_sq_dist(
      x1,
      x2,
      params.continuous_length_scale_squared,
      z1,
      z2,
      params.categorical_length_scale_squared,
      cont_dim_valid,
      cat_dim_valid,
  )
  k = matern52_from_d2(d2, params.signal_variance)
  if params.linear is not None:
    k += linear_kernel(params, x1, x2, cont_dim_valid)
  return k


def linear_features(params: GPParams, x, cont_dim_valid=None) -> np.ndarray:
  """u_id = x_id / l_d - coef * shift  (masked dimensions are zeroed BEFORE the kernel: u = -coef*shift)."""
  x = np.asarray(x, np.float64)
  if cont_dim_valid is not None:
    x = np.where(np.asarray(cont_dim_valid, bool)[None, :], x, 0.0)
  lin = params.linear
  return x / np.sqrt(params.continuous_length_scale_squared)[None, :] - lin.coef * lin.shift


def linear_kernel(params: GPParams, x1, x2, cont_dim_valid=None) -> np.ndarray:
  lin = params.linear
  a = (lin.coef * lin.slope_amplitude) ** 2
  return a * (linear_features(params, x1, cont_dim_valid) @ linear_features(params, x2, cont_dim_valid).T)


def prior_mean(params: GPParams) -> float:
  return 0.0 if params.linear is None else params.linear.coef * params.linear.mean_constant


def kernel_matrix(
    params: GPParams, x, z=None, row_valid=None, cont_dim_valid=None, cat_dim_valid=None
) -> np.ndarray:
  """K_y = K(X,X) + sn2*I; rows/cols of padded observations replaced by identity.

  stochastic_process_model.py:962-964 (is_missing) and Appendix A.2 [T].
  """
  k = kernel(params, x, x, z, z, cont_dim_valid, cat_dim_valid)
  n = k.shape[0]
  k[np.diag_indices(n)] += params.observation_noise_variance
  if row_valid is not None:
    inv = ~np.asarray(row_valid, bool)
    k[inv, :] = 0.0
    k[:, inv] = 0.0
    k[inv, inv] = 1.0
  return k


# ----------------------------------------------------------------------------
# Cholesky with retry (tuned_gp_models.py:272-280; Appendix A.5 [T])
# ----------------------------------------------------------------------------
def retrying_cholesky(
    a: np.ndarray, jitter: float = 1e-4, max_iters: int = 5
) -> tuple[np.ndarray, float, int]:
  """Returns (L, shift, n_retries).  On repeated failure L contains NaNs."""

  def _try(m):
    try:
      return np.linalg.cholesky(m), True
    except np.linalg.LinAlgError:
      return np.full_like(m, np.nan), False

  l, ok = _try(a)
  shift = 0.0
  it = 0
  while (not ok) and it < max_iters:
    shift = jitter if shift == 0.0 else shift * 10.0
    l, ok = _try(a + shift * np.eye(a.shape[0]))
    it += 1
  return l, shift, it


# ----------------------------------------------------------------------------
# Loss and gradient (stochastic_process_model.py:940-966; Appendix A.2/A.3)
# ----------------------------------------------------------------------------
def regularizer(params: GPParams, cont_dim_valid=None, cat_dim_valid=None) -> float:
  r = REG_WEIGHT * math.log(params.signal_variance / REG_CENTER_SIGNAL) ** 2
  # NOTE: the reference regularises every length-scale entry, padded or not
  # (tuned_gp_models.py:180: jnp.sum over the whole vector).
  r += float(np.sum(REG_WEIGHT * np.log(params.continuous_length_scale_squared / REG_CENTER_LENGTH) ** 2))
  r += float(np.sum(REG_WEIGHT * np.log(params.categorical_length_scale_squared / REG_CENTER_LENGTH) ** 2))
  r += REG_WEIGHT * math.log(params.observation_noise_variance / REG_CENTER_NOISE) ** 2
  if params.linear is not None:   # tuned_gp_models.py:214, 219, 242
    r += REG_WEIGHT * math.log(params.linear.slope_amplitude / REG_CENTER_SIGNAL) ** 2
    r += 0.5 * params.linear.shift ** 2 + 0.5 * params.linear.mean_constant ** 2
  return r


def nll(
    params: GPParams, x, y, z=None, row_valid=None, cont_dim_valid=None, cat_dim_valid=None
) -> float:
  """-log N(y; 0, K_y) over the valid rows (no regulariser).  y [N] or, for M metrics, [N, M]: the
  independent multi-task GP (tuned_gp_models.py:282-288, tfpke.Independent) is M GPs sharing K_y, so
  the terms simply add up."""
  y = np.asarray(y, np.float64)
  n = y.shape[0]
  y2 = y.reshape(n, -1)
  if row_valid is None:
    row_valid = np.ones(n, bool)
  yv = np.where(np.asarray(row_valid, bool)[:, None], y2 - prior_mean(params), 0.0)
  ky = kernel_matrix(params, x, z, row_valid, cont_dim_valid, cat_dim_valid)
  l, _, _ = retrying_cholesky(ky)
  w = sla.solve_triangular(l, yv, lower=True)
  nv = int(np.sum(row_valid))
  m = y2.shape[1]
  return float(0.5 * np.sum(w * w) + m * np.sum(np.log(np.diag(l))) + 0.5 * m * nv * math.log(2 * math.pi))


def loss(params: GPParams, x, y, z=None, row_valid=None, cont_dim_valid=None, cat_dim_valid=None) -> float:
  return nll(params, x, y, z, row_valid, cont_dim_valid, cat_dim_valid) + regularizer(params)


def loss_and_grad(
    theta: np.ndarray, x, y, z=None, row_valid=None, cont_dim_valid=None, cat_dim_valid=None,
    linear_coef: Optional[float] = None,
) -> tuple[float, np.ndarray]:
  """loss(theta) and d loss / d theta in GPParams.to_vector() order (A.3).

  G = K_y^-1 - alpha alpha^T; dNLL/dp = 0.5*sum_ij G_ij dK_ij/dp.
  E_ij = dk/d(d2) = -(5/6)*sf2*(1+s)*exp(-s).
  y [N, M] (independent multi-task GP, one shared K_y): G = M K_y^-1 - sum_m alpha_m alpha_m^T.
  """
  x = np.asarray(x, np.float64)
  n, dc = x.shape
  y = np.asarray(y, np.float64).reshape(n, -1)
  n_metrics = y.shape[1]
  dk = 0 if z is None else z.shape[1]
  p = GPParams.from_vector(theta, dc, dk, linear_coef)
  if row_valid is None:
    row_valid = np.ones(n, bool)
  row_valid = np.asarray(row_valid, bool)
  yv = np.where(row_valid[:, None], y - prior_mean(p), 0.0)

  d2 = scaled_sq_dist(
      x, x, p.continuous_length_scale_squared, z, z, p.categorical_length_scale_squared,
      cont_dim_valid, cat_dim_valid,
  )
  s = SQRT5 * np.sqrt(d2)
  es = np.exp(-s)
  kmat = p.signal_variance * (1.0 + s + s * s / 3.0) * es
  e = -(5.0 / 6.0) * p.signal_variance * (1.0 + s) * es
  ky = kmat.copy()
  if p.linear is not None:
    u = linear_features(p, x, cont_dim_valid)                 # [N, Dc]
    lin_a = (p.linear.coef * p.linear.slope_amplitude) ** 2
    uu = u @ u.T
    ky += lin_a * uu
  ky[np.diag_indices(n)] += p.observation_noise_variance
  inv = ~row_valid
  ky[inv, :] = 0.0
  ky[:, inv] = 0.0
  ky[inv, inv] = 1.0
  l, _, _ = retrying_cholesky(ky)
  w = sla.solve_triangular(l, yv, lower=True)
  alpha = sla.solve_triangular(l.T, w, lower=False)
  nv = int(np.sum(row_valid))
  val = 0.5 * np.sum(w * w) + n_metrics * np.sum(np.log(np.diag(l))) + 0.5 * n_metrics * nv * math.log(2 * math.pi)
  val += regularizer(p)

  linv = sla.solve_triangular(l, np.eye(n), lower=True)
  kinv = linv.T @ linv
  g = n_metrics * kinv - alpha @ alpha.T
  vm = np.outer(row_valid, row_valid).astype(np.float64)
  g = g * vm  # padded rows/cols carry no dependence on theta

  grad_cat = np.zeros(dk)
  for k in range(dk):
    if cat_dim_valid is not None and not cat_dim_valid[k]:
      dterm = 0.0
    else:
      neq = (z[:, k][:, None] != z[:, k][None, :]).astype(np.float64)
      dterm = 0.5 * np.sum(g * e * (-neq / p.categorical_length_scale_squared[k] ** 2))
    lk = p.categorical_length_scale_squared[k]
    grad_cat[k] = dterm + 2 * REG_WEIGHT * math.log(lk / REG_CENTER_LENGTH) / lk
  grad_cont = np.zeros(dc)
  ge = g * e
  for d in range(dc):
    ld = p.continuous_length_scale_squared[d]
    if cont_dim_valid is not None and not cont_dim_valid[d]:
      dterm = 0.0
    else:
      diff = x[:, d][:, None] - x[:, d][None, :]
      dterm = 0.5 * np.sum(ge * (-(diff * diff) / ld**2))
      if p.linear is not None:
        # K_lin = A sum_d u_id u_jd, u_id = x_id w_d - b, w_d = ld^-1/2: dK_lin/d ld = A (x_id u_jd + x_jd u_id) (-w_d^3 / 2)
        w3 = ld ** -1.5
        dterm += 0.5 * np.sum(g * lin_a * (np.outer(x[:, d], u[:, d]) + np.outer(u[:, d], x[:, d]))) * (-0.5 * w3)
    grad_cont[d] = dterm + 2 * REG_WEIGHT * math.log(ld / REG_CENTER_LENGTH) / ld
  sn2 = p.observation_noise_variance
  sf2 = p.signal_variance
  g_noise = 0.5 * np.sum(np.diag(g)) + 2 * REG_WEIGHT * math.log(sn2 / REG_CENTER_NOISE) / sn2
  g_signal = 0.5 * np.sum(g * kmat) / sf2 + 2 * REG_WEIGHT * math.log(sf2 / REG_CENTER_SIGNAL) / sf2
  lin_grads = []
  if p.linear is not None:
    lp = p.linear
    su = u.sum(axis=1)                                           # sum_d u_id
    # d/d shift: b = coef*shift, d(u_i.u_j)/db = -(su_i + su_j)
    g_shift = 0.5 * np.sum(g * (-(su[:, None] + su[None, :]))) * lin_a * lp.coef + lp.shift
    # d/d slope: A = (coef*slope)^2
    g_slope = 0.5 * np.sum(g * uu) * 2.0 * lp.coef**2 * lp.slope_amplitude \
        + 2 * REG_WEIGHT * math.log(lp.slope_amplitude / REG_CENTER_SIGNAL) / lp.slope_amplitude
    # d/d mean_constant: mu = coef*m on the valid rows, dNLL/dmu_i = -alpha_i (summed over metrics)
    g_mean = -lp.coef * float(np.sum(alpha[row_valid])) + lp.mean_constant
    lin_grads = [g_shift, g_slope, g_mean]
  grad = np.concatenate([grad_cat, grad_cont, lin_grads, [g_noise], [g_signal]])
  return float(val), grad


# ----------------------------------------------------------------------------
# ARD driver (jaxopt_wrappers.py:108-199; optimizers/core.py:104-132; A.4)
# ----------------------------------------------------------------------------
def ard_fit(
    x, y, z=None, *, init_thetas: np.ndarray, maxiter: int = 50, row_valid=None,
    cont_dim_valid=None, cat_dim_valid=None,
) -> tuple[np.ndarray, np.ndarray]:
  """Runs SciPy L-BFGS-B from each row of init_thetas; returns (best theta, losses)."""
  x = np.asarray(x, np.float64)
  dc = x.shape[1]
  dk = 0 if z is None else z.shape[1]
  lo, hi = param_bounds(dc, dk)
  finals, losses = [], []
  for t0 in np.atleast_2d(init_thetas):
    res = sopt.minimize(
        loss_and_grad, t0, args=(x, y, z, row_valid, cont_dim_valid, cat_dim_valid), jac=True,
        method='L-BFGS-B', bounds=list(zip(lo, hi)),
        options={'maxiter': maxiter, 'gtol': 1e-8, 'maxls': 20},
    )
    finals.append(res.x)
    losses.append(res.fun)
  losses = np.asarray(losses)
  return finals[int(np.argsort(losses)[0])], losses


# ----------------------------------------------------------------------------
# Posterior (stochastic_process_model.py:968-997, :800-832; Appendix A.6)
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class Predictive:
  params: GPParams
  x: np.ndarray
  z: Optional[np.ndarray]
  chol: np.ndarray  # L, lower
  alpha: np.ndarray  # K_y^-1 y
  row_valid: np.ndarray
  cont_dim_valid: Optional[np.ndarray] = None
  cat_dim_valid: Optional[np.ndarray] = None
  n_retries: int = 0


def precompute_predictive(
    params: GPParams, x, y, z=None, row_valid=None, cont_dim_valid=None, cat_dim_valid=None
) -> Predictive:
  x = np.asarray(x, np.float64)
  n = x.shape[0]
  y = np.asarray(y, np.float64)
  y = y.reshape(-1) if y.ndim == 1 or y.shape[1] == 1 else y   # [N], or [N, M] for M metrics (alpha [N, M])
  if row_valid is None:
    row_valid = np.ones(n, bool)
  row_valid = np.asarray(row_valid, bool)
  ky = kernel_matrix(params, x, z, row_valid, cont_dim_valid, cat_dim_valid)
  l, _, it = retrying_cholesky(ky)
  yv = np.where(row_valid if y.ndim == 1 else row_valid[:, None], y - prior_mean(params), 0.0)
  w = sla.solve_triangular(l, yv, lower=True)
  alpha = sla.solve_triangular(l.T, w, lower=False)
  return Predictive(params, x, z, l, alpha, row_valid, cont_dim_valid, cat_dim_valid, it)


def predict(pred: Predictive, xs, zs=None) -> tuple[np.ndarray, np.ndarray]:
  """Posterior mean and stddev at xs.

  mu = K* alpha; var = sf2 - ||L^-1 K*^T||^2 + sn2 (predictive noise =
  observation noise [T]).  The reference would give NaN for var<0; we clamp at
  0 (documented deviation).
  """
  ks = kernel(pred.params, xs, pred.x, zs, pred.z, pred.cont_dim_valid, pred.cat_dim_valid)
  ks = ks * pred.row_valid[None, :]
  mu = ks @ pred.alpha + prior_mean(pred.params)
  v = sla.solve_triangular(pred.chol, ks.T, lower=True)
  kss = pred.params.signal_variance
  if pred.params.linear is not None:   # k(x*, x*) is no longer constant
    us = linear_features(pred.params, xs, pred.cont_dim_valid)
    kss = kss + (pred.params.linear.coef * pred.params.linear.slope_amplitude) ** 2 * np.sum(us * us, axis=1)
  var = kss - np.sum(v * v, axis=0) + pred.params.observation_noise_variance
  return mu, np.sqrt(np.maximum(var, 0.0))


def predict_ensemble(preds: Sequence[Predictive], xs, zs=None):
  """Equal-weight mixture (stochastic_process_model.py:858-867; A.6 [T])."""
  mus, sds = zip(*[predict(p, xs, zs) for p in preds])
  mus, sds = np.stack(mus), np.stack(sds)
  mean = mus.mean(0)
  var = (sds**2 + mus**2).mean(0) - mean**2
  return mean, np.sqrt(np.maximum(var, 0.0))


# ----------------------------------------------------------------------------
# Acquisition (acquisitions.py:213-225 UCB; :152-174 trust region application;
# :691-820 TrustRegion)
# ----------------------------------------------------------------------------
def ucb(mu: np.ndarray, sd: np.ndarray, coefficient: float = 1.8) -> np.ndarray:
  return mu + coefficient * sd


TR_MIN_RADIUS = 0.2
TR_DIMENSION_FACTOR = 5.0


def trust_region_dim_mask(continuous_feasible_values: Sequence[np.ndarray]) -> np.ndarray:
  """acquisitions.py:734-749."""
  mask = []
  for fv in continuous_feasible_values:
    fv = np.asarray(fv, np.float64)
    if fv.size == 0:
      mask.append(True)
    elif fv.size == 1:
      mask.append(False)
    else:
      mask.append(bool(np.max(np.diff(np.sort(fv))) <= TR_MIN_RADIUS))
  return np.asarray(mask, bool)


def trust_radius(num_obs: int, continuous_dof: int, categorical_dof: int) -> float:
  """acquisitions.py:757-777."""
  if num_obs == 0:
    return 1.0
  dof = continuous_dof + categorical_dof
  original_num_obs = 0.1 * num_obs + 0.9 * num_obs
  trust_level = original_num_obs / (TR_DIMENSION_FACTOR * (dof + 1))
  return TR_MIN_RADIUS + (0.5 - TR_MIN_RADIUS) * trust_level


def min_linf_distance(
    xs: np.ndarray, trusted: np.ndarray, dim_mask: np.ndarray, row_valid=None
) -> np.ndarray:
  """acquisitions.py:779-820.  xs [..., D], trusted [N, D] -> [...]."""
  xs = np.asarray(xs, np.float64)
  trusted = np.asarray(trusted, np.float64)
  if trusted.size == 0 or xs.shape[-1] == 0:
    return -np.inf * np.ones(xs.shape[:-1])
  # Same arithmetic as the reference's broadcasted (..., N, D) array, evaluated one feature
  # dimension at a time so the CPU baseline does not pay for a 3-D temporary.
  flat = xs.reshape(-1, xs.shape[-1])
  linf = np.zeros((flat.shape[0], trusted.shape[0]), np.float64)
  tmp = np.empty_like(linf)
  for d in range(flat.shape[1]):
    if not dim_mask[d]:
      continue
    np.subtract(flat[:, d][:, None], trusted[:, d][None, :], out=tmp)
    np.abs(tmp, out=tmp)
    np.maximum(linf, tmp, out=linf)
  if row_valid is not None:
    linf[:, ~np.asarray(row_valid, bool)] = np.inf
  return np.min(linf, axis=-1).reshape(xs.shape[:-1])


def apply_trust_region(acq: np.ndarray, distance: np.ndarray, radius: float) -> np.ndarray:
  """acquisitions.py:160-166."""
  return np.where((distance <= radius) | (radius > 0.5), acq, -1e4 - distance)


def score_with_aux(
    pred: Predictive, xs, zs=None, *, coefficient: float = 1.8,
    tr_dim_mask: Optional[np.ndarray] = None, categorical_dof: int = 0,
    use_trust_region: bool = True,
):
  """BayesianScoringFunction.score_with_aux (acquisitions.py:190-207)."""
  mu, sd = predict(pred, xs, zs)
  acq = ucb(mu, sd, coefficient)
  aux = {}
  if use_trust_region:
    xs = np.asarray(xs, np.float64)
    if tr_dim_mask is None:
      tr_dim_mask = np.ones(xs.shape[-1], bool)
    dist = min_linf_distance(xs, pred.x, tr_dim_mask, pred.row_valid)
    radius = trust_radius(int(np.sum(pred.row_valid)), int(np.sum(tr_dim_mask)), categorical_dof)
    raw = acq
    acq = apply_trust_region(acq, dist, radius)
    aux = {
        'mean': mu, 'stddev': sd, 'raw_acquisition': raw,
        'linf_distance': dist, 'radius': np.ones_like(dist) * radius,
    }
  return acq, aux


def top_k(scores: np.ndarray, count: int) -> np.ndarray:
  """Indices of the `count` largest scores, descending; ties -> lowest index.

  vectorized_base.py:580 uses argpartition (unordered) followed by an argsort
  in best_candidates_to_trials (:598); the net effect is this ordering.
  """
  scores = np.asarray(scores, np.float64)
  order = np.lexsort((np.arange(scores.shape[0]), -scores))
  return order[:count]


# ----------------------------------------------------------------------------
# GP-UCB-PE acquisition (vizier/_src/algorithms/designers/gp_ucb_pe.py)
# ----------------------------------------------------------------------------
def ucb_threshold(pred_a: Predictive, pred_b: Predictive, ucb_coefficient: float = 1.8) -> float:
  """_compute_ucb_threshold (:175-218): mean of model A at the point of B's features with max UCB."""
  mu, sd = predict(pred_a, pred_b.x, pred_b.z)
  u = np.where(pred_b.row_valid, mu + ucb_coefficient * sd, -np.inf)
  return float(mu[int(np.argmax(u))])


def ucb_pe_score(pred_a: Predictive, pred_b: Predictive, xs, zs=None, *, mode: int, ucb_coefficient=1.8,
                 explore_coefficient=0.5, penalty_coefficient=10.0, threshold=0.0, tr_dim_mask=None,
                 tr_rows=None, trust_radius_value=None, use_trust_region=True):
  """UCBScoreFunction.score_with_aux (:344-381, mode 0) / PEScoreFunction.score_with_aux (:434-492,
  mode 1), single metric, with the strict trust region of :221-242 over the first tr_rows rows of B."""
  mu, sd = predict(pred_a, xs, zs)
  _, sd_all = predict(pred_b, xs, zs)
  if mode == 0:
    acq = mu + ucb_coefficient * sd_all
  else:
    acq = sd_all + penalty_coefficient * np.minimum(mu + sd * explore_coefficient - threshold, 0.0)
  if use_trust_region:
    xs = np.asarray(xs, np.float64)
    if tr_dim_mask is None:
      tr_dim_mask = np.ones(xs.shape[-1], bool)
    n_tr = pred_b.x.shape[0] if tr_rows is None else tr_rows
    dist = min_linf_distance(xs, pred_b.x[:n_tr], tr_dim_mask)
    acq = np.where((dist < trust_radius_value) | (trust_radius_value > 0.5), acq, -1e4 - dist)
  return acq, {'mean': mu, 'stddev': sd, 'stddev_from_all': sd_all}


# ----------------------------------------------------------------------------
# Transfer learning: stacked residual GPs (gp/gp_models.py:91-140, :245-300; gp/transfer_learning.py:38-152)
# ----------------------------------------------------------------------------
def transfer_dof(n: int, num_hyperparameters: int) -> float:
  """_compute_dof (gp/transfer_learning.py:38-59)."""
  return max(n - num_hyperparameters, n / (1 + num_hyperparameters))


def transfer_alpha(n_top: int, n_base: int, num_hyperparameters: int, expected_base_stddev_mismatch: float = 1.0) -> float:
  """combine_predictions_with_aux (gp/transfer_learning.py:96-118): weight of the top stddev in the geometric mean."""
  dof_base, dof_top = transfer_dof(n_base, num_hyperparameters), transfer_dof(n_top, num_hyperparameters)
  beta_squared = (dof_top / dof_base) * (1 + dof_base + expected_base_stddev_mismatch ** 2)
  return beta_squared / (1 + beta_squared)


def predict_stack(preds: Sequence[Predictive], xs, zs=None):
  """StackedResidualGP.predict_with_aux, recursively (gp/gp_models.py:110-140): preds[0] is the first prior study's GP,
  preds[e] was trained on the residuals of study e against preds[:e]; the last one belongs to the current study.
  mean = sum of the means; stddev = stddev_e^alpha_e * (stddev of the stack below)^(1 - alpha_e), bottom-up, with
  alpha_e from the training-set sizes of levels e and e-1 and D + 2 hyper-parameters (gp/gp_models.py:77-88)."""
  mean, sd = predict(preds[0], xs, zs)
  for e in range(1, len(preds)):
    p = preds[e]
    h = p.x.shape[1] + (0 if p.z is None else p.z.shape[1]) + 2
    alpha = transfer_alpha(int(np.sum(p.row_valid)), int(np.sum(preds[e - 1].row_valid)), h)
    mu_e, sd_e = predict(p, xs, zs)
    mean = mean + mu_e
    sd = sd_e ** alpha * sd ** (1.0 - alpha)
  return mean, sd


def stack_residual_labels(preds: Sequence[Predictive], x, y, z=None) -> np.ndarray:
  """Labels of the next level: y minus the mean of the stack built so far (gp/gp_models.py:268-283)."""
  if not preds:
    return np.asarray(y, np.float64)
  return np.asarray(y, np.float64) - predict_stack(preds, x, z)[0]


def predictive_covariance(pred: Predictive, xs) -> np.ndarray:
  """Joint posterior predictive covariance at xs [m, D]: K** - V^T V + sn2 I (the GPRM of
  stochastic_process_model.py:800-868 with predictive noise = observation noise [T])."""
  xs = np.asarray(xs, np.float64)
  ks = kernel(pred.params, xs, pred.x, None, pred.z, pred.cont_dim_valid, pred.cat_dim_valid)
  ks = ks * pred.row_valid[None, :]
  v = sla.solve_triangular(pred.chol, ks.T, lower=True)
  kss = kernel(pred.params, xs, xs, None, None, pred.cont_dim_valid, pred.cat_dim_valid)
  return kss - v.T @ v + pred.params.observation_noise_variance * np.eye(xs.shape[0])


def set_pe_score(pred_a: Predictive, pred_b: Predictive, xs_sets, *, explore_coefficient=0.5, penalty_coefficient=10.0,
                 threshold=0.0, tr_dim_mask=None, tr_rows=None, trust_radius_value=None, use_trust_region=True):
  """SetPEScoreFunction.score_with_aux (gp_ucb_pe.py:510-594) for xs_sets [S, q, D]: logdet of the joint predictive
  covariance under model B (-inf if not positive definite, :495-507) + penalty * sum_i min(mean_A + explore *
  stddev_A - threshold, 0) + the set trust-region term (:245-269: sum_i (dist_i > r) & (r <= 0.5) * (-1e4 - dist_i))."""
  xs_sets = np.asarray(xs_sets, np.float64)
  n_sets, q, d = xs_sets.shape
  flat = xs_sets.reshape(n_sets * q, d)
  mu, sd = predict(pred_a, flat)
  acq = np.zeros(n_sets)
  sd_all = np.zeros(n_sets * q)
  for s in range(n_sets):
    cov = predictive_covariance(pred_b, xs_sets[s])
    sd_all[s * q:(s + 1) * q] = np.sqrt(np.maximum(np.diag(cov), 0.0))
    try:
      chol = np.linalg.cholesky(cov)
      logdet = 2.0 * np.sum(np.log(np.diag(chol)))
      if not np.isfinite(logdet):
        logdet = -np.inf
    except np.linalg.LinAlgError:
      logdet = -np.inf
    viol = np.minimum(mu[s * q:(s + 1) * q] + explore_coefficient * sd[s * q:(s + 1) * q] - threshold, 0.0)
    acq[s] = logdet + penalty_coefficient * np.sum(viol)
  if use_trust_region:
    if tr_dim_mask is None:
      tr_dim_mask = np.ones(d, bool)
    n_tr = pred_b.x.shape[0] if tr_rows is None else tr_rows
    dist = min_linf_distance(flat, pred_b.x[:n_tr], tr_dim_mask).reshape(n_sets, q)
    acq = acq + np.sum(((dist > trust_radius_value) & (trust_radius_value <= 0.5)) * (-1e4 - dist), axis=1)
  return acq, {'mean': mu, 'stddev': sd, 'stddev_from_all': sd_all}


# ----------------------------------------------------------------------------
# Multi-metric: hyper-volume scalarised UCB (gp_bandit.py:214-242; acquisitions.py:132-149, 571-625;
# scalarization.py:85-111).  The scalarisation weights are an input (the reference draws them with
# jax.random.normal: abs, then rows normalised to unit L2 norm).
# ----------------------------------------------------------------------------
def hv_reference_point(labels: np.ndarray, scale: float = 0.01) -> np.ndarray:
  """worst - scale * (best - worst) per metric (acquisitions.py:132-149; labels are to be maximised)."""
  labels = np.asarray(labels, np.float64)
  best, worst = labels.max(axis=0), labels.min(axis=0)
  return worst - scale * (best - worst)


def hv_scalarize(objectives: np.ndarray, weights: np.ndarray, reference_point=None) -> np.ndarray:
  """HyperVolumeScalarization.__call__ (scalarization.py:95-111): objectives [..., M], weights [S, M]
  -> [S, ...]:  min_m(max(obj_m - ref_m, 0) / w_sm) ** M."""
  obj = np.asarray(objectives, np.float64)
  if reference_point is not None:
    obj = obj - reference_point
  obj = np.maximum(obj, 0.0)
  w = np.asarray(weights, np.float64)
  prod = obj[None, ...] * (1.0 / w).reshape((w.shape[0],) + (1,) * (obj.ndim - 1) + (w.shape[1],))
  return np.min(prod, axis=-1) ** obj.shape[-1]


def scalarized_ucb(mu: np.ndarray, sd: np.ndarray, weights: np.ndarray, reference_point, max_scalarized=None,
                   coefficient: float = 1.8) -> np.ndarray:
  """ScalarizeOverAcquisitions(UCB, HV scalarizer, mean over scalarisations, max with the best observed
  scalarised label) - mu [B, M], sd [B] (shared by the metrics of the independent multi-task GP)."""
  u = mu + coefficient * np.asarray(sd, np.float64)[:, None]
  sc = hv_scalarize(u, weights, reference_point)      # [S, B]
  if max_scalarized is not None:
    sc = np.maximum(sc, np.asarray(max_scalarized, np.float64)[:, None])
  return sc.mean(axis=0)


def hv_max_scalarized(labels: np.ndarray, weights: np.ndarray, reference_point) -> np.ndarray:
  """max over the observed label vectors of the scalarisation, per weight vector (gp_bandit.py:229-232)."""
  return hv_scalarize(labels, weights, reference_point).max(axis=-1)
