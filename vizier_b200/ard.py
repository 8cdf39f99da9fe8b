"""ARD: fits the GP hyper-parameters by multi-restart L-BFGS-B on the device loss/gradient.

Mirrors `gp_models._train_gp` (vizier/_src/algorithms/designers/gp/gp_models.py:169-223) with the
default optimiser `JaxoptScipyLbfgsB(LbfgsBOptions(maxiter=50))`
(vizier/_src/jax/optimizers/jaxopt_wrappers.py:108-199; vizier/jax/optimizers.py:30-37):
  * `random_restarts` initial points drawn log-uniformly inside the box
    (tuned_gp_models.py:42-63),
  * each refined by SciPy's L-BFGS-B in the CONSTRAINED space with box bounds, gtol=1e-8,
    maxls=20 (the same Fortran routine jaxopt.ScipyBoundedMinimize reaches),
  * the `best_n` lowest final losses are kept (optimizers/core.py:104-132).
loss(theta) and its gradient come from `vzgp_nll_grad` (CUDA); SciPy only drives.
The draws use NumPy's Generator rather than JAX threefry: seeded trajectories are reproducible
but not bit-identical to a JAX run.
"""

from __future__ import annotations

import dataclasses
from typing import List, Optional, Tuple

import numpy as np
import scipy.optimize as sopt

from vizier_b200 import gp

DEFAULT_RANDOM_RESTARTS = 4


@dataclasses.dataclass(frozen=True)
class LbfgsBOptions:
  num_line_search_steps: int = 20
  tol: float = 1e-8
  maxiter: int = 50


def log_uniform_init(rng: np.random.Generator, dc: int, dk: int, n: int) -> np.ndarray:
  lo, hi = gp.param_bounds(dc, dk)
  u = rng.uniform(size=(n, lo.shape[0]))
  return np.exp(u * np.log(hi / lo) + np.log(lo))


@dataclasses.dataclass
class ScipyLbfgsB:
  """The reference's default ARD optimiser, driving the CUDA loss."""

  options: LbfgsBOptions = LbfgsBOptions()

  def __call__(self, init_thetas: np.ndarray, loss_and_grad, bounds, best_n: int = 1):
    finals, losses = [], []
    for t0 in np.atleast_2d(init_thetas):
      res = sopt.minimize(loss_and_grad, t0, jac=True, method='L-BFGS-B', bounds=bounds,
                          options={'maxiter': self.options.maxiter, 'gtol': self.options.tol,
                                   'maxls': self.options.num_line_search_steps})
      finals.append(res.x)
      losses.append(float(res.fun))
    losses = np.asarray(losses)
    order = np.argsort(losses)[:max(1, best_n)]
    return [finals[i] for i in order], losses


def train_gp(dev: gp.DeviceGP, x, y, z=None, *, rng: np.random.Generator,
             random_restarts: int = DEFAULT_RANDOM_RESTARTS, ensemble_size: int = 1,
             optimizer: Optional[ScipyLbfgsB] = None, n_valid: Optional[int] = None
             ) -> Tuple[List[gp.GPHyperParams], np.ndarray]:
  """Returns the best `ensemble_size` hyper-parameter sets and all final losses.

  x [N,Dc] float64, z [N,Dk] int32 or None, y [N]: device tensors or arrays (copied once).
  """
  import torch  # device-memory handles only
  optimizer = optimizer or ScipyLbfgsB()
  xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev.device)
  yt = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(y).reshape(-1), dtype=np.float64)).to(dev.device)
  zt = None
  if z is not None and np.prod(tuple(z.shape)) > 0:
    zt = z if isinstance(z, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(z, dtype=np.int32)).to(dev.device)
  dc = xt.shape[1]
  dk = 0 if zt is None else zt.shape[1]
  lo, hi = gp.param_bounds(dc, dk)
  inits = log_uniform_init(rng, dc, dk, random_restarts)

  def f(theta):
    p = gp.GPHyperParams.from_vector(theta, dc, dk)
    loss, grad, _ = dev.loss_and_grad(xt, yt, p, z=zt, n_valid=n_valid)
    if not np.isfinite(loss):
      return 1e300, np.zeros_like(theta)
    return loss, grad

  best, losses = optimizer(inits, f, list(zip(lo, hi)), best_n=ensemble_size)
  return [gp.GPHyperParams.from_vector(t, dc, dk) for t in best], losses
