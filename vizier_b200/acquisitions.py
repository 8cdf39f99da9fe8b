"""Host-side acquisition parameters: UCB coefficient and the trust region's scalar state.

Mirrors vizier/_src/algorithms/designers/gp/acquisitions.py: `UCB` (:213-225, coefficient 1.8),
`TrustRegion.__post_init__` (:734-749, which dimensions take part), `TrustRegion.trust_radius`
(:757-777).  The per-candidate work (min L-inf distance, thresholding, -1e4 - distance penalty)
runs inside the CUDA scoring kernel; only these O(D) scalars are computed here.
"""

from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from vizier_b200 import gp

TR_MIN_RADIUS = 0.2        # TrustRegion.min_radius (acquisitions.py:751-754)
TR_DIMENSION_FACTOR = 5.0  # acquisitions.py:760
DEFAULT_UCB_COEFFICIENT = 1.8


def trust_region_dim_mask(continuous_feasible_values: Sequence[np.ndarray]) -> np.ndarray:
  """True for dimensions used in the L-inf distance.

  Continuous parameters (empty feasible list) always take part; a discrete parameter takes part
  only if its scaled feasible values have no gap larger than min_radius; single-valued ones never.
  """
  mask = []
  for fv in continuous_feasible_values:
    fv = np.asarray(fv, dtype=np.float64).reshape(-1)
    if fv.size == 0:
      mask.append(True)
    elif fv.size == 1:
      mask.append(False)
    else:
      mask.append(bool(np.max(np.diff(np.sort(fv))) <= TR_MIN_RADIUS))
  return np.asarray(mask, dtype=bool)


def trust_radius(num_obs: int, continuous_dof: int, categorical_dof: int = 0) -> float:
  """0.2 + 0.3 * num_obs / (5 * (dof + 1)); 1.0 with no observations."""
  if num_obs == 0:
    return 1.0
  dof = continuous_dof + categorical_dof
  trust_level = (0.1 * num_obs + 0.9 * num_obs) / (TR_DIMENSION_FACTOR * (dof + 1))
  return TR_MIN_RADIUS + (0.5 - TR_MIN_RADIUS) * trust_level


def make_acquisition(num_obs: int, continuous_feasible_values: Optional[Sequence[np.ndarray]],
                     n_continuous: int, n_categorical: int = 0, *, use_trust_region: bool = True,
                     ucb_coefficient: float = DEFAULT_UCB_COEFFICIENT) -> gp.Acquisition:
  """What `bayesian_scoring_function_factory` (acquisitions.py:368-387) builds, as kernel params."""
  if continuous_feasible_values is None:
    mask = np.ones(n_continuous, dtype=bool)
  else:
    mask = trust_region_dim_mask(continuous_feasible_values)
    if mask.shape[0] != n_continuous:
      raise ValueError(f'{mask.shape[0]} feasible-value lists for {n_continuous} continuous features')
  radius = trust_radius(num_obs, int(mask.sum()), n_categorical)
  return gp.Acquisition(ucb_coefficient, use_trust_region, radius, mask)


def hv_reference_point(labels: np.ndarray, scale: float = 0.01) -> np.ndarray:
  """worst - scale * (best - worst) per metric (acquisitions.py:132-149; labels are maximised)."""
  labels = np.asarray(labels, np.float64)
  best, worst = labels.max(axis=0), labels.min(axis=0)
  return worst - scale * (best - worst)


def hv_scalarize(objectives: np.ndarray, weights: np.ndarray, reference_point=None) -> np.ndarray:
  """HyperVolumeScalarization (scalarization.py:95-111): objectives [N, M], weights [S, M] -> [S, N]:
  min_m(max(obj_m - ref_m, 0) / w_sm) ** M.  Host NumPy: only the observed labels go through here, the
  candidates are scalarised on the device (csrc/multi.cu)."""
  obj = np.asarray(objectives, np.float64)
  if reference_point is not None:
    obj = obj - reference_point
  obj = np.maximum(obj, 0.0)
  prod = obj[None, :, :] / np.asarray(weights, np.float64)[:, None, :]
  return np.min(prod, axis=-1) ** obj.shape[-1]
