"""Trials <-> scaled feature arrays for the GP designer (host NumPy, O(N*D)).

Mirrors the slice of vizier/pyvizier/converters/{core,jnp_converters}.py used by
`VizierGPBandit` (`TrialToModelInputConverter.from_problem(problem, scale=True,
max_discrete_indices=0, flip_sign_for_minimization_metrics=True)`, gp_bandit.py:190-196):
  * DOUBLE                -> one continuous feature scaled to [0,1] (LINEAR / LOG / REVERSE_LOG,
                             core.py:425-482; low==high maps to 0.5),
  * INTEGER / DISCRETE    -> continuified (max_discrete_indices=0), same scaling; suggestions are
                             rounded to the nearest feasible value (core.py:667-682),
  * CATEGORICAL           -> one int32 index feature; unknown values map to len(feasible_values)
                             (core.py:700-708),
  * labels                -> [N, n_metrics] float64, sign-flipped for MINIMIZE metrics, NaN for
                             infeasible trials / missing metrics (core.py:796-823).
Feature padding (PaddingSchedule) is not reproduced: the default schedule is "no padding"
(gp_bandit.py:139-141) and the CUDA kernels take explicit sizes instead of padded shapes.
"""

from __future__ import annotations

import dataclasses
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np

from vizier_b200 import vz


def _tname(t) -> str:
  return getattr(t, 'name', str(t))


@dataclasses.dataclass
class _ContinuousSpec:
  name: str
  ptype: str                   # DOUBLE / INTEGER / DISCRETE
  low: float
  high: float
  feasible_values: Tuple[float, ...]   # () for DOUBLE
  scale: Optional[str]
  forward: Callable[[np.ndarray], np.ndarray]
  backward: Callable[[np.ndarray], np.ndarray]


@dataclasses.dataclass
class _CategoricalSpec:
  name: str
  feasible_values: Tuple[str, ...]

  @property
  def size(self) -> int:
    # bounds[1] of the reference's DISCRETE spec (core.py:183-191) = number of categories; the
    # optimisers sample indices in [0, size) (eagle_strategy.py:209-211, :293-302)
    return len(self.feasible_values)


def _make_scaler(low: float, high: float, scale: Optional[str]):
  """ModelInputArrayBijector.scaler_from_spec (core.py:425-482)."""
  if low == high:
    return (lambda x: np.where(np.isfinite(x), x - low + 0.5, x)), (lambda y: np.where(np.isfinite(y), y + low - 0.5, y))
  if scale == 'LOG':
    if low <= 0 or high <= 0:
      raise ValueError(f'Log scale requires positive bounds, got [{low}, {high}].')
    llo, lhi = np.log(low), np.log(high)
    denom = (lhi - llo) or 1.0
    return (lambda x: (np.log(x) - llo) / denom), (lambda y: np.exp(y * denom + llo))
  if scale == 'REVERSE_LOG':
    raw_sum = low + high
    llo, lhi = np.log(low), np.log(high)
    denom = (lhi - llo) or 1.0
    return (lambda x: 1.0 - (np.log(raw_sum - x) - llo) / denom), (lambda y: raw_sum - np.exp(lhi - denom * y))
  if high - low == 1.0 and low == 0:
    return (lambda x: x), (lambda y: y)
  return (lambda x: (x - low) / (high - low)), (lambda y: y * (high - low) + low)


class TrialToModelInputConverter:
  """See module docstring.  Features are returned as (continuous f64 [N,Dc], categorical i32 [N,Dk])."""

  def __init__(self, problem):
    self._problem = problem
    self.continuous_specs: List[_ContinuousSpec] = []
    self.categorical_specs: List[_CategoricalSpec] = []
    self._param_order: List[Tuple[str, str, int]] = []  # (name, 'c'|'k', index)
    for p in problem.search_space.parameters:
      t = _tname(p.type)
      if t == 'CATEGORICAL':
        self._param_order.append((p.name, 'k', len(self.categorical_specs)))
        self.categorical_specs.append(_CategoricalSpec(p.name, tuple(p.feasible_values)))
      elif t in ('DOUBLE', 'INTEGER', 'DISCRETE'):
        low, high = float(p.bounds[0]), float(p.bounds[1])
        scale = _tname(p.scale_type) if p.scale_type is not None else None
        if scale == 'UNIFORM_DISCRETE':
          scale = None
        fwd, bwd = _make_scaler(low, high, scale)
        fv = () if t == 'DOUBLE' else tuple(float(v) for v in p.feasible_values)
        self._param_order.append((p.name, 'c', len(self.continuous_specs)))
        self.continuous_specs.append(_ContinuousSpec(p.name, t, low, high, fv, scale, fwd, bwd))
      else:
        raise ValueError(f'Unsupported parameter type {t} for {p.name}')
    self.metric_specs = list(problem.metric_information)

  @classmethod
  def from_problem(cls, problem, *, scale: bool = True, max_discrete_indices: int = 0,
                   flip_sign_for_minimization_metrics: bool = True, dtype=np.float64, padding_schedule=None):
    if not scale or max_discrete_indices != 0 or not flip_sign_for_minimization_metrics:
      raise ValueError('Only the GP-bandit configuration (scale=True, max_discrete_indices=0, flip sign) is supported.')
    del dtype, padding_schedule
    return cls(problem)

  @property
  def n_continuous(self) -> int:
    return len(self.continuous_specs)

  @property
  def n_categorical(self) -> int:
    return len(self.categorical_specs)

  @property
  def categorical_sizes(self) -> List[int]:
    return [s.size for s in self.categorical_specs]

  # -- trials -> arrays --------------------------------------------------------
  @staticmethod
  def _value(trial, name):
    pv = trial.parameters.get(name) if hasattr(trial.parameters, 'get') else None
    if pv is None:
      return None
    return getattr(pv, 'value', pv)

  def to_features(self, trials: Sequence[Any]) -> Tuple[np.ndarray, np.ndarray]:
    n = len(trials)
    cont = np.full((n, self.n_continuous), np.nan, dtype=np.float64)
    cat = np.zeros((n, self.n_categorical), dtype=np.int32)
    for j, s in enumerate(self.continuous_specs):
      col = np.array([np.nan if (v := self._value(t, s.name)) is None else float(v) for t in trials], dtype=np.float64)
      if n:
        cont[:, j] = s.forward(col)
    for j, s in enumerate(self.categorical_specs):
      index = {v: i for i, v in enumerate(s.feasible_values)}
      cat[:, j] = [index.get(self._value(t, s.name), len(s.feasible_values)) for t in trials]
    return cont, cat

  def to_labels(self, trials: Sequence[Any]) -> np.ndarray:
    out = np.full((len(trials), len(self.metric_specs)), np.nan, dtype=np.float64)
    for j, m in enumerate(self.metric_specs):
      sign = -1.0 if int(m.goal) == int(vz.ObjectiveMetricGoal.MINIMIZE) else 1.0
      for i, t in enumerate(trials):
        fm = getattr(t, 'final_measurement', None)
        if getattr(t, 'infeasible', False) or fm is None or m.name not in fm.metrics:
          continue
        out[i, j] = sign * float(fm.metrics[m.name].value)
    return out

  def to_xy(self, trials: Sequence[Any]):
    return self.to_features(trials), self.to_labels(trials)

  def to_xy_cached(self, trials: Sequence[Any]):
    """`to_xy` for trials the caller owns and never mutates (the designers' deep copies of COMPLETED
    trials): rows are converted once per trial object and re-stacked afterwards, so a study that grows by
    a few trials per `update` does not pay the O(N*D) Python conversion loop on every `suggest`."""
    cache = self.__dict__.setdefault('_row_cache', {})
    new = [t for t in trials if id(t) not in cache]
    if new:
      (cont, cat), labels = self.to_xy(new)
      for i, t in enumerate(new):
        cache[id(t)] = (t, cont[i], cat[i], labels[i])   # holding t keeps id(t) unique
    n = len(trials)
    cont = np.empty((n, self.n_continuous), np.float64)
    cat = np.empty((n, self.n_categorical), np.int32)
    labels = np.empty((n, len(self.metric_specs)), np.float64)
    for i, t in enumerate(trials):
      _, c, z, y = cache[id(t)]
      cont[i], cat[i], labels[i] = c, z, y
    return (cont, cat), labels

  # -- arrays -> parameters ------------------------------------------------------
  def to_parameters(self, continuous: np.ndarray, categorical: Optional[np.ndarray] = None) -> List[Any]:
    if self.n_continuous:
      continuous = np.asarray(continuous, dtype=np.float64).reshape(-1, self.n_continuous)
      n = continuous.shape[0]
    else:
      n = 0 if categorical is None else np.asarray(categorical).reshape(-1, self.n_categorical).shape[0]
      continuous = np.zeros((n, 0))
    if categorical is None:
      categorical = np.zeros((n, self.n_categorical), dtype=np.int32)
    categorical = np.asarray(categorical).reshape(n, self.n_categorical)
    out = []
    for i in range(n):
      pd = vz.ParameterDict()
      for name, kind, j in self._param_order:
        if kind == 'c':
          s = self.continuous_specs[j]
          v = float(s.backward(np.asarray(continuous[i, j])))
          if not np.isfinite(v):
            continue
          if s.ptype == 'DOUBLE':
            pd[name] = float(np.clip(v, s.low, s.high))
          else:
            fv = np.asarray(s.feasible_values)
            nearest = fv[int(np.argmin(np.abs(fv - v)))]
            pd[name] = int(nearest) if s.ptype == 'INTEGER' else float(nearest)
        else:
          s = self.categorical_specs[j]
          idx = int(categorical[i, j])
          if 0 <= idx < len(s.feasible_values):
            pd[name] = s.feasible_values[idx]
      out.append(pd)
    return out

  def continuous_feasible_values(self, max_num_feasible_values: Optional[int] = None) -> List[np.ndarray]:
    """jnp_converters.py:249-299: scaled feasible values per continuified parameter ([] = continuous)."""
    res = []
    for s in self.continuous_specs:
      if s.ptype == 'DOUBLE':
        res.append(np.asarray([]))
        continue
      if max_num_feasible_values is not None and len(s.feasible_values) > max_num_feasible_values:
        res.append(np.asarray([]))
      else:
        res.append(np.asarray(s.forward(np.asarray(s.feasible_values, dtype=np.float64))))
    return res


def trials_to_sorted_features(trials: Sequence[Any], converter: TrialToModelInputConverter, features=None):
  """vectorized_base.py:655-665: prior trials ordered by creation time.

  `features` = (continuous, categorical) already converted for `trials` in the given order: only the
  permutation is applied then (the conversion is an O(N*D) Python loop, worth not repeating per suggestion)."""
  if not trials:
    return None
  order = sorted(range(len(trials)), key=lambda i: (trials[i].creation_time, getattr(trials[i], '_seq', 0)))
  if features is not None:
    cont, cat = features
    idx = np.asarray(order, dtype=np.int64)
    return cont[idx], cat[idx]
  return converter.to_features([trials[i] for i in order])
