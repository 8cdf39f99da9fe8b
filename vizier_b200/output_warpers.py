"""Label warping for the GP designer (host NumPy; O(N) work outside the hot path).

Mirrors vizier/_src/algorithms/designers/gp/output_warpers.py: `create_default_warper`
(:185-213) = HalfRankComponent (:285-365) -> LogWarperComponent (:368-409) ->
InfeasibleWarperComponent (:412-494) inside OutputWarperPipeline (:118-182), with `unwarp`
for predict/sample.  Labels are (num_points, 1) arrays, maximisation convention, NaN = infeasible.
The half-rank step is vectorised (the reference loops in Python, :348-357) but produces the same
numbers; tests/test_output_warpers.py pins it to the reference's golden arrays.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
from scipy import stats


def _validate_labels(labels: np.ndarray) -> np.ndarray:
  labels = np.array(labels, dtype=float)  # copy
  if not (labels.ndim == 2 and labels.shape[-1] == 1):
    raise ValueError(f'Labels need to be an array of shape (num_points, 1). Got shape: {labels.shape}')
  if np.isposinf(labels).any():
    raise ValueError('Infinity metric value is not valid.')
  labels[np.isneginf(labels)] = np.nan
  return labels


class HalfRankComponent:
  """Maps the below-median half onto a Gaussian tail fitted to the above-median half."""

  def __init__(self):
    self._orig: Optional[np.ndarray] = None
    self._warped: Optional[np.ndarray] = None
    self._orig_median: float = 0.0

  @staticmethod
  def _std_of_good_half(unique: np.ndarray, threshold: float) -> float:
    good = unique[unique >= threshold]
    std = np.sqrt(((good - threshold) ** 2).sum() / good.shape[0])
    if std > 0:
      return float(std)
    std = np.sqrt(((unique - threshold) ** 2).sum() / unique.shape[0])
    if np.isfinite(std):
      return float(std)
    return float(np.abs(unique - threshold).sum() / unique.shape[0])

  def warp(self, labels: np.ndarray) -> np.ndarray:
    labels = _validate_labels(labels)
    if labels.size == 1:
      return labels
    y = labels.flatten()
    median = np.nanmedian(y)
    finite = np.isfinite(y)
    unique, unique_idx = np.unique(y[finite], return_index=True)
    ranks = stats.rankdata(y, method='dense', nan_policy='omit')
    med_idx = unique.searchsorted(median, 'left')
    denom = med_idx + (unique[med_idx] == median) * 0.5
    std = self._std_of_good_half(unique, median)
    below = finite & (y < median)
    if below.any():
      q = 0.5 * (ranks[below] - 0.5) / denom
      y[below] = stats.norm.ppf(q) * std + median
    self._orig = unique
    self._warped = y[finite][unique_idx]
    self._orig_median = float(unique[len(unique) // 2])
    return y[:, None]

  def _unwarp_one(self, label: float) -> float:
    orig, warped = self._orig, self._warped
    if label >= self._orig_median:
      return label
    idx = np.searchsorted(warped, label)
    cand = warped[max(0, idx - 1):min(len(warped), idx + 1)]
    best = int(np.argmin(np.abs(cand - label)))
    if np.isclose(warped[best], label):
      return float(orig[best])
    if label < np.min(warped):
      return float(orig[0] - (np.abs(label - warped[0]) / (warped[-1] - warped[0])) * (orig[-1] - orig[0]))
    lower = np.searchsorted(warped, label) - 1
    upper = lower + 1
    return float(orig[lower] + (label - warped[lower]) * (orig[upper] - orig[lower]) / (warped[upper] - warped[lower]))

  def unwarp(self, labels: np.ndarray) -> np.ndarray:
    if self._orig is None:
      raise ValueError('warp() needs to be called before unwarp() is called.')
    y = _validate_labels(labels).flatten()
    if np.isnan(y).any():
      raise ValueError('unwarp does not support nan values.')
    return self._unwarp_many(y)[:, None]

  def _unwarp_many(self, y: np.ndarray) -> np.ndarray:
    """`_unwarp_one` for a whole array at once (`sample()` unwarps num_samples x num_trials values): the same
    branches in the same order, evaluated with masks; `tests/test_output_warpers.py` pins it to the scalar form."""
    orig, warped = self._orig, self._warped
    nw = len(warped)
    out = y.astype(np.float64).copy()
    todo = y < self._orig_median
    if not todo.any():
      return out
    lab = y[todo]
    idx = np.searchsorted(warped, lab)
    lo_c = np.maximum(0, idx - 1)
    hi_c = np.minimum(nw, idx + 1)                      # candidates warped[lo_c:hi_c], one or two values
    d0 = np.abs(warped[lo_c] - lab)
    has2 = hi_c - lo_c > 1
    d1 = np.where(has2, np.abs(warped[np.minimum(lo_c + 1, nw - 1)] - lab), np.inf)
    best = np.where(d1 < d0, 1, 0)                      # argmin inside the candidate window ...
    best = np.minimum(best, nw - 1)
    close = np.isclose(warped[best], lab)               # ... used, as in the reference, as an index into `warped`
    res = np.empty_like(lab)
    res[close] = orig[best[close]]
    below = ~close & (lab < np.min(warped))
    if below.any():
      res[below] = orig[0] - (np.abs(lab[below] - warped[0]) / (warped[-1] - warped[0])) * (orig[-1] - orig[0])
    mid = ~close & ~below
    if mid.any():
      lower = np.searchsorted(warped, lab[mid]) - 1
      upper = lower + 1
      res[mid] = orig[lower] + (lab[mid] - warped[lower]) * (orig[upper] - orig[lower]) / (warped[upper] - warped[lower])
    out[todo] = res
    return out


class LogWarperComponent:
  def __init__(self, offset: float = 1.5):
    if offset <= 0:
      raise ValueError('offset must be positive')
    self.offset = offset
    self._min: Optional[float] = None
    self._max: Optional[float] = None

  def warp(self, labels: np.ndarray) -> np.ndarray:
    labels = _validate_labels(labels)
    self._min, self._max = np.nanmin(labels), np.nanmax(labels)
    y = labels.flatten()
    f = np.isfinite(y)
    norm_diff = (self._max - y[f]) / (self._max - self._min)
    y[f] = 0.5 - np.log1p(norm_diff * (self.offset - 1)) / np.log(self.offset)
    return y[:, None]

  def unwarp(self, labels: np.ndarray) -> np.ndarray:
    if self._max is None:
      raise ValueError('warp() needs to be called before unwarp() is called.')
    y = np.asarray(labels, dtype=float).flatten()
    y = self._max - (np.exp(np.log(self.offset) * (0.5 - y)) - 1) * (self._max - self._min) / (self.offset - 1)
    return y[:, None]


class InfeasibleWarperComponent:
  def __init__(self):
    self._shift: Optional[float] = None

  def warp(self, labels: np.ndarray) -> np.ndarray:
    y = _validate_labels(labels).flatten()
    if np.isnan(y).all():
      self._shift = np.nan
      y[:] = 0
      return y[:, None]
    rng = np.nanmax(y) - np.nanmin(y)
    bad_value = np.nanmin(y) - (0.5 * rng + 1)
    n_feasible = y.size - np.isnan(y).sum()
    p_feasible = (0.5 + n_feasible) / (1 + y.size)
    self._shift = -np.nanmean(y) * p_feasible - bad_value * (1 - p_feasible)
    y[np.isnan(y)] = bad_value
    # (the reference shifts every entry, including the freshly imputed ones: :487-488)
    y = y + self._shift
    return y[:, None]

  def unwarp(self, labels: np.ndarray) -> np.ndarray:
    if self._shift is None:
      raise ValueError('warp() needs to be called before unwarp() is called.')
    return np.asarray(labels, dtype=float) - self._shift


class OutputWarperPipeline:
  def __init__(self, warpers: Optional[Sequence] = None):
    self.warpers: List = list(warpers or [])

  def warp(self, labels: np.ndarray) -> np.ndarray:
    y = _validate_labels(labels)
    if np.isfinite(y).all() and len(np.unique(y).flatten()) == 1:
      return np.zeros(y.shape)
    if np.isnan(y).all():
      return -1 * np.ones(shape=y.shape)
    for w in self.warpers:
      y = w.warp(y)
    return y

  def unwarp(self, labels: np.ndarray) -> np.ndarray:
    y = _validate_labels(labels)
    if np.isfinite(y).all() and len(np.unique(y).flatten()) == 1:
      u = np.unique(y).item()
      if u == 0.0:
        return y
      if u == -1.0:
        return np.nan * np.ones(shape=y.shape)
    for w in self.warpers[::-1]:
      y = w.unwarp(y)
    return y


def create_default_warper(*, half_rank_warp: bool = True, log_warp: bool = True, infeasible_warp: bool = True) -> OutputWarperPipeline:
  if not (half_rank_warp or log_warp or infeasible_warp):
    raise ValueError('At least one of "half_rank_warp", "log_warp" or "infeasible_warp" must be True.')
  ws = []
  if half_rank_warp:
    ws.append(HalfRankComponent())
  if log_warp:
    ws.append(LogWarperComponent())
  if infeasible_warp:
    ws.append(InfeasibleWarperComponent())
  return OutputWarperPipeline(ws)
