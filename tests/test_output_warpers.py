"""Host label warpers against the reference's golden arrays
(vizier/_src/algorithms/designers/gp/output_warpers_test.py:131-157 default pipeline,
:262-328 HalfRankComponent; values transcribed)."""
import numpy as np
import pytest

from vizier_b200 import output_warpers as ow


def test_default_pipeline_known_arrays():
  w = ow.create_default_warper()
  x = np.array([[1.0], [1.0], [5.0], [-1e80], [np.nan], [-np.inf]])
  want = np.array([[0.61848423], [0.61848423], [1.25966537], [0.25966537], [-1.24033463], [-1.24033463]])
  np.testing.assert_allclose(w.warp(x), want, rtol=1e-7)
  np.testing.assert_array_equal(w.warp(np.array([[np.nan], [np.nan]])), [[-1.0], [-1.0]])
  for c in (0.0, 1.0, 100.0, -100.0):
    np.testing.assert_array_equal(w.warp(c * np.ones((5, 1))), 0.0)
  with pytest.raises(ValueError):
    ow.create_default_warper(half_rank_warp=False, log_warp=False, infeasible_warp=False)


@pytest.mark.parametrize('unwarped,expected', [
    ([np.nan, 1, 4, 2, 10, 12, -np.inf, 2, 3, 5, 6],
     [np.nan, -2.7145447657886415, 4.0, 0.3722561569665319, 10.0, 12.0, np.nan, 0.3722561569665319,
      2.322289907556879, 5.0, 6.0]),
    ([np.nan, -4, -3, -2, 1.1, 1.2, 1.3, 1.4, 1.5],
     [np.nan, 0.7984888240158797, 0.9467291870388195, 1.0380072549079085, 1.1139555940074284, 1.2, 1.3, 1.4, 1.5]),
    ([np.nan, 1, 2, 3, 4, 4, 6, 7, 10, 11, 12],
     [np.nan, -2.3573836671676096, 0.7453945664588675, 2.655910679724611, 4.2455644597926385,
      4.2455644597926385, 6.0, 7.0, 10.0, 11.0, 12.0]),
])
def test_half_rank_known_arrays(unwarped, expected):
  got = ow.HalfRankComponent().warp(np.array([unwarped], dtype=float).T)
  np.testing.assert_allclose(got, np.array([expected]).T)


def test_warp_preserves_rank_and_unwarp_inverts():
  rng = np.random.default_rng(0)
  y = rng.normal(size=(40, 1)) * 3 + 1
  w = ow.create_default_warper()
  z = w.warp(y)
  assert np.all(np.isfinite(z))
  np.testing.assert_array_equal(np.argsort(y[:, 0]), np.argsort(z[:, 0]))
  np.testing.assert_allclose(w.unwarp(z), y, rtol=1e-6, atol=1e-6)
  # input is not mutated
  y2 = y.copy(); w.warp(y); np.testing.assert_array_equal(y, y2)


def test_infeasible_are_worst_and_mean_shift():
  y = np.array([[1.0], [2.0], [np.nan], [4.0]])
  z = ow.InfeasibleWarperComponent().warp(y)
  assert z[2, 0] < z[[0, 1, 3], 0].min()


def test_vectorised_half_rank_unwarp_equals_the_scalar_form():
  from vizier_b200 import output_warpers as ow
  rng = np.random.default_rng(0)
  for n in (3, 10, 57):
    labels = rng.normal(size=(n, 1)) * 3
    labels[rng.integers(0, n, size=2)] = labels[0]          # ties
    w = ow.HalfRankComponent()
    warped = w.warp(labels.copy())
    q = np.concatenate([warped.flatten(), rng.normal(size=200) * 4, [warped.min() - 1.0, warped.max() + 1.0]])
    want = np.array([w._unwarp_one(v) for v in q])
    np.testing.assert_array_equal(w._unwarp_many(q), want)
    np.testing.assert_array_equal(w.unwarp(q[:, None]).flatten(), want)
