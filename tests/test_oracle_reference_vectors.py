"""Pins the oracle to the known-answer tests the REFERENCE holds for this path.

Sources (paths under /root/reference, values transcribed; nothing is read at run time):
  vizier/_src/algorithms/designers/gp/acquisitions_test.py:53-55   (UCB)
  vizier/_src/algorithms/designers/gp/acquisitions_test.py:349-545 (TrustRegion)
  vizier/_src/algorithms/optimizers/eagle_strategy_test.py:258-419 (update / trim)
  vizier/_src/algorithms/optimizers/eagle_strategy_test.py:39-134  (naive force restatement)
"""
import numpy as np

from oracle import eagle_oracle as eo
from oracle import gp_oracle as go


def test_ucb_known_answer():
  # acquisitions_test.py:53-55: UCB(coefficient=2.0) on Normal(0.1, 1) == 2.1
  np.testing.assert_allclose(go.ucb(np.array([0.1]), np.array([1.0]), 2.0), [2.1])


def test_trust_region_small():
  trusted = np.array([[0.0, 0.0], [1.0, 1.0]])
  xs = np.array([[0.0, 0.3], [0.9, 0.8], [1.0, 1.0]])
  mask = go.trust_region_dim_mask([np.array([]), np.array([])])
  np.testing.assert_allclose(go.min_linf_distance(xs, trusted, mask), [0.3, 0.2, 0.0], atol=1e-15)
  # 2 continuous + 2 categorical dof, 2 observations -> 0.224
  assert abs(go.trust_radius(2, int(mask.sum()), 2) - 0.224) < 1e-3


def test_trust_region_ignores_sparse_feasible_dimensions():
  trusted = np.array([[0.0, 0.0, 100.0], [0.6, 0.0, -120.0]])
  mask = go.trust_region_dim_mask([np.linspace(0.0, 1.0, 11), np.array([0.0, 1.0]), np.array([5.0])])
  assert mask.tolist() == [True, False, False]
  xs = np.array([[0.5, 1.0, 5.0], [0.2, 1.0, 5.0]])
  np.testing.assert_allclose(go.min_linf_distance(xs, trusted, mask), [0.1, 0.2], atol=1e-15)
  assert abs(go.trust_radius(2, int(mask.sum()), 0) - 0.26) < 1e-3
  # with_padding variant: padded trusted rows are excluded, padded query rows give 0 distance
  trusted_p = np.zeros((5, 3)); trusted_p[:2] = trusted
  valid = np.array([True, True, False, False, False])
  xs_p = np.zeros((5, 3)); xs_p[:2] = xs
  got = go.min_linf_distance(xs_p, trusted_p, mask, valid)
  np.testing.assert_allclose(got, [0.1, 0.2, 0.0, 0.0, 0.0], atol=1e-15)


def test_trust_region_bigger():
  trusted = np.vstack([[0.0, 0.0], [1.0, 1.0]] * 10)
  xs = np.array([[0.0, 0.3], [0.9, 0.8], [1.0, 1.0]])
  mask = np.array([True, True])
  np.testing.assert_allclose(go.min_linf_distance(xs, trusted, mask), [0.3, 0.2, 0.0], atol=1e-15)
  assert abs(go.trust_radius(20, 2, 2) - 0.44) < 1e-3


def test_trust_region_multi_batch_and_all_categorical():
  trusted = np.array([[0.0, 0.0], [1.0, 1.0]])
  xs = np.array([[[0.0, 0.3], [0.9, 0.8], [1.0, 1.0]], [[1.0, 1.0], [0.0, 0.3], [0.9, 0.8]]])
  got = go.min_linf_distance(xs, trusted, np.array([True, True]))
  np.testing.assert_allclose(got, [[0.3, 0.2, 0.0], [0.0, 0.3, 0.2]], atol=1e-15)
  # no continuous features -> -inf (acquisitions.py:815-816)
  got = go.min_linf_distance(np.zeros((3, 3, 0)), np.zeros((2, 0)), np.zeros((0,), bool))
  assert np.all(np.isneginf(got)) and got.shape == (3, 3)


def test_apply_trust_region():
  acq = np.array([1.0, 2.0, 3.0])
  dist = np.array([0.1, 0.3, 0.25])
  np.testing.assert_allclose(go.apply_trust_region(acq, dist, 0.25), [1.0, -1e4 - 0.3, 3.0])
  np.testing.assert_allclose(go.apply_trust_region(acq, dist, 0.6), acq)  # radius > 0.5: inactive


def _state(iterations=2):
  feats = np.array([[1, 2], [3, 4], [7, 7], [8, 8]], dtype=np.float64)
  rewards = np.array([2, 3, 4, 1], dtype=np.float64)
  return eo.EagleState(iterations, feats, rewards, float(rewards.max()), np.ones(4))


def test_eagle_update_pool_features_and_rewards():
  # eagle_strategy_test.py:258-309 (pool 4, batch 2; iteration past the init phase)
  cfg = eo.EagleConfig()
  st = _state(iterations=2)  # batch_id = 2 % 2 = 0 -> flies 0,1
  new = eo.update(st, 2, np.array([[9.0, 9.0], [10.0, 10.0]]), np.array([5.0, 0.5]),
                  np.full((2, 2), 0.123), cfg)
  np.testing.assert_array_equal(new.features[:2], [[9, 9], [3, 4]])
  np.testing.assert_array_equal(new.rewards[:2], [5, 3])
  np.testing.assert_allclose(new.perturbations[:2], [1, cfg.penalize_factor])
  np.testing.assert_array_equal(new.features[2:], st.features[2:])
  assert new.iterations == 3


def test_eagle_update_best_reward():
  # eagle_strategy_test.py:311-342
  st = _state(iterations=0)
  bf = np.array([[9.0, 9.0], [10.0, 10.0]])
  new = eo.update(st, 2, bf, np.array([5.0, 0.5]), np.zeros((2, 2)), eo.EagleConfig())
  assert new.best_reward == 5.0
  new2 = eo.update(new, 2, bf, np.array([2.0, 4.0]), np.zeros((2, 2)), eo.EagleConfig())
  assert new2.best_reward == 5.0


def test_eagle_trim_pool():
  # eagle_strategy_test.py:365-419: perturbation 0 < lower bound and reward != best -> replaced.
  cfg = eo.EagleConfig()
  st = eo.EagleState(2, np.array([[1.0, 2.0], [3.0, 4.0]]), np.array([2.0, 3.0]), 4.0,
                     np.array([cfg.perturbation, 0.0]))
  rnd = np.array([[0.11, 0.22], [0.33, 0.44]])
  # batch rewards lower than previous -> no improvement; fly 1 has perturbation 0 -> trimmed.
  new = eo.update(st, 2, np.array([[5.0, 5.0], [6.0, 6.0]]), np.array([-1.0, -1.0]), rnd, cfg)
  np.testing.assert_array_equal(new.features[0], [1.0, 2.0])
  np.testing.assert_array_equal(new.features[1], rnd[1])
  np.testing.assert_array_equal(new.rewards, [2.0, -np.inf])
  np.testing.assert_allclose(new.perturbations, [cfg.perturbation * cfg.penalize_factor, cfg.perturbation])
  # The best firefly is never removed.
  st2 = eo.EagleState(2, st.features.copy(), np.array([2.0, 4.0]), 4.0, np.array([cfg.perturbation, 0.0]))
  new2 = eo.update(st2, 2, np.array([[5.0, 5.0], [6.0, 6.0]]), np.array([-1.0, -1.0]), rnd, cfg)
  np.testing.assert_array_equal(new2.features[1], [3.0, 4.0])
  assert new2.rewards[1] == 4.0


def test_eagle_create_features_matches_naive_unnormalised_structure():
  # eagle_strategy_test.py:39-134 restates the force as a materialised sum_j scale_ij (p_j - b_i).
  rng = np.random.default_rng(0)
  pool = rng.uniform(size=(12, 5)); rewards = rng.normal(size=12); rewards[3] = -np.inf
  batch = pool[4:8]; rb = rewards[4:8]
  cfg = eo.EagleConfig()
  got = eo.create_features(pool, rewards, batch, rb, np.zeros((4, 5)), cfg)
  # naive: same normalised scale, materialised differences
  diffs = pool[None, :, :] - batch[:, None, :]
  d2 = np.sum(diffs**2, -1)
  with np.errstate(invalid='ignore'):
    direc = rewards[None, :] - rb[:, None]
  sd = np.where(direc >= 0, cfg.gravity, -cfg.negative_gravity)
  sc = np.isfinite(rewards)[None, :] * sd * np.exp(-cfg.visibility * d2 / 5 * 10.0)
  pull = np.maximum(sc, 0); push = np.minimum(sc, 0)
  npl = np.where(pull > 0, pull / np.maximum((pull > 0).sum(1, keepdims=True), 1), 0) * 0.5
  nps = np.where(push < 0, push / np.maximum((push < 0).sum(1, keepdims=True), 1), 0) * 0.5
  want = batch + np.sum(diffs * (npl + nps)[..., None], axis=1)
  np.testing.assert_allclose(got, want, atol=1e-14)


def test_mask_flip():
  # eagle_strategy.py:484-486 docstring example.
  f = np.arange(5, dtype=np.float64)[:, None]
  r = np.array([1, -np.inf, 3, -np.inf, 2])
  ff, fr = eo.mask_flip(f, r)
  np.testing.assert_array_equal(fr, [2, 3, 1, -np.inf, -np.inf])
  np.testing.assert_array_equal(ff[:3, 0], [4, 2, 0])


def test_default_pool_size():
  cfg = eo.EagleConfig()
  assert eo.default_pool_size(4, 25, cfg) == 25      # C1
  assert eo.default_pool_size(20, 25, cfg) == 75     # D=20
  assert eo.default_pool_size(100, 5, eo.EagleConfig(max_pool_size=50)) == 50
  assert eo.default_pool_size(100, None, eo.EagleConfig(max_pool_size=10)) == 10


def test_categorical_logits_match_reference_restatement():
  """eagle_strategy_test.py:38-71 (`_create_logits_vector_simple`, one-hot matmul form) vs the
  library form (`_create_logits_vector`, eagle_strategy.py:954-985) restated in the oracle."""
  rng = np.random.default_rng(1)
  cfg = eo.EagleConfig()
  p, b = 9, 4
  sizes = np.array([2, 3, 5])
  pool_z = np.stack([rng.integers(0, s, size=p) for s in sizes], axis=1)
  batch_z = pool_z[:b].copy()
  scale = rng.normal(size=(b, p)) * 0.1
  got = eo.categorical_logits(pool_z, batch_z, scale, sizes, cfg)
  max_size = 5
  for i, s in enumerate(sizes):
    oh_f = np.zeros((p, s)); oh_f[np.arange(p), pool_z[:, i]] = 1
    oh_b = np.zeros((b, s)); oh_b[np.arange(b), batch_z[:, i]] = 1
    change = scale @ oh_f - oh_b * np.sum(scale, axis=-1, keepdims=True)
    diff_logit = np.log((1.0 - cfg.prob_same_category_without_perturbation) / (s - 1))
    want = np.zeros((b, max_size)) + diff_logit
    want[:, s:] = -np.inf
    want[np.arange(b), batch_z[:, i]] = np.log(cfg.prob_same_category_without_perturbation)
    want[:, :s] = want[:, :s] + change
    np.testing.assert_allclose(got[:, i, :], want, atol=1e-14)


def test_noise_transforms():
  u = eo.philox_uniform(3, eo.STREAM_CAT_LAPLACE, 0, 200000)
  lap = eo.laplace_from_uniform(u)
  assert abs(lap.mean()) < 2e-2 and abs(lap.var() - 2.0) < 5e-2      # Laplace(0,1): var 2
  g = eo.gumbel_from_uniform(eo.philox_uniform(3, eo.STREAM_CAT_GUMBEL, 0, 200000))
  assert abs(g.mean() - 0.5772156649) < 1e-2
  z = eo.uniform_categories(u[:3000].reshape(1000, 3), np.array([2, 3, 7]))
  assert z.min() == 0 and z[:, 0].max() == 1 and z[:, 2].max() == 6
