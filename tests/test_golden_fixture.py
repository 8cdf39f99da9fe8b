"""Product host code and oracle against the committed golden fixture
(tests/golden/reference_known_answers.json: known answers transcribed from the reference's tests)."""
import json
import os

import numpy as np

from oracle import eagle_oracle as eo
from oracle import gp_oracle as go
from vizier_b200 import acquisitions as acq_lib
from vizier_b200 import output_warpers as ow

G = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_known_answers.json')))


def _arr(v):
  return np.array([float(x) for x in v], dtype=float)


def test_ucb_and_trust_region_fixture():
  u = G['ucb']
  np.testing.assert_allclose(go.ucb(np.array([u['mean']]), np.array([u['stddev']]), u['coefficient']), [u['expected']])
  for key in ('trust_region_small', 'trust_region_bigger'):
    t = G[key]
    assert abs(go.trust_radius(t['n_obs'], t['continuous_dof'], t['categorical_dof']) - t['radius']) < 1e-3
    assert abs(acq_lib.trust_radius(t['n_obs'], t['continuous_dof'], t['categorical_dof']) - t['radius']) < 1e-3
  t = G['trust_region_small']
  np.testing.assert_allclose(go.min_linf_distance(np.array(t['xs']), np.array(t['trusted']), np.ones(2, bool)), t['distances'], atol=1e-15)
  t = G['trust_region_sparse']
  for lib in (go, acq_lib):
    assert lib.trust_region_dim_mask([np.array(f) for f in t['feasible']]).tolist() == t['mask']
  np.testing.assert_allclose(go.min_linf_distance(np.array(t['xs']), np.array(t['trusted']), np.array(t['mask'])), t['distances'], atol=1e-15)
  assert abs(acq_lib.trust_radius(2, 1, 0) - t['radius']) < 1e-3


def test_warper_fixture():
  c = G['default_warper_case1']
  np.testing.assert_allclose(ow.create_default_warper().warp(_arr(c['unwarped'])[:, None])[:, 0], _arr(c['expected']), rtol=1e-7)
  for key in ('half_rank_case1', 'half_rank_case2', 'half_rank_case3'):
    c = G[key]
    np.testing.assert_allclose(ow.HalfRankComponent().warp(_arr(c['unwarped'])[:, None])[:, 0], _arr(c['expected']))


def test_eagle_and_philox_fixture():
  c = G['eagle_update']
  st = eo.EagleState(2, np.array(c['pool'], float), np.array(c['rewards'], float), 4.0, np.ones(4))
  new = eo.update(st, 2, np.array(c['batch'], float), np.array(c['batch_rewards'], float), np.zeros((2, 2)), eo.EagleConfig())
  np.testing.assert_array_equal(new.features[:2], c['new_features'])
  np.testing.assert_array_equal(new.rewards[:2], c['new_rewards'])
  np.testing.assert_allclose(new.perturbations[:2], c['new_perturbations_factor'])
  m = G['mask_flip']
  _, fr = eo.mask_flip(np.zeros((5, 1)), _arr(m['rewards']))
  np.testing.assert_array_equal(fr, _arr(m['flipped']))
  for case in G['philox4x32_10']['cases']:
    h = lambda v: int(v, 16) if isinstance(v, str) else int(v)
    out = eo.philox4x32(np.array([[h(v) for v in case['ctr']]], dtype=np.uint32), np.array([h(v) for v in case['key']], dtype=np.uint32))[0]
    assert [int(v) for v in out] == [int(v, 16) for v in case['out']]
