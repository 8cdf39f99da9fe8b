"""Multi-metric path on the GPU against the oracle: the independent multi-task GP (one factor, M alphas;
tuned_gp_models.py:282-288), its NLL + gradient, the hyper-volume scalarised UCB (gp_bandit.py:214-242) and
the Eagle loop with that scorer; then the designer end to end."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason='no CUDA device')]

from oracle import eagle_oracle as eo  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402


def _setup(n, d, m, dk=0, seed=0):
  rng = np.random.default_rng(seed)
  x = rng.uniform(size=(n, d))
  y = np.stack([-np.sum((x - 0.3 - 0.2 * k) ** 2, axis=1) + 0.05 * rng.normal(size=n) for k in range(m)], axis=1)
  z = rng.integers(0, 3, size=(n, dk)).astype(np.int32) if dk else None
  ls2 = 0.5 * (1 + np.arange(d) / d)
  lk = np.linspace(0.7, 1.3, dk) if dk else None
  return rng, x, y, z, ls2, lk


@pytest.mark.parametrize('n,d,m,dk,nv', [(40, 3, 2, 0, 40), (130, 5, 3, 1, 120), (300, 6, 2, 0, 300)])
def test_multi_nll_grad_matches_oracle(n, d, m, dk, nv):
  from vizier_b200 import gp
  _, x, y, z, ls2, lk = _setup(n, d, m, dk, 1)
  po = go.GPParams(0.7, ls2, 2e-3, lk); pg = gp.GPHyperParams(0.7, ls2, 2e-3, lk)
  valid = np.arange(n) < nv
  want_l, want_g = go.loss_and_grad(po.to_vector(), x, y, z, valid)
  dev = gp.DeviceGP(0)
  for _ in range(2):   # capture, then graph replay
    loss, grad, retries = dev.loss_and_grad(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), pg,
                                            z=torch.from_numpy(z).cuda() if z is not None else None, n_valid=nv)
    assert retries == 0
    assert abs(loss - want_l) < 1e-9 * max(1.0, abs(want_l))
    np.testing.assert_allclose(grad, want_g, atol=1e-8 * max(1.0, np.max(np.abs(want_g))), rtol=0)


@pytest.mark.parametrize('n,d,m,dk,mc', [(60, 4, 2, 0, 300), (200, 6, 3, 2, 5000), (1000, 20, 2, 0, 20000)])
def test_scalarized_ucb_score_matches_oracle(n, d, m, dk, mc):
  from vizier_b200 import gp
  rng, x, y, z, ls2, lk = _setup(n, d, m, dk, 2)
  po = go.GPParams(1.0, ls2, 1e-3, lk); pg = gp.GPHyperParams(1.0, ls2, 1e-3, lk)
  dev = gp.DeviceGP(0)
  assert dev.fit(x, y, pg, z=z) == 0
  pred = go.precompute_predictive(po, x, y, z)
  xs = rng.uniform(size=(mc, d))
  zs = rng.integers(0, 3, size=(mc, dk)).astype(np.int32) if dk else None
  w = np.abs(rng.normal(size=(200, m))); w /= np.linalg.norm(w, axis=1, keepdims=True)
  ref = go.hv_reference_point(y)
  best = go.hv_max_scalarized(y, w, ref)
  mu, sd = go.predict(pred, xs, zs)
  for floor in (best, None):
    acq = gp.ScalarizedUcbAcquisition(w, ref, floor, 1.8)
    out = dev.score_multi(xs, acq, zs=zs, with_aux=True)
    dev.synchronize()
    want = go.scalarized_ucb(mu, sd, w, ref, floor)
    np.testing.assert_allclose(out['mean'].cpu().numpy().T, mu, atol=1e-10, rtol=0)
    np.testing.assert_allclose(out['stddev'].cpu().numpy(), sd, atol=1e-10, rtol=0)
    np.testing.assert_allclose(out['score'].cpu().numpy(), want, atol=1e-10 * max(1.0, np.max(np.abs(want))), rtol=0)
  pm, pc = dev.posterior(xs[:50], zs[:50] if zs is not None else None)
  np.testing.assert_allclose(pm.cpu().numpy().T, mu[:50], atol=1e-10, rtol=0)
  np.testing.assert_allclose(np.sqrt(np.diag(pc.cpu().numpy())), sd[:50], atol=1e-9, rtol=0)


def test_eagle_with_scalarized_ucb_matches_oracle():
  from vizier_b200 import _lib, gp
  n, d, m = 80, 4, 2
  rng, x, y, _, ls2, _ = _setup(n, d, m, 0, 3)
  po = go.GPParams(1.0, ls2, 1e-3); pg = gp.GPHyperParams(1.0, ls2, 1e-3)
  dev = gp.DeviceGP(0)
  dev.fit(x, y, pg)
  pred = go.precompute_predictive(po, x, y)
  w = np.abs(rng.normal(size=(64, m))); w /= np.linalg.norm(w, axis=1, keepdims=True)
  ref = go.hv_reference_point(y); best = go.hv_max_scalarized(y, w, ref)

  def score_fn(q):
    mu, sd = go.predict(pred, q)
    return go.scalarized_ucb(mu, sd, w, ref, best)

  cfg_o = eo.EagleConfig()
  pool, batch, steps = 25, 25, 7
  wx, wr, _ = eo.run_eagle_optimizer(score_fn, dim=d, pool_size=pool, batch_size=batch, max_evaluations=steps * batch,
                                     count=3, seed=11, cfg=cfg_o, prior_features=x)
  cfg = _lib.EagleConfig(cfg_o.visibility, cfg_o.gravity, cfg_o.negative_gravity, cfg_o.perturbation,
                         cfg_o.perturbation_lower_bound, cfg_o.penalize_factor, cfg_o.normalization_scale,
                         cfg_o.prior_trials_pool_pct, pool, batch, steps * batch)
  bx, _, br = dev.eagle_run(cfg, gp.ScalarizedUcbAcquisition(w, ref, best, 1.8), count=3, seed=11, prior=x)
  np.testing.assert_allclose(br, wr, atol=1e-9 * max(1.0, np.max(np.abs(wr))))
  np.testing.assert_allclose(bx, wx, atol=1e-9)


def test_multi_metric_designer_suggests_and_predicts():
  """The shape of gp_bandit_test.py's multi-objective cases: two metrics (one minimised), suggest / update
  rounds with both optimiser strategies, predictions of shape (n, num_metrics)."""
  from vizier_b200 import optimizers as vb
  from vizier_b200 import vz
  from vizier_b200.designers import gp_bandit
  p = vz.ProblemStatement()
  for i in range(3):
    p.search_space.root.add_float_param(f'x{i}', 0.0, 1.0)
  p.metric_information.append(vz.MetricInformation(name='gain', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  p.metric_information.append(vz.MetricInformation(name='cost', goal=vz.ObjectiveMetricGoal.MINIMIZE))
  rng = np.random.default_rng(0)

  def evaluate(params):
    x = np.array([params[f'x{i}'].value for i in range(3)])
    return {'gain': float(-np.sum((x - 0.7) ** 2)), 'cost': float(np.sum((x - 0.2) ** 2))}

  for factory in (vb.VectorizedOptimizerFactory(strategy_factory=vb.VectorizedEagleStrategyFactory(), max_evaluations=1500,
                                                suggestion_batch_size=25),
                  vb.VectorizedOptimizerFactory(strategy_factory=vb.random_strategy_factory, max_evaluations=5000,
                                                suggestion_batch_size=5000)):
    d = gp_bandit.VizierGPBandit(p, rng=5, acquisition_optimizer_factory=factory, num_scalarizations=100)
    trials = []
    for i in range(12):
      xv = rng.uniform(size=3)
      t = vz.Trial(parameters={f'x{j}': float(xv[j]) for j in range(3)}, id=i + 1)
      t.complete(vz.Measurement(evaluate(t.parameters)))
      trials.append(t)
    d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
    for step in range(3):
      sugg = d.suggest(2)
      assert len(sugg) == 2
      new = []
      for k, s in enumerate(sugg):
        assert p.search_space.contains(s.parameters)
        t = s.to_trial(100 + 10 * step + k)
        t.complete(vz.Measurement(evaluate(t.parameters)))
        new.append(t)
      d.update(vz.CompletedTrials(new), vz.ActiveTrials())
    pred = d.predict(trials[:5], rng=1, num_samples=200)
    assert pred.mean.shape == (5, 2) and pred.stddev.shape == (5, 2)
    # predictions live in the model's label space: MINIMIZE metrics are sign-flipped by the converter
    # (converters/core.py:539-737), exactly like the single-metric path
    truth = np.array([[evaluate(t.parameters)['gain'], -evaluate(t.parameters)['cost']] for t in trials[:5]])
    assert np.max(np.abs(pred.mean - truth)) < 0.4
