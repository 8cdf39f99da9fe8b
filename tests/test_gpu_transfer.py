"""Transfer learning: stacked residual GPs (`VizierGPBandit.set_priors`, gp_bandit.py:289-318; gp/gp_models.py:91-140,
:245-300; gp/transfer_learning.py) through `vzgp_score_stack` / `vzgp_eagle_run_stack` against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason='no CUDA device')]

from oracle import eagle_oracle as eo  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402


def _f(x, shift=0.0):
  return -np.sum((x - 0.3 - shift) ** 2, axis=1)


def _build(sizes, d, seed, dk=0):
  """A stack with fixed hyper-parameters per level on both sides; returns (StackedGP, oracle predictives, top x)."""
  from vizier_b200 import gp
  rng = np.random.default_rng(seed)
  stack = gp.StackedGP(0)
  preds = []
  x = None
  for e, n in enumerate(sizes):
    x = rng.uniform(size=(n, d))
    z = rng.integers(0, 3, size=(n, dk)).astype(np.int32) if dk else None
    y = _f(x, 0.05 * e) + 0.03 * rng.normal(size=n) - (0.2 * (z[:, 0] == 1) if dk else 0.0)
    ls2 = (0.3 + 0.2 * e) * (1 + np.arange(d) / d)
    lk = np.linspace(0.7, 1.2, dk) if dk else None
    po = go.GPParams(0.8 + 0.3 * e, ls2, 1e-3 * (e + 1), lk)
    pg = gp.GPHyperParams(0.8 + 0.3 * e, ls2, 1e-3 * (e + 1), lk)
    resid_o = go.stack_residual_labels(preds, x, y, z)
    resid_d = y - stack.mean(x, z)
    np.testing.assert_allclose(resid_d, resid_o, atol=1e-9)
    preds.append(go.precompute_predictive(po, x, resid_o, z))
    level = stack.new_level()
    level.fit(x, resid_d, pg, z=z)
    stack.push(level, n)
  return stack, preds, x


@pytest.mark.parametrize('sizes,d,dk,radius', [((80, 40), 4, 0, 0.3), ((120, 60, 25), 6, 0, 0.9), ((50, 30), 3, 2, 0.25)])
def test_stack_score_matches_oracle(sizes, d, dk, radius):
  from vizier_b200 import gp
  stack, preds, x_top = _build(sizes, d, 17, dk)
  rng = np.random.default_rng(3)
  m = 700
  xs = rng.uniform(size=(m, d))
  zs = rng.integers(0, 3, size=(m, dk)).astype(np.int32) if dk else None
  mu, sd = go.predict_stack(preds, xs, zs)
  for e in range(1, len(sizes)):
    np.testing.assert_allclose(stack.alphas[e], go.transfer_alpha(sizes[e], sizes[e - 1], d + dk + 2), rtol=1e-14)
  mask = np.ones(d, bool)
  dist = go.min_linf_distance(xs, x_top, mask, np.ones(sizes[-1], bool))
  want = go.apply_trust_region(go.ucb(mu, sd, 1.8), dist, radius)
  out = stack.score(xs, gp.Acquisition(1.8, True, radius), zs=zs, with_aux=True)
  stack.synchronize()
  np.testing.assert_allclose(out['mean'].cpu().numpy(), mu, atol=1e-9)
  np.testing.assert_allclose(out['stddev'].cpu().numpy(), sd, atol=1e-9)
  np.testing.assert_array_equal(out['linf_distance'].cpu().numpy(), dist)
  np.testing.assert_allclose(out['score'].cpu().numpy(), want, atol=1e-9)
  stack.close()


def test_stack_eagle_and_random_search_match_oracle():
  from vizier_b200 import _lib, gp
  d, pool, batch, steps = 4, 50, 25, 8
  stack, preds, x_top = _build((90, 35), d, 23)
  radius = go.trust_radius(35, d, 0)
  mask = np.ones(d, bool)

  def score_fn(q):
    mu, sd = go.predict_stack(preds, q)
    dist = go.min_linf_distance(q, x_top, mask, np.ones(35, bool))
    return go.apply_trust_region(go.ucb(mu, sd, 1.8), dist, radius)

  cfg_o = eo.EagleConfig()
  wx, wr, _ = eo.run_eagle_optimizer(score_fn, dim=d, pool_size=pool, batch_size=batch, max_evaluations=steps * batch,
                                     count=2, seed=5, cfg=cfg_o, prior_features=x_top)
  cfg = _lib.EagleConfig(cfg_o.visibility, cfg_o.gravity, cfg_o.negative_gravity, cfg_o.perturbation,
                         cfg_o.perturbation_lower_bound, cfg_o.penalize_factor, cfg_o.normalization_scale,
                         cfg_o.prior_trials_pool_pct, pool, batch, steps * batch)
  acq = gp.Acquisition(1.8, True, radius)
  bx, _, br = stack.eagle_run(cfg, acq, count=2, seed=5, prior=x_top)
  np.testing.assert_allclose(br, wr, atol=1e-9)
  np.testing.assert_allclose(bx, wx, atol=1e-9)
  rx, _, rs, _ = stack.random_search(3000, acq, 2, seed=9)
  wrx, wrs, _ = eo.run_random_optimizer(score_fn, dim=d, num_candidates=3000, count=2, seed=9)
  np.testing.assert_allclose(rs, wrs, atol=1e-9)
  np.testing.assert_allclose(rx, wrx, atol=0)
  stack.close()


def test_designer_with_priors_predicts_better_from_few_trials():
  """`set_priors`: a well-sampled prior study of the same objective; with 6 current trials the stacked prediction is
  much closer to the truth than the prediction without priors, and suggest() keeps working (gp_bandit_test's
  transfer-learning shape: priors, then update, suggest, predict)."""
  from vizier_b200 import vz
  from vizier_b200.designers import gp_bandit
  from vizier_b200 import optimizers as vb
  d = 3
  p = vz.ProblemStatement()
  for i in range(d):
    p.search_space.root.add_float_param(f'x{i}', 0.0, 1.0)
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  rng = np.random.default_rng(0)

  def make(n, start_id=1):
    out = []
    for i in range(n):
      xv = rng.uniform(size=d)
      t = vz.Trial(parameters={f'x{j}': float(xv[j]) for j in range(d)}, id=start_id + i)
      t.complete(vz.Measurement({'obj': float(_f(xv[None, :])[0])}))
      out.append(t)
    return out

  prior, current, test = make(120), make(6, 1000), make(40, 2000)
  truth = np.array([t.final_measurement.metrics['obj'].value for t in test])
  fac = vb.VectorizedOptimizerFactory(strategy_factory=vb.VectorizedEagleStrategyFactory(), max_evaluations=2500,
                                      suggestion_batch_size=25)
  errs = {}
  for with_prior in (False, True):
    des = gp_bandit.VizierGPBandit(p, acquisition_optimizer_factory=fac, rng=1)
    if with_prior:
      des.set_priors([vz.CompletedTrials(prior)])
    des.update(vz.CompletedTrials(current), vz.ActiveTrials())
    sugg = des.suggest(2)
    assert len(sugg) == 2
    pred = des.predict(test, rng=3, num_samples=400)
    errs[with_prior] = float(np.sqrt(np.mean((pred.mean - truth) ** 2)))
    assert np.all(pred.stddev > 0)
  assert errs[True] < 0.6 * errs[False], errs


def _lambda_search(f, num_trials, seed):
  """gp_bandit_test.py:64-105 (_setup_lambda_search): 1-D problem on [-5, 5), `num_trials` evaluated trials of f."""
  from vizier_b200 import vz
  from vizier_b200.designers import gp_bandit
  from vizier_b200 import optimizers as vb
  p = vz.ProblemStatement()
  p.search_space.root.add_float_param('x0', -5.0, 5.0)
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  xs = np.random.default_rng(seed).uniform(-5.0, 5.0, num_trials)
  trials = []
  for i, xv in enumerate(xs):
    t = vz.Trial(parameters={'x0': float(xv)}, id=i + 1)
    t.complete(vz.Measurement({'obj': float(f(xv))}))
    trials.append(t)
  fac = vb.VectorizedOptimizerFactory(strategy_factory=vb.VectorizedEagleStrategyFactory(), max_evaluations=1000,
                                      suggestion_batch_size=25)
  return gp_bandit.VizierGPBandit(p, acquisition_optimizer_factory=fac, rng=1), trials, p


@pytest.mark.xfail(strict=False, reason="skipped upstream too ('The current transfer learning seems broken and test failing', "
                   "gp_bandit_test.py:546-547): each study's labels are warped on their own scale, so the residuals of "
                   "a linearly transformed study against the prior are not small; this port reproduces that behaviour "
                   "(measured: MSE 0.139 with the prior against 0.004 without)")
def test_prior_warping_like_the_reference_test():
  """gp_bandit_test.py:550-595 (test_prior_warping): the prior study samples f, the current study the linearly
  transformed 3 f + 10; upstream expects the designer with the prior to predict the transformed function better."""
  from vizier_b200 import vz
  f = lambda x: -((x / 5.0 - 0.5) ** 2)
  g = lambda x: 3.0 * f(x) + 10.0
  x_test = np.random.default_rng(1).uniform(-5.0, 5.0, 100)
  y_test = np.array([g(x) for x in x_test])
  test_trials = [vz.Trial(parameters={'x0': float(x)}) for x in x_test]
  with_prior, prior_trials, _ = _lambda_search(f, 100, 11)
  with_prior.set_priors([vz.CompletedTrials(prior_trials)])
  no_prior, obs, _ = _lambda_search(g, 20, 12)
  no_prior.update(vz.CompletedTrials(obs), vz.ActiveTrials())
  with_prior.update(vz.CompletedTrials(obs), vz.ActiveTrials())
  mse = lambda d: float(np.mean((d.predict(test_trials, rng=2, num_samples=300).mean - y_test) ** 2))
  assert mse(with_prior) < mse(no_prior)


@pytest.mark.parametrize('iters,batch_size', [(3, 5), (5, 1)])
def test_run_with_priors_like_the_reference_test(iters, batch_size):
  """gp_bandit_test.py:597-621 (test_run_with_priors): a designer with a prior study survives a suggest / complete /
  update loop with random metrics; every suggestion lies in the search space."""
  from vizier_b200 import vz
  des, prior_trials, p = _lambda_search(lambda x: -((x / 5.0 - 0.5) ** 2), 100, 21)
  des.set_priors([vz.CompletedTrials(prior_trials)])
  rng = np.random.default_rng(1)
  n, tid = 0, 1000
  for _ in range(iters):
    sugg = des.suggest(batch_size)
    assert len(sugg) == batch_size
    done = []
    for s in sugg:
      assert -5.0 <= s.parameters['x0'].value <= 5.0
      tid += 1
      t = s.to_trial(tid)
      t.complete(vz.Measurement({'obj': float(rng.uniform(-1.0, 1.0))}))
      done.append(t)
    n += len(done)
    des.update(vz.CompletedTrials(done), vz.ActiveTrials())
  assert n == iters * batch_size
