"""The `linear_coef` model (Matern + feature-scaled linear kernel, constant mean; tuned_gp_models.py:203-245)
on the GPU against the oracle: kernel matrices, NLL + the 3 extra gradient entries, fit, scoring through the
general path (explicit K*), Eagle, posterior, and the designer."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason='no CUDA device')]

from oracle import eagle_oracle as eo  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402

COEF = 0.7


def _gp():
  from vizier_b200 import gp
  return gp


def _setup(n, d, dk=0, seed=0):
  rng = np.random.default_rng(seed)
  x = rng.uniform(size=(n, d))
  y = 1.5 * x[:, 0] - np.sum((x - 0.4) ** 2, axis=1) + 0.05 * rng.normal(size=n) + 0.8
  z = rng.integers(0, 3, size=(n, dk)).astype(np.int32) if dk else None
  ls2 = 0.5 * (1 + np.arange(d) / d)
  lk = np.linspace(0.7, 1.3, dk) if dk else None
  return rng, x, y, z, ls2, lk


def _params(ls2, lk, sn2=2e-3):
  from vizier_b200 import gp
  po = go.GPParams(0.8, ls2, sn2, lk, go.LinearParams(COEF, 0.9, 0.3, -0.4))
  pg = gp.GPHyperParams(0.8, ls2, sn2, lk, COEF, 0.9, 0.3, -0.4)
  return po, pg


def test_linear_kernel_matrices():
  from vizier_b200 import gp
  _, x, _, z, ls2, lk = _setup(150, 5, 2, 1)
  po, pg = _params(ls2, lk)
  dev = gp.DeviceGP(0)
  want = go.kernel_matrix(po, x, z, row_valid=np.arange(150) < 140)
  got = dev.kernel_matrix(x, pg, z=z, n_valid=140, diag_add=po.observation_noise_variance).cpu().numpy()
  np.testing.assert_allclose(got, want, atol=1e-12, rtol=0)
  xs = np.random.default_rng(2).uniform(size=(70, 5)); zs = np.random.default_rng(3).integers(0, 3, size=(70, 2)).astype(np.int32)
  np.testing.assert_allclose(dev.cross_kernel(xs, x, pg, zs=zs, z=z).cpu().numpy(), go.kernel(po, xs, x, zs, z), atol=1e-12, rtol=0)


@pytest.mark.parametrize('n,d,dk,nv', [(40, 3, 0, 40), (130, 6, 2, 120), (300, 20, 0, 300)])
def test_linear_nll_grad(n, d, dk, nv):
  from vizier_b200 import gp
  _, x, y, z, ls2, lk = _setup(n, d, dk, 4)
  po, pg = _params(ls2, lk)
  valid = np.arange(n) < nv
  want_l, want_g = go.loss_and_grad(po.to_vector(), x, y, z, valid, linear_coef=COEF)
  dev = gp.DeviceGP(0)
  loss, grad, retries = dev.loss_and_grad(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), pg,
                                          z=torch.from_numpy(z).cuda() if z is not None else None, n_valid=nv)
  assert retries == 0 and grad.shape == want_g.shape
  assert abs(loss - want_l) < 1e-9 * max(1.0, abs(want_l))
  np.testing.assert_allclose(grad, want_g, atol=1e-8 * max(1.0, np.max(np.abs(want_g))), rtol=0)
  f = dev.make_loss_fn(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(),
                       torch.from_numpy(z).cuda() if z is not None else None, nv, linear_coef=COEF)
  l2, g2 = f(pg.to_vector())
  assert l2 == loss
  np.testing.assert_array_equal(g2, grad)


@pytest.mark.parametrize('n,d,dk,m,radius', [(60, 4, 0, 300, None), (200, 6, 2, 5000, None), (50, 3, 0, 700, 0.3)])
def test_linear_score_fit_posterior(n, d, dk, m, radius):
  from vizier_b200 import gp
  rng, x, y, z, ls2, lk = _setup(n, d, dk, 5)
  po, pg = _params(ls2, lk, 1e-3)
  dev = gp.DeviceGP(0)
  assert dev.fit(x, y, pg, z=z) == 0
  pred = go.precompute_predictive(po, x, y, z)
  np.testing.assert_allclose(dev.cholesky().cpu().numpy(), pred.chol, atol=1e-11, rtol=0)
  xs = rng.uniform(size=(m, d))
  zs = rng.integers(0, 3, size=(m, dk)).astype(np.int32) if dk else None
  mu, sd = go.predict(pred, xs, zs)
  r = go.trust_radius(n, d, dk) if radius is None else radius
  out = dev.score(xs, gp.Acquisition(1.8, True, r), zs=zs, with_aux=True)
  dev.synchronize()
  np.testing.assert_allclose(out['mean'].cpu().numpy(), mu, atol=1e-10, rtol=0)
  np.testing.assert_allclose(out['stddev'].cpu().numpy(), sd, atol=1e-10, rtol=0)
  dist = go.min_linf_distance(xs, x, np.ones(d, bool), np.ones(n, bool))
  np.testing.assert_array_equal(out['linf_distance'].cpu().numpy(), dist)
  want = go.apply_trust_region(mu + 1.8 * sd, dist, r)
  np.testing.assert_allclose(out['score'].cpu().numpy(), want, atol=1e-10, rtol=0)
  pm, pc = dev.posterior(xs[:40], zs[:40] if zs is not None else None)
  np.testing.assert_allclose(pm.cpu().numpy(), mu[:40], atol=1e-10, rtol=0)
  np.testing.assert_allclose(np.sqrt(np.diag(pc.cpu().numpy())), sd[:40], atol=1e-9, rtol=0)


def test_linear_eagle_matches_oracle():
  from vizier_b200 import _lib, gp
  n, d = 50, 4          # N <= 64: would take the fused single-CTA loop if the model were plain Matern
  _, x, y, _, ls2, _ = _setup(n, d, 0, 6)
  po, pg = _params(ls2, None, 1e-3)
  dev = gp.DeviceGP(0)
  dev.fit(x, y, pg)
  pred = go.precompute_predictive(po, x, y)
  radius = go.trust_radius(n, d, 0)

  def score_fn(q):
    mu, sd = go.predict(pred, q)
    return go.apply_trust_region(mu + 1.8 * sd, go.min_linf_distance(q, x, np.ones(d, bool), np.ones(n, bool)), radius)

  cfg_o = eo.EagleConfig()
  wx, wr, _ = eo.run_eagle_optimizer(score_fn, dim=d, pool_size=25, batch_size=25, max_evaluations=150, count=3, seed=9,
                                     cfg=cfg_o, prior_features=x)
  cfg = _lib.EagleConfig(cfg_o.visibility, cfg_o.gravity, cfg_o.negative_gravity, cfg_o.perturbation,
                         cfg_o.perturbation_lower_bound, cfg_o.penalize_factor, cfg_o.normalization_scale,
                         cfg_o.prior_trials_pool_pct, 25, 25, 150)
  bx, _, br = dev.eagle_run(cfg, gp.Acquisition(1.8, True, radius), count=3, seed=9, prior=x)
  np.testing.assert_allclose(br, wr, atol=1e-9)
  np.testing.assert_allclose(bx, wx, atol=1e-9)


def test_linear_designer_runs_and_predicts():
  """gp_bandit_test.py's linear_coef cases in shape: suggest / update rounds, predictions near the data."""
  from vizier_b200 import optimizers as vb
  from vizier_b200 import vz
  from vizier_b200.designers import gp_bandit
  p = vz.ProblemStatement()
  for i in range(3):
    p.search_space.root.add_float_param(f'x{i}', 0.0, 1.0)
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  rng = np.random.default_rng(0)
  f = lambda xv: float(2.0 * xv[0] - np.sum((xv - 0.5) ** 2))
  fac = vb.VectorizedOptimizerFactory(strategy_factory=vb.VectorizedEagleStrategyFactory(), max_evaluations=1500, suggestion_batch_size=25)
  d = gp_bandit.VizierGPBandit(p, rng=3, linear_coef=1.0, acquisition_optimizer_factory=fac)
  trials = []
  for i in range(80):            # > 64 trials: the general (multi-kernel) NLL path
    xv = rng.uniform(size=3)
    t = vz.Trial(parameters={f'x{j}': float(xv[j]) for j in range(3)}, id=i + 1)
    t.complete(vz.Measurement({'obj': f(xv)}))
    trials.append(t)
  d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  for step in range(2):
    sugg = d.suggest(2)
    assert len(sugg) == 2 and all(p.search_space.contains(s.parameters) for s in sugg)
    new = []
    for k, s in enumerate(sugg):
      xv = np.array([s.parameters[f'x{j}'].value for j in range(3)])
      t = s.to_trial(200 + 10 * step + k)
      t.complete(vz.Measurement({'obj': f(xv)}))
      new.append(t)
    d.update(vz.CompletedTrials(new), vz.ActiveTrials())
  assert d._last_params.linear_coef == 1.0
  pred = d.predict(trials[:6], rng=1, num_samples=300)
  truth = np.array([t.final_measurement.metrics['obj'].value for t in trials[:6]])
  assert np.max(np.abs(pred.mean - truth)) < 0.15


def test_ucb_pe_with_linear_kernel():
  """GP-UCB-PE `mixes_linear_kernel=True` (gp_ucb_pe.py:805-809, :844-853): the PE / UCB acquisition over two models
  with the Matern + linear kernel and constant mean against the oracle, and the designer end to end."""
  from vizier_b200 import vz
  from vizier_b200.designers import gp_ucb_pe
  from vizier_b200 import optimizers as vb
  gp = _gp()
  rng = np.random.default_rng(4)
  n, n_pend, d, m = 60, 5, 4, 300
  x = rng.uniform(size=(n + n_pend, d))
  y = 2.0 * x[:n, 0] - np.sum((x[:n] - 0.4) ** 2, axis=1) + 0.02 * rng.normal(size=n)
  ls2 = 0.4 * (1 + np.arange(d) / d)
  lin = go.LinearParams(1.0, 0.7, 0.2, 0.3)
  po = go.GPParams(0.9, ls2, 2e-3, None, lin)
  pg = gp.GPHyperParams(0.9, ls2, 2e-3, None, linear_coef=1.0, linear_slope_amplitude=0.7, linear_shift=0.2, mean_constant=0.3)
  pred_a = go.precompute_predictive(po, x[:n], y)
  y_all = np.r_[y, np.zeros(n_pend)]
  pred_b = go.precompute_predictive(po, x, y_all)
  dev_a = gp.DeviceGP(0)
  dev_b = gp.DeviceGP(0, stream=dev_a.stream)
  dev_a.fit(x[:n], y, pg)
  dev_b.fit(x, y_all, pg)
  xs = rng.uniform(size=(m, d))
  for mode in (0, 1):
    pe = gp.UcbPeAcquisition(mode=mode, threshold=0.1, use_trust_region=True, trust_radius=0.3, tr_rows=n + 2)
    out = dev_a.score_pe(dev_b, xs, pe)
    want, aux = go.ucb_pe_score(pred_a, pred_b, xs, mode=mode, threshold=0.1, tr_rows=n + 2, trust_radius_value=0.3)
    np.testing.assert_allclose(out['score'].cpu().numpy(), want, atol=1e-9)
    np.testing.assert_allclose(out['stddev_from_all'].cpu().numpy(), aux['stddev_from_all'], atol=1e-9)
  dev_a.close(); dev_b.close()

  p = vz.ProblemStatement()
  for i in range(3):
    p.search_space.root.add_float_param(f'x{i}', 0.0, 1.0)
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  trials = []
  for i in range(25):
    xv = rng.uniform(size=3)
    t = vz.Trial(parameters={f'x{j}': float(xv[j]) for j in range(3)}, id=i + 1)
    t.complete(vz.Measurement({'obj': float(3.0 * xv[0] + xv[1] - 0.5 * xv[2])}))
    trials.append(t)
  fac = vb.VectorizedOptimizerFactory(strategy_factory=vb.VectorizedEagleStrategyFactory(eagle_config=gp_ucb_pe.default_eagle_config),
                                      max_evaluations=2500, suggestion_batch_size=25)
  des = gp_ucb_pe.VizierGPUCBPEBandit(p, acquisition_optimizer_factory=fac, mixes_linear_kernel=True, rng=3)
  des.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  sugg = des.suggest(2)
  assert len(sugg) == 2
  params_txt = sugg[0].metadata.ns('google_gp_ucb_pe_bandit').ns('prediction_in_warped_y_space')['params']
  assert 'linear_coef=1.0' in params_txt
  pred = des.predict(trials[:5])
  assert np.all(np.isfinite(pred.mean)) and np.all(pred.stddev > 0)
