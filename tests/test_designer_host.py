"""CPU tests of the designer's host logic (no GPU): converters, seeding, validation."""
import numpy as np
import pytest

from vizier_b200 import converters, vz
from vizier_b200 import acquisitions as acq_lib


def _problem():
  p = vz.ProblemStatement()
  r = p.search_space.root
  r.add_float_param('lr', 1e-4, 1e-1, scale_type=vz.ScaleType.LOG)
  r.add_float_param('x', -5.0, 5.0)
  r.add_float_param('rl', 1.0, 10.0, scale_type=vz.ScaleType.REVERSE_LOG)
  r.add_int_param('layers', 1, 5)
  r.add_discrete_param('bs', [16, 32, 128])
  r.add_categorical_param('opt', ['adam', 'sgd'])
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MINIMIZE))
  return p


def test_converter_scaling_and_round_trip():
  p = _problem()
  c = converters.TrialToModelInputConverter.from_problem(p)
  assert (c.n_continuous, c.n_categorical) == (5, 1)
  t = vz.Trial(parameters={'lr': 1e-3, 'x': 0.0, 'rl': 10.0, 'layers': 3, 'bs': 32, 'opt': 'sgd'})
  t.complete(vz.Measurement({'obj': 2.0}))
  bad = vz.Trial(parameters={'lr': 1e-1, 'x': 5.0, 'rl': 1.0, 'layers': 5, 'bs': 128, 'opt': 'unknown'})
  bad.complete(vz.Measurement(), infeasibility_reason='diverged')
  (cont, cat), y = c.to_xy([t, bad])
  # LOG: (log x - log lo)/(log hi - log lo);  LINEAR; REVERSE_LOG: 1 - (log(lo+hi-x) - log lo)/(...)
  np.testing.assert_allclose(cont[0], [1 / 3, 0.5, 1.0, 0.5, (32 - 16) / (128 - 16)], rtol=1e-12)
  np.testing.assert_allclose(cont[1], [1.0, 1.0, 0.0, 1.0, 1.0], atol=1e-12)
  assert cat.dtype == np.int32 and cat[:, 0].tolist() == [1, 2]  # 'sgd' -> 1, unknown -> len(feasible)
  np.testing.assert_array_equal(y[:, 0], [-2.0, np.nan])      # sign flipped for MINIMIZE; infeasible -> NaN
  back = c.to_parameters(cont, cat)
  assert back[0].as_dict() == pytest.approx({'lr': 1e-3, 'x': 0.0, 'rl': 10.0, 'layers': 3, 'bs': 32.0, 'opt': 'sgd'})
  assert 'opt' not in back[1].as_dict()  # out-of-vocabulary index maps to no value
  # continuified values round to the nearest feasible value and DOUBLEs are clipped
  q = c.to_parameters(np.array([[0.5, 1.3, 0.5, 0.6, 0.2]]), np.array([[0]], np.int32))[0].as_dict()
  assert q['layers'] == 3 and q['bs'] == 32.0 and q['x'] == 5.0 and q['opt'] == 'adam'


def test_continuous_feasible_values_and_trust_region_mask():
  p = _problem()
  c = converters.TrialToModelInputConverter.from_problem(p)
  fv = c.continuous_feasible_values(1000)
  assert [len(v) for v in fv] == [0, 0, 0, 5, 3]
  mask = acq_lib.trust_region_dim_mask(fv)
  # ints 1..5 scaled have gaps 0.25 > 0.2 -> excluded; bs has a gap of 0.857 -> excluded
  assert mask.tolist() == [True, True, True, False, False]
  a = acq_lib.make_acquisition(10, fv, 5, 1)
  assert a.trust_radius == pytest.approx(0.2 + 0.3 * 10 / (5 * (3 + 1 + 1)))


def test_designer_validation_errors_need_no_gpu():
  from vizier_b200.designers import gp_bandit
  with pytest.raises(ValueError):
    gp_bandit.VizierGPBandit(vz.ProblemStatement())
  p = _problem()
  p.metric_information.append(vz.MetricInformation(name='m2', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  d = gp_bandit.VizierGPBandit(p, num_scalarizations=16)     # multi-metric: scalarised UCB, no trust region
  assert d._scal_weights.shape == (16, 2) and not d._use_trust_region
  np.testing.assert_allclose(np.linalg.norm(d._scal_weights, axis=1), 1.0)
  with pytest.raises(NotImplementedError):
    gp_bandit.VizierGPBandit(p, ensemble_size=2)
  assert gp_bandit.VizierGPBandit(_problem(), linear_coef=0.1)._linear_coef == 0.1
  with pytest.raises(NotImplementedError):
    gp_bandit.VizierGPBandit(p, linear_coef=0.1)     # linear kernel + several metrics


def test_seed_trials_centre_then_quasi_random():
  from vizier_b200.designers import gp_bandit
  d = gp_bandit.VizierGPBandit.from_problem(_problem(), seed=1, num_seed_trials=3)
  s = d.suggest(3)
  assert len(s) == 3
  centre = s[0].parameters.as_dict()
  assert centre['x'] == 0.0 and centre['layers'] == 3 and s[0].metadata['seeded'] == 'center'
  space = _problem().search_space
  for sug in s:
    assert space.contains(sug.parameters)


def test_lean_lbfgsb_driver_is_scipy_minimize():
  """ard._lean_lbfgsb drives the same compiled routine as scipy.optimize.minimize(method='L-BFGS-B'):
  identical iterates, so identical final point and value (jaxopt_wrappers.py:139-152 uses the latter)."""
  import scipy.optimize as sopt
  from vizier_b200 import ard
  if ard._setulb is None:
    pytest.skip('scipy.optimize._lbfgsb not importable')
  rng = np.random.default_rng(3)

  def f(t):
    return float(np.sum((np.log(t) - 0.3) ** 2) + 1e-3 * np.sum(t ** 4)), 2 * (np.log(t) - 0.3) / t + 4e-3 * t ** 3

  for d, maxiter in ((6, 50), (22, 5), (52, 500)):
    bounds = [(1e-3, 10.0)] * d
    x0 = np.exp(rng.uniform(np.log(1e-3), np.log(10.0), d))
    want = sopt.minimize(f, x0, jac=True, method='L-BFGS-B', bounds=bounds,
                         options={'maxiter': maxiter, 'gtol': 1e-8, 'maxls': 20})
    x, fv = ard._lean_lbfgsb(f, x0, bounds, maxiter=maxiter, gtol=1e-8, maxls=20)
    np.testing.assert_array_equal(x, want.x)
    assert fv == want.fun


def test_sorted_prior_features_reuse_converted_arrays():
  """trials_to_sorted_features(trials, converter, features) only permutes already converted arrays
  (vectorized_base.py:655-665 order: creation time, then insertion)."""
  import datetime
  p = vz.ProblemStatement()
  p.search_space.root.add_float_param('a', 0, 1)
  p.search_space.root.add_categorical_param('c', ['x', 'y', 'z'])
  p.metric_information.append(vz.MetricInformation(name='m', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  rng = np.random.default_rng(0)
  t0 = datetime.datetime(2024, 1, 1)
  ts = [vz.Trial(parameters={'a': float(rng.uniform()), 'c': ['x', 'y', 'z'][i % 3]}, id=i + 1,
                 creation_time=t0 + datetime.timedelta(seconds=int(rng.integers(0, 10)))) for i in range(30)]
  c = converters.TrialToModelInputConverter.from_problem(p)
  a = converters.trials_to_sorted_features(ts, c)
  b = converters.trials_to_sorted_features(ts, c, c.to_features(ts))
  np.testing.assert_array_equal(a[0], b[0])
  np.testing.assert_array_equal(a[1], b[1])
  assert converters.trials_to_sorted_features([], c, None) is None


def test_to_xy_cached_equals_to_xy():
  p = _problem()
  c = converters.TrialToModelInputConverter.from_problem(p)
  rng = np.random.default_rng(1)
  names = [pc.name for pc in p.search_space.parameters]
  trials = []
  for i in range(40):
    params = {}
    for pc in p.search_space.parameters:
      tname = getattr(pc.type, 'name', str(pc.type))
      if tname == 'CATEGORICAL':
        params[pc.name] = pc.feasible_values[int(rng.integers(len(pc.feasible_values)))]
      elif tname == 'DOUBLE':
        params[pc.name] = float(rng.uniform(pc.bounds[0], pc.bounds[1]))
      else:
        params[pc.name] = pc.feasible_values[int(rng.integers(len(pc.feasible_values)))]
    t = vz.Trial(parameters=params, id=i + 1)
    metric = p.metric_information[0].name if hasattr(p.metric_information, '__getitem__') else list(p.metric_information)[0].name
    t.complete(vz.Measurement({metric: float(rng.normal())}))
    trials.append(t)
  for upto in (10, 25, 40, 40):        # growing study, then an unchanged one
    (c1, z1), y1 = c.to_xy(trials[:upto])
    (c2, z2), y2 = c.to_xy_cached(trials[:upto])
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_array_equal(z1, z2)
    np.testing.assert_array_equal(y1, y2)
  assert names


def test_designer_policy_call_sequence():
  """DesignerPolicy.suggest (designer_policy.py:77-112): factory(problem) -> update(Completed, Active) ->
  suggest(count), a fresh designer per request, trials split by status."""
  from vizier_b200 import designer_policy as dp
  from vizier_b200 import vz
  calls = []

  class Recorder(vz.Designer):
    def __init__(self, problem):
      calls.append(('factory', problem))
    def update(self, completed, all_active):
      assert isinstance(completed, vz.CompletedTrials) and isinstance(all_active, vz.ActiveTrials)
      calls.append(('update', [t.id for t in completed.trials], [t.id for t in all_active.trials]))
    def suggest(self, count=None):
      calls.append(('suggest', count))
      return [vz.TrialSuggestion({'x': 0.5}) for _ in range(count or 1)]

  p = vz.ProblemStatement()
  p.search_space.root.add_float_param('x', 0.0, 1.0)
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  sup = dp.InRamPolicySupporter(p)
  done = [vz.Trial(parameters={'x': 0.1 * i}).complete(vz.Measurement({'obj': float(i)})) for i in range(3)]
  sup.AddTrials(done + [vz.Trial(parameters={'x': 0.9})])
  policy = dp.DesignerPolicy(sup, Recorder)
  for count in (2, 1):
    decision = policy.suggest(dp.SuggestRequest(study_config=p, count=count))
    assert len(decision.suggestions) == count
  assert [c[0] for c in calls] == ['factory', 'update', 'suggest'] * 2
  assert calls[1] == ('update', [1, 2, 3], [4]) and calls[2] == ('suggest', 2)
  with pytest.raises(NotImplementedError):
    policy.early_stop(None)


def test_designers_implement_the_designer_and_predictor_interfaces():
  from vizier_b200 import vz
  from vizier_b200.designers import gp_bandit, gp_ucb_pe
  for cls in (gp_bandit.VizierGPBandit, gp_ucb_pe.VizierGPUCBPEBandit):
    assert issubclass(cls, vz.Designer) and issubclass(cls, vz.Predictor)
    assert not getattr(cls, '__abstractmethods__', None), cls.__abstractmethods__


def test_lockstep_driver_equals_one_restart_at_a_time():
  """ard._lockstep (all restarts advance together, one batched evaluation per round) gives every restart
  exactly the iterates it has alone - the batched ARD changes the schedule, not the optimisation."""
  from vizier_b200 import ard
  if ard._setulb is None:
    pytest.skip('SciPy without the C L-BFGS-B step routine')
  rng = np.random.default_rng(0)
  a = rng.normal(size=(6, 6)); q = a @ a.T + np.eye(6)
  c = rng.normal(size=6)

  def f(x):
    return float(0.5 * x @ q @ x + c @ x + 0.1 * np.sum(np.cos(3 * x))), q @ x + c - 0.3 * np.sin(3 * x)

  bounds = [(-2.0, 2.0)] * 6
  inits = rng.uniform(-2, 2, size=(5, 6))
  calls = []

  def batch(indices, points):
    calls.append(list(indices))
    out = [f(p) for p in points]
    return [o[0] for o in out], [o[1] for o in out]

  kw = dict(maxiter=50, gtol=1e-8, maxls=20)
  together = ard._lockstep(batch, inits, bounds, **kw)
  for t0, (x, fx) in zip(inits, together):
    xs, fs = ard._lean_lbfgsb(f, t0, bounds, **kw)
    np.testing.assert_array_equal(x, xs)
    assert fx == fs
  assert calls[0] == [0, 1, 2, 3, 4] and len(calls[-1]) >= 1 and len(calls) < sum(len(c) for c in calls)


def test_lean_lbfgsb_with_unbounded_variables():
  """The linear_coef model has unconstrained parameters (+-inf bounds): the lean driver must tell setulb so
  (nbd codes) and still reproduce scipy.optimize.minimize."""
  import scipy.optimize as sopt
  from vizier_b200 import ard
  if ard._setulb is None:
    pytest.skip('SciPy without the C L-BFGS-B step routine')

  def f(x):
    return float(np.sum((x - np.array([0.3, -2.0, 5.0])) ** 2) + 0.1 * x[0] * x[1]), 2 * (x - np.array([0.3, -2.0, 5.0])) + 0.1 * np.array([x[1], x[0], 0.0])

  bounds = [(0.0, 1.0), (-np.inf, np.inf), (-np.inf, 4.0)]
  x0 = np.array([0.9, 0.0, 0.0])
  x, fx = ard._lean_lbfgsb(f, x0, bounds, maxiter=50, gtol=1e-8, maxls=20)
  ref = sopt.minimize(f, x0, jac=True, method='L-BFGS-B', bounds=bounds, options={'maxiter': 50, 'gtol': 1e-8, 'maxls': 20})
  np.testing.assert_array_equal(x, ref.x)
  assert fx == ref.fun
