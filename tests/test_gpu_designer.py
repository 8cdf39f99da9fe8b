"""GPU tests of the drop-in designer, modelled on the reference's gp_bandit_test.py:
runs (`:130-235`), prediction accuracy (`:373-379`), convergence (`:513-544`)."""
import json

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from vizier_b200 import optimizers as vb  # noqa: E402
from vizier_b200 import profiler, vz  # noqa: E402
from vizier_b200.designers import gp_bandit  # noqa: E402


@pytest.fixture(autouse=True)
def _need_cuda():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')


def _problem(d=4, lo=-5.0, hi=5.0, goal=vz.ObjectiveMetricGoal.MAXIMIZE):
  p = vz.ProblemStatement()
  for i in range(d):
    p.search_space.root.add_float_param(f'x{i}', lo, hi)
  p.metric_information.append(vz.MetricInformation(name='obj', goal=goal))
  return p


def _complete(suggestions, f, start_id):
  out = []
  for i, s in enumerate(suggestions):
    t = s.to_trial(start_id + i)
    x = np.array([t.parameters[k].value for k in sorted(t.parameters)])
    t.complete(vz.Measurement({'obj': float(f(x))}))
    out.append(t)
  return out


small_opt = vb.VectorizedOptimizerFactory(strategy_factory=vb.VectorizedEagleStrategyFactory(),
                                          max_evaluations=2000, suggestion_batch_size=25)


def test_suggest_update_loop_and_metadata():
  p = _problem(4)
  d = gp_bandit.VizierGPBandit.from_problem(p, seed=0, acquisition_optimizer_factory=small_opt)
  f = lambda x: -np.sum((x - 1.0) ** 2)
  tid = 1
  with profiler.collect_events() as ev:
    for _ in range(6):
      sugg = d.suggest(2)
      assert len(sugg) == 2
      for s in sugg:
        assert p.search_space.contains(s.parameters)
      trials = _complete(sugg, f, tid); tid += len(trials)
      d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  assert 'VizierGPBandit.suggest' in ev and 'VizierGPBandit._update_gp' in ev
  md = sugg[0].metadata
  info = json.loads(md.ns('devinfo')['acquisition_optimization'])
  for k in ('acquisition', 'mean', 'stddev', 'raw_acquisition', 'linf_distance', 'radius'):
    assert k in info and np.isfinite(info[k])
  assert 'time_spent' in md.ns('oss_gp_bandit').ns('devinfo')


def test_prediction_accuracy_on_quadratic():
  # gp_bandit_test.py:373-379: 100 quasi-random points of f(x)=x^2 on [-1,1]; |mu(0) - f(0)| < 2e-2
  p = _problem(1, -1.0, 1.0)
  d = gp_bandit.VizierGPBandit.from_problem(p, seed=1, acquisition_optimizer_factory=small_opt)
  xs = (np.arange(100) + 0.5) / 100 * 2 - 1
  trials = [vz.Trial(parameters={'x0': float(x)}, id=i + 1).complete(vz.Measurement({'obj': float(x * x)})) for i, x in enumerate(xs)]
  d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  pred = d.predict([vz.TrialSuggestion({'x0': 0.0}), vz.TrialSuggestion({'x0': 0.5})], num_samples=2000)
  assert pred.mean.shape == (2,) and pred.stddev.shape == (2,)
  assert abs(pred.mean[0] - 0.0) < 2e-2
  assert abs(pred.mean[1] - 0.25) < 3e-2
  s = d.sample([vz.TrialSuggestion({'x0': 0.0})], num_samples=7)
  assert s.shape == (7, 1) and np.all(np.isfinite(s))


def test_minimisation_converges_faster_than_random():
  p = _problem(3, 0.0, 1.0, goal=vz.ObjectiveMetricGoal.MINIMIZE)
  f = lambda x: np.sum((x - 0.7) ** 2)
  d = gp_bandit.VizierGPBandit.from_problem(p, seed=3, acquisition_optimizer_factory=small_opt)
  best = np.inf
  tid = 1
  for _ in range(15):
    trials = _complete(d.suggest(1), f, tid); tid += 1
    best = min(best, trials[0].final_measurement.metrics['obj'].value)
    d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  rng = np.random.default_rng(0)
  rand_best = min(f(rng.uniform(size=3)) for _ in range(15))
  assert best < 0.02 and best < rand_best


def test_ensemble_designer_suggest_sample_predict():
  """gp_bandit_test.py:128-235 shape: ensemble_size 2 / 3, a padding schedule (no numerical effect
  here), batches, then sample() / predict() on fresh points."""
  p = _problem(3)
  f = lambda x: -np.sum((x - 0.5) ** 2)
  for ens, batch in ((2, 2), (3, 1)):
    d = gp_bandit.VizierGPBandit.from_problem(p, seed=1, acquisition_optimizer_factory=small_opt, ensemble_size=ens,
                                              padding_schedule='POWERS_OF_2', use_trust_region=False)
    tid = 1
    for _ in range(4):
      sugg = d.suggest(batch)
      assert len(sugg) == batch
      trials = _complete(sugg, f, tid); tid += len(trials)
      d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
    assert len(d._last_params) == ens
    pts = [vz.Trial(parameters={f'x{i}': v for i in range(3)}) for v in (-2.0, 0.5, 3.0)]
    s = d.sample(pts, num_samples=5)
    assert s.shape == (5, 3) and np.isfinite(s).all()
    assert d.sample([], num_samples=5).shape == (5, 0)
    pr = d.predict(pts)
    assert len(pr.mean) == 3 and np.isfinite(pr.mean).all() and np.isfinite(pr.stddev).all()
  with pytest.raises(ValueError):
    gp_bandit.VizierGPBandit.from_problem(p, ensemble_size=9)


def test_gp_ucb_pe_only_active_trials():
  """Parallel workers at study start: suggestions are requested while the seed trials are still ACTIVE and
  nothing has completed.  The reference runs its normal code on an empty data set (gp_ucb_pe.py:1356-1445,
  :1006-1155); here that is pure exploration on the GP conditioned on the pending points."""
  import datetime
  from vizier_b200.designers import gp_ucb_pe
  p = _problem(2, lo=0.0, hi=1.0)
  d = gp_ucb_pe.VizierGPUCBPEBandit(p, rng=5, acquisition_optimizer_factory=vb.VectorizedOptimizerFactory(
      strategy_factory=vb.VectorizedEagleStrategyFactory(eagle_config=gp_ucb_pe.default_eagle_config),
      max_evaluations=3000, suggestion_batch_size=25))
  seeds = d.suggest(1)
  active = []
  for i, sgg in enumerate(seeds):
    t = sgg.to_trial(i + 1)
    t.creation_time = datetime.datetime.now()
    active.append(t)
  d.update(vz.CompletedTrials([]), vz.ActiveTrials(active))
  out = d.suggest(3)
  assert len(out) == 3
  pts = np.array([[s.parameters['x0'].value, s.parameters['x1'].value] for s in out])
  pend = np.array([[t.parameters['x0'].value, t.parameters['x1'].value] for t in active])
  assert ((pts >= 0) & (pts <= 1)).all()
  every = np.vstack([pend, pts])
  dist = np.linalg.norm(every[:, None, :] - every[None, :, :], axis=-1) + np.eye(len(every))
  assert dist.min() > 0.05                                       # exploration: nothing piles up on a pending point
  info = json.loads(out[0].metadata.ns('devinfo')['acquisition_optimization'])
  assert info['mean'] == 0.0 and info['stddev_from_all'] <= info['stddev'] + 1e-12


def test_gp_ucb_pe_sample_and_predict():
  """gp_ucb_pe.py:1262-1354: sample() / predict() shapes, finiteness, and a sane posterior mean."""
  from vizier_b200.designers import gp_ucb_pe
  p = _problem(2)
  f = lambda x: -np.sum((x - 1.0) ** 2)
  rng = np.random.default_rng(5)
  trials = []
  for i in range(25):
    t = vz.Trial(parameters={f'x{j}': float(v) for j, v in enumerate(rng.uniform(-5, 5, 2))}, id=i + 1)
    t.complete(vz.Measurement({'obj': float(f(np.array([t.parameters['x0'].value, t.parameters['x1'].value])))}))
    trials.append(t)
  d = gp_ucb_pe.VizierGPUCBPEBandit(p, rng=3)
  d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  pts = [vz.Trial(parameters={'x0': a, 'x1': b}) for a, b in ((1.0, 1.0), (-4.0, 4.0), (0.0, 2.0))]
  s = d.sample(pts, num_samples=7)
  assert s.shape == (7, 3) and np.isfinite(s).all()
  assert d.sample([], num_samples=4).shape == (4, 0)
  pr = d.predict(pts, num_samples=400)
  assert len(pr.mean) == 3 and np.isfinite(pr.mean).all() and (pr.stddev >= 0).all()
  assert pr.mean[0] > pr.mean[1]          # the optimum region predicts higher than a far corner
  assert abs(pr.mean[0] - f(np.array([1.0, 1.0]))) < 3.0


def test_two_studies_concurrently_match_sequential():
  """Different studies may call suggest() concurrently (vizier_service.py:297 only serialises per study):
  two designers driven from two threads on one GPU give exactly what they give one after the other."""
  import concurrent.futures as cf
  from vizier_b200.designers import gp_ucb_pe

  def run(kind_seed):
    kind, seed = kind_seed
    p = _problem(3)
    f = lambda x: -np.sum((x - 0.7) ** 2)
    rng = np.random.default_rng(seed)
    trials = []
    for i in range(70 if kind == 'bandit' else 30):     # 70 trials: grid-persistent loop, 30: single-CTA loop
      t = vz.Trial(parameters={f'x{j}': float(v) for j, v in enumerate(rng.uniform(-5, 5, 3))}, id=i + 1)
      t.complete(vz.Measurement({'obj': float(f(np.array([t.parameters[k].value for k in sorted(t.parameters)])))}))
      trials.append(t)
    if kind == 'bandit':
      d = gp_bandit.VizierGPBandit.from_problem(p, seed=seed, acquisition_optimizer_factory=small_opt)
    else:
      d = gp_ucb_pe.VizierGPUCBPEBandit(p, rng=seed, acquisition_optimizer_factory=vb.VectorizedOptimizerFactory(
          strategy_factory=vb.VectorizedEagleStrategyFactory(eagle_config=gp_ucb_pe.default_eagle_config),
          max_evaluations=2000, suggestion_batch_size=25))
    d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
    return [[s.parameters[f'x{j}'].value for j in range(3)] for s in d.suggest(2)]

  jobs = [('bandit', 11), ('ucbpe', 12), ('bandit', 13), ('ucbpe', 14)]
  sequential = [run(j) for j in jobs]
  with cf.ThreadPoolExecutor(max_workers=4) as pool:
    concurrent = list(pool.map(run, jobs))
  assert concurrent == sequential


def test_random_pool_optimizer_factory():
  # the C2 shape through the designer API: score one M-candidate uniform pool, take the top `count`
  p = _problem(5, 0.0, 1.0)
  opt = vb.VectorizedOptimizerFactory(strategy_factory=vb.random_strategy_factory, max_evaluations=20_000, suggestion_batch_size=20_000)
  d = gp_bandit.VizierGPBandit.from_problem(p, seed=5, acquisition_optimizer_factory=opt)
  rng = np.random.default_rng(2)
  trials = [vz.Trial(parameters={f'x{j}': float(v) for j, v in enumerate(x)}, id=i + 1).complete(
      vz.Measurement({'obj': float(-np.sum((x - 0.4) ** 2))})) for i, x in enumerate(rng.uniform(size=(30, 5)))]
  d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  sugg = d.suggest(3)
  assert len(sugg) == 3 and len({tuple(sorted(s.parameters.as_dict().items())) for s in sugg}) == 3


def test_all_parameter_types_search_space():
  """gp_bandit_test.py:130-235 runs the designer on `flat_space_with_all_types`; same idea here."""
  p = vz.ProblemStatement()
  r = p.search_space.root
  r.add_float_param('lr', 1e-4, 1e-1, scale_type=vz.ScaleType.LOG)
  r.add_float_param('x', -1.0, 1.0)
  r.add_int_param('layers', 1, 6)
  r.add_discrete_param('bs', [16, 32, 64, 128])
  r.add_categorical_param('opt', ['adam', 'sgd', 'lion'])
  r.add_categorical_param('norm', ['bn', 'ln'])
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))

  def f(t):
    v = t.parameters
    return (-(np.log10(v['lr'].value) + 2.5) ** 2 - v['x'].value ** 2 - 0.1 * (v['layers'].value - 4) ** 2
            + (0.5 if v['opt'].value == 'lion' else 0.0) + (0.2 if v['norm'].value == 'ln' else 0.0))

  d = gp_bandit.VizierGPBandit.from_problem(p, seed=7, acquisition_optimizer_factory=small_opt)
  tid, best = 1, -np.inf
  for _ in range(12):
    sugg = d.suggest(2)
    trials = []
    for s in sugg:
      assert p.search_space.contains(s.parameters), s.parameters
      t = s.to_trial(tid); tid += 1
      t.complete(vz.Measurement({'obj': float(f(t))}))
      best = max(best, t.final_measurement.metrics['obj'].value)
      trials.append(t)
    d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  assert best > -0.6


def test_gp_ucb_pe_designer_batches_and_pending():
  """Service DEFAULT algorithm (policy_factory.py:40-47): batch suggestions are spread out by the
  pure-exploration acquisition, pending trials are respected, metadata carries the decision."""
  from vizier_b200.designers import gp_ucb_pe
  p = _problem(3, 0.0, 1.0)
  f = lambda x: -np.sum((x - 0.6) ** 2)
  opt = vb.VectorizedOptimizerFactory(strategy_factory=vb.VectorizedEagleStrategyFactory(eagle_config=gp_ucb_pe.default_eagle_config),
                                      max_evaluations=3000, suggestion_batch_size=25)
  d = gp_ucb_pe.VizierGPUCBPEBandit.from_problem(p, seed=4, acquisition_optimizer_factory=opt)
  first = d.suggest(1)
  assert first[0].metadata['seeded'] == 'center'
  rng = np.random.default_rng(0)
  tid = 1
  trials = []
  for x in rng.uniform(size=(8, 3)):
    t = vz.Trial(parameters={f'x{j}': float(v) for j, v in enumerate(x)}, id=tid); tid += 1
    trials.append(t.complete(vz.Measurement({'obj': float(f(x))})))
  d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
  best = max(t.final_measurement.metrics['obj'].value for t in trials)
  for it in range(5):
    sugg = d.suggest(3)
    assert len(sugg) == 3
    pts = np.array([[s.parameters[f'x{j}'].value for j in range(3)] for s in sugg])
    assert np.min(np.linalg.norm(pts[:, None] - pts[None], axis=-1) + 10 * np.eye(3)) > 1e-3   # distinct points
    flags = [s.metadata.ns('google_gp_ucb_pe_bandit').ns('prediction_in_warped_y_space')['use_ucb'] for s in sugg]
    assert flags[0] in ('True', 'False') and 'False' in flags[1:]       # later members of a batch explore
    new = []
    for s in sugg:
      assert p.search_space.contains(s.parameters)
      t = s.to_trial(tid); tid += 1
      x = np.array([t.parameters[f'x{j}'].value for j in range(3)])
      new.append(t.complete(vz.Measurement({'obj': float(f(x))})))
      best = max(best, new[-1].final_measurement.metrics['obj'].value)
    d.update(vz.CompletedTrials(new), vz.ActiveTrials())
  assert best > -0.05
  # an ACTIVE (pending) trial steers the next suggestion away from itself
  pending = vz.Trial(parameters={'x0': 0.6, 'x1': 0.6, 'x2': 0.6}, id=tid)
  d.update(vz.CompletedTrials([]), vz.ActiveTrials([pending]))
  s = d.suggest(1)[0]
  x = np.array([s.parameters[f'x{j}'].value for j in range(3)])
  assert np.linalg.norm(x - 0.6) > 1e-3


def test_designer_policy_drives_the_gpu_designer():
  """The service's path (policy_factory.py:48-53 -> DesignerPolicy.suggest, designer_policy.py:77-112):
  a fresh VizierGPBandit per request, updated with all COMPLETED + ACTIVE trials, asked for suggestions."""
  from vizier_b200 import designer_policy as dp
  p = _problem(3)
  sup = dp.InRamPolicySupporter(p)
  policy = dp.DesignerPolicy(sup, lambda problem: gp_bandit.VizierGPBandit.from_problem(problem, seed=3))
  rng = np.random.default_rng(0)
  for step in range(4):
    decision = policy.suggest(dp.SuggestRequest(study_config=p, count=2))
    assert len(decision.suggestions) == 2
    for s in decision.suggestions:
      assert p.search_space.contains(s.parameters)
      x = np.array([s.parameters[f'x{i}'].value for i in range(3)])
      sup.AddTrials([vz.Trial(parameters=s.parameters).complete(vz.Measurement({'obj': float(-np.sum((x - 0.3) ** 2))}))])
  assert len(sup.GetTrials(status_matches=vz.TrialStatus.COMPLETED)) == 8
  assert 'oss_gp_bandit' in [ns[0] for ns in decision.suggestions[0].metadata.namespaces() if ns]
