"""cProfile of one VizierGPBandit.suggest() at N=1000, D=20 (random 100k pool), after warm-up."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vizier_b200 import optimizers as vb, vz, ard
from vizier_b200.designers import gp_bandit
x, y, _ = bench.make_problem()
p = vz.ProblemStatement()
for i in range(bench.DIM):
  p.search_space.root.add_float_param(f'x{i}', 0.0, 1.0)
p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
def trials(xs, ys, first):
  out = []
  for i, (xv, yv) in enumerate(zip(xs, ys)):
    t = vz.Trial(parameters={f'x{j}': float(xv[j]) for j in range(bench.DIM)}, id=first + i)
    t.complete(vz.Measurement({'obj': float(yv)})); out.append(t)
  return out
fac = vb.VectorizedOptimizerFactory(strategy_factory=vb.random_strategy_factory, max_evaluations=100000, suggestion_batch_size=100000)
d = gp_bandit.VizierGPBandit(p, rng=1, acquisition_optimizer_factory=fac)
d.update(vz.CompletedTrials(trials(x, y, 1)), vz.ActiveTrials())
d.suggest(1)
rng = np.random.default_rng(5)
evals = {'n': 0}
orig = ard._lockstep
def counting(batch_fn, inits, bounds, **kw):
  def wrapped(idx, pts):
    evals['n'] += 1
    return batch_fn(idx, pts)
  wrapped.n_restarts = batch_fn.n_restarts
  return orig(wrapped, inits, bounds, **kw)
ard._lockstep = counting
for rep in range(3):
  xn = rng.uniform(size=(1, bench.DIM)); yn = -np.sum((xn - 0.3) ** 2, axis=1)
  d.update(vz.CompletedTrials(trials(xn, yn, 1001 + rep)), vz.ActiveTrials())
  evals['n'] = 0
  if rep == 2:
    pr = cProfile.Profile(); pr.enable()
  t0 = time.perf_counter(); d.suggest(1); dt = time.perf_counter() - t0
  if rep == 2:
    pr.disable()
  print('suggest', rep, round(dt, 4), 's; lock-step rounds', evals['n'])
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
