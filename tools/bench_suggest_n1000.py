"""End-to-end `suggest()` of the default designers on a 20-D study with 1000 completed trials."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vizier_b200 import vz, profiler
from vizier_b200.designers import gp_bandit, gp_ucb_pe

def problem(d=20):
  p = vz.ProblemStatement()
  for i in range(d):
    p.search_space.root.add_float_param(f'x{i}', 0.0, 1.0)
  p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
  return p

def trials(n, d, seed=0):
  rng = np.random.default_rng(seed)
  out = []
  for i in range(n):
    x = rng.uniform(size=d)
    t = vz.Trial(parameters={f'x{j}': float(x[j]) for j in range(d)}, id=i + 1)
    t.complete(vz.Measurement({'obj': float(-np.sum((x - 0.3) ** 2) + 0.05 * rng.normal())}))
    out.append(t)
  return out

out = {}
for n in (200, 1000):
  p = problem()
  ts = trials(n, 20)
  for name, mk in (('gp_bandit', lambda: gp_bandit.VizierGPBandit.from_problem(p, seed=1)),
                   ('gp_ucb_pe', lambda: gp_ucb_pe.VizierGPUCBPEBandit(p, rng=1))):
    d = mk()
    d.update(vz.CompletedTrials(ts), vz.ActiveTrials())
    d.suggest(1)                                   # first call: workspaces, worker handles
    d.update(vz.CompletedTrials(trials(1, 20, seed=n + 7)), vz.ActiveTrials())   # one new trial -> refit
    with profiler.collect_events() as ev:
      t0 = time.perf_counter(); d.suggest(1); dt = time.perf_counter() - t0
    out[f'{name}_N{n}_suggest_s'] = dt
    out[f'{name}_N{n}_breakdown'] = {k.split('.')[-1]: round(float(np.sum(v)), 4) for k, v in ev.items()}
print(json.dumps(out, indent=1))
