"""cProfile of one GP-UCB-PE suggest() at 1000 trials (host-side view: where the wall time is spent)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vizier_b200 import vz
from vizier_b200.designers import gp_ucb_pe
p = vz.ProblemStatement()
for i in range(20):
  p.search_space.root.add_float_param(f'x{i}', 0.0, 1.0)
p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
rng = np.random.default_rng(0)
ts = []
for i in range(1000):
  x = rng.uniform(size=20)
  t = vz.Trial(parameters={f'x{j}': float(x[j]) for j in range(20)}, id=i + 1)
  t.complete(vz.Measurement({'obj': float(-np.sum((x - 0.3) ** 2) + 0.05 * rng.normal())}))
  ts.append(t)
d = gp_ucb_pe.VizierGPUCBPEBandit(p, rng=1)
d.update(vz.CompletedTrials(ts), vz.ActiveTrials())
d.suggest(1)
pr = cProfile.Profile()
pr.enable(); d._last_params_key = None if hasattr(d, '_last_params_key') else None; out = d.suggest(1); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
