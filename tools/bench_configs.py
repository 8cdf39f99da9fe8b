"""Timings for the BASELINE.json configs other than the headline (C1, C3, C4) + fit latency.
Prints one JSON object; run on the GPU box."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vizier_b200 import gp, ard, _lib, vz
from vizier_b200 import optimizers as vb
from vizier_b200.acquisitions import trust_radius
from vizier_b200.designers import gp_bandit

out = {}
rng = np.random.default_rng(0)

def sync_time(fn, reps=3):
  fn(); torch.cuda.synchronize()
  ts = []
  for _ in range(reps):
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
  return min(ts)

# ---- C1: designer suggest() on 4-D, 50 trials, default settings (75k evals, batch 25, 4x50 ARD)
p = vz.ProblemStatement()
for i in range(4):
  p.search_space.root.add_float_param(f'x{i}', -5.0, 5.0)
p.metric_information.append(vz.MetricInformation(name='obj', goal=vz.ObjectiveMetricGoal.MAXIMIZE))
xs = rng.uniform(-5, 5, size=(50, 4))
trials = [vz.Trial(parameters={f'x{j}': float(v) for j, v in enumerate(x)}, id=i + 1).complete(
    vz.Measurement({'obj': float(-np.sum((x / 10 + 0.5 - 0.3) ** 2) + 0.05 * rng.normal())})) for i, x in enumerate(xs)]
d = gp_bandit.VizierGPBandit.from_problem(p, seed=0)
d.update(vz.CompletedTrials(trials), vz.ActiveTrials())
t0 = time.perf_counter(); s = d.suggest(1); t1 = time.perf_counter()
from vizier_b200 import profiler
d2 = gp_bandit.VizierGPBandit.from_problem(p, seed=1)
d2.update(vz.CompletedTrials(trials), vz.ActiveTrials())
with profiler.collect_events() as ev:
  t2 = time.perf_counter(); d2.suggest(1); t3 = time.perf_counter()
out['C1_suggest_s_first_call'] = t1 - t0
out['C1_suggest_s'] = t3 - t2
out['C1_breakdown_s'] = {k: sum(v) for k, v in ev.items()}

# ---- GP-UCB-PE (service DEFAULT): batch of 4 suggestions, same 4-D / 50-trial study, default settings
from vizier_b200.designers import gp_ucb_pe
d3 = gp_ucb_pe.VizierGPUCBPEBandit.from_problem(p, seed=2)
d3.update(vz.CompletedTrials(trials), vz.ActiveTrials())
d3.suggest(1)
d3.update(vz.CompletedTrials([]), vz.ActiveTrials())
t2 = time.perf_counter(); d3.suggest(4); t3 = time.perf_counter()
out['UCBPE_suggest4_N50_D4_s'] = t3 - t2

# ---- fit latency and C3: eagle 1000 fireflies x 200 iterations against the C2 posterior
import bench
x, y, th = bench.make_problem()
dev = gp.DeviceGP(0)
params = gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2'])
out['fit_N1000_D20_ms'] = 1e3 * sync_time(lambda: dev.fit(x, y, params))
acq = gp.Acquisition(1.8, True, trust_radius(1000, 20, 0))
cfg = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, 1000, 1000, 200_000)
t = sync_time(lambda: dev.eagle_run(cfg, acq, 1, 7, prior=x), reps=2)
out['C3_eagle_P1000_B1000_200it_s'] = t
out['C3_iterations_per_s'] = 200 / t
cfgd = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, 75, 25, 75_000)
t = sync_time(lambda: dev.eagle_run(cfgd, acq, 1, 7, prior=x), reps=1)
out['default_eagle_N1000_D20_3000it_s'] = t

# ---- C4: NLL+grad at N=2000, D=50, and a full 4 x 50 ARD fit
n, dd = 2000, 50
x4 = rng.uniform(size=(n, dd)); y4 = -np.sum((x4 - 0.3) ** 2, axis=1) + 0.05 * rng.normal(size=n)
y4 = (y4 - y4.mean()) / y4.std()
xt = torch.from_numpy(x4).cuda(); yt = torch.from_numpy(y4).cuda()
p4 = gp.GPHyperParams(1.0, np.full(dd, 2.0), 1e-2)
out['C4_nll_grad_N2000_D50_ms'] = 1e3 * sync_time(lambda: dev.loss_and_grad(xt, yt, p4))
t0 = time.perf_counter()
best, losses = ard.train_gp(dev, xt, yt, rng=np.random.default_rng(0), workers=1)
out['C4_ard_fit_4x50_sequential_s'] = time.perf_counter() - t0
ard.train_gp(dev, xt, yt, rng=np.random.default_rng(0))   # creates the worker handles
t0 = time.perf_counter()
best4, losses4 = ard.train_gp(dev, xt, yt, rng=np.random.default_rng(0))
out['C4_ard_fit_4x50_s'] = time.perf_counter() - t0
out['C4_final_losses'] = [float(v) for v in losses4]
out['C4_concurrent_equals_sequential'] = bool(np.array_equal(losses, losses4))
n, dd = 1000, 20
xt = torch.from_numpy(x).cuda(); yt = torch.from_numpy(y).cuda()
out['nll_grad_N1000_D20_ms'] = 1e3 * sync_time(lambda: dev.loss_and_grad(xt, yt, params))
ard.train_gp(dev, xt, yt, rng=np.random.default_rng(1))
t0 = time.perf_counter()
ard.train_gp(dev, xt, yt, rng=np.random.default_rng(1), workers=1)
out['ard_N1000_D20_4x50_sequential_s'] = time.perf_counter() - t0
t0 = time.perf_counter()
ard.train_gp(dev, xt, yt, rng=np.random.default_rng(1))
out['ard_N1000_D20_4x50_s'] = time.perf_counter() - t0
print(json.dumps(out, indent=1))
