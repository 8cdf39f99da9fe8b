"""Default-designer eagle loop (pool 75, batch 25) against the C2 posterior: timing split and, under ncu,
the per-kernel launch list.  Usage: python tools/profile_eagle.py [n_iters]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vizier_b200 import gp, _lib
from vizier_b200.multi_gpu import trust_radius

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
x, y, th = bench.make_problem()
dev = gp.DeviceGP(0)
dev.fit(x, y, gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2']))
acq = gp.Acquisition(1.8, True, trust_radius(1000, 20, 0))

def run(evals, prior):
  cfg = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, 75, 25, evals)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  r = dev.eagle_run(cfg, acq, 1, 7, prior=prior)
  torch.cuda.synchronize(); return time.perf_counter() - t0, r

out = {}
if os.environ.get('VZ_PROFILE_ONLY'):
  run(25 * iters, x)
  sys.exit(0)
run(25 * 10, x)
out['seed_priors_plus_1it_s'] = run(25, x)[0]
out['no_prior_1it_s'] = run(25, None)[0]
t, _ = run(25 * iters, x)
out[f'{iters}it_s'] = t
out['us_per_iteration'] = 1e6 * (t - out['seed_priors_plus_1it_s']) / (iters - 1)
for n in (50, 200, 500):
  xs, ys = x[:n], y[:n]
  dev.fit(xs, ys, gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2']))
  a = gp.Acquisition(1.8, True, trust_radius(n, 20, 0))
  cfg = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, 75, 25, 25 * iters)
  dev.eagle_run(cfg, a, 1, 7, prior=xs)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  dev.eagle_run(cfg, a, 1, 7, prior=xs)
  torch.cuda.synchronize()
  out[f'N{n}_us_per_iteration'] = 1e6 * (time.perf_counter() - t0) / iters
print(json.dumps(out))
