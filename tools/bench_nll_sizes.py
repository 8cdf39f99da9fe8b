"""nll_grad latency by N (D=20), and evaluation counts of the GP-UCB-PE ARD (maxiter 500)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vizier_b200 import gp, ard
out = {}
rng = np.random.default_rng(0)
dev = gp.DeviceGP(0)
for n in (64, 128, 200, 256, 512, 1000):
  x = rng.uniform(size=(n, 20)); y = -np.sum((x - 0.3) ** 2, axis=1) + 0.05 * rng.normal(size=n); y = (y - y.mean()) / y.std()
  xt = torch.from_numpy(x).cuda(); yt = torch.from_numpy(y).cuda()
  f = dev.make_loss_fn(xt, yt)
  th = gp.GPHyperParams(1.0, np.full(20, 0.5), 1e-2).to_vector()
  for _ in range(5): f(th)
  t0 = time.perf_counter()
  for _ in range(50): f(th)
  out[f'nll_grad_N{n}_us'] = 1e6 * (time.perf_counter() - t0) / 50
  if n in (200, 1000):
    cnt = [0]
    def g(t, f=f, cnt=cnt):
      cnt[0] += 1
      return f(t)
    lo, hi = gp.param_bounds(20, 0)
    inits = ard.log_uniform_init(np.random.default_rng(1), 20, 0, 4)
    opt = ard.ScipyLbfgsB(ard.LbfgsBOptions(num_line_search_steps=20, tol=1e-5, maxiter=500))
    per = []
    for t0_ in inits:
      cnt[0] = 0; t1 = time.perf_counter(); opt._one(g, t0_, list(zip(lo, hi))); per.append((cnt[0], round(time.perf_counter() - t1, 4)))
    out[f'ucbpe_ard_N{n}_evals_and_seconds_per_restart'] = per
print(json.dumps(out, indent=1))
