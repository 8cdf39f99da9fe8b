"""Latency of the fit path: vzgp_fit and one NLL + gradient evaluation at N=1000/D=20 and C4 (N=2000/D=50),
alone and with 4 concurrent restarts (the ARD configuration).  Prints one JSON object; run on the GPU box."""
import json, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vizier_b200 import gp, ard

out = {}
rng = np.random.default_rng(0)

def best_of(fn, reps=20, warm=3):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(reps):
    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
  return 1e3 * float(np.median(ts)), 1e3 * float(np.min(ts))

for n, d in ((1000, 20), (2000, 50), (500, 10), (200, 5)):
  x = rng.uniform(size=(n, d)); y = -np.sum((x - 0.3) ** 2, axis=1) + 0.05 * rng.normal(size=n)
  y = (y - y.mean()) / y.std()
  dev = gp.DeviceGP(0)
  p = gp.GPHyperParams(1.0, np.full(d, 2.0), 1e-2)
  xt = torch.from_numpy(x).cuda(); yt = torch.from_numpy(y).cuda()
  f = dev.make_loss_fn(xt, yt)
  th = p.to_vector()
  med, mn = best_of(lambda: f(th))
  out[f'nll_grad_N{n}_D{d}_ms'] = {'median': med, 'min': mn}
  med, mn = best_of(lambda: dev.fit(xt, yt, p), reps=10)
  out[f'fit_N{n}_D{d}_ms'] = {'median': med, 'min': mn}
  # 4 concurrent evaluation chains (one handle / stream / host thread each), 30 evaluations per chain
  fns = ard.loss_functions(dev, xt, yt, None, d, 0, workers=4)
  for g in fns:
    g(th)
  devs = [dev] + list(getattr(dev, '_ard_workers', []))[:3]
  gpu_ms = [[] for _ in fns]
  def chain(i, g):
    st = devs[i].stream
    for _ in range(30):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(st); g(th); e1.record(st); e1.synchronize()
      gpu_ms[i].append(e0.elapsed_time(e1))
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  ths = [threading.Thread(target=chain, args=(i, g)) for i, g in enumerate(fns)]
  [t.start() for t in ths]; [t.join() for t in ths]
  dt = time.perf_counter() - t0
  out[f'nll_grad_N{n}_D{d}_4_concurrent_ms_per_eval'] = 1e3 * dt / 120
  out[f'nll_grad_N{n}_D{d}_4_concurrent_gpu_ms_of_one_eval'] = float(np.median(np.concatenate(gpu_ms)))
  gpu1 = []
  for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(dev.stream); f(th); e1.record(dev.stream); e1.synchronize(); gpu1.append(e0.elapsed_time(e1))
  out[f'nll_grad_N{n}_D{d}_alone_gpu_ms_of_one_eval'] = float(np.median(gpu1))
  for mode in ('batched', 'threaded'):
    ard.BATCHED_ARD = mode == 'batched'
    ard.train_gp(dev, xt, yt, rng=np.random.default_rng(1))      # buffers, graphs
    t0 = time.perf_counter()
    ard.train_gp(dev, xt, yt, rng=np.random.default_rng(1))
    out[f'ard_4x50_N{n}_D{d}_{mode}_s'] = time.perf_counter() - t0
  ard.BATCHED_ARD = True
  dev.close()
print(json.dumps(out, indent=1))
