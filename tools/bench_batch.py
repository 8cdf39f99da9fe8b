"""One lock-step ARD round (vzgp_nll_grad_batch) for R = 1..5 concurrent evaluations: ms per round."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vizier_b200 import gp, ard
out = {}
rng = np.random.default_rng(0)
for n, d in ((1000, 20), (2000, 50)):
  x = rng.uniform(size=(n, d)); y = rng.normal(size=n)
  xt = torch.from_numpy(x).cuda(); yt = torch.from_numpy(y).cuda()
  th = gp.GPHyperParams(1.0, np.full(d, 2.0), 1e-2).to_vector()
  for R in (1, 2, 3, 4, 5):
    dev = gp.DeviceGP(0)
    shares = [os.environ.get('VZGP_ARD_SHARE')]
    f = ard.batch_loss_function(dev, xt, yt, None, R) if R > 1 else None
    if R == 1:
      g = dev.make_loss_fn(xt, yt)
      call = lambda: g(th)
    else:
      call = lambda: f(list(range(R)), [th] * R)
    for _ in range(3):
      call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
      t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    out[f'N{n}_R{R}_ms_per_round'] = 1e3 * float(np.median(ts))
    for w in getattr(dev, '_ard_workers', []):
      w.close()
    dev.close()
print(json.dumps(out, indent=1))
