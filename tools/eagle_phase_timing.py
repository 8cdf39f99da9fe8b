"""Phase split of the persistent small-study Eagle kernel (build libvzgp with `make EXTRA=-DVZ_EAGLE_TIMING`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vizier_b200 import gp, _lib
from vizier_b200.multi_gpu import trust_radius
rng = np.random.default_rng(0)
for n, d in ((50, 4), (60, 20), (1000, 20), (200, 20)):
  x = rng.uniform(size=(n, d)); y = rng.normal(size=n)
  dev = gp.DeviceGP(0)
  dev.fit(x, y, gp.GPHyperParams(1.0, np.full(d, 0.5), 1e-3))
  acq = gp.Acquisition(1.8, True, trust_radius(n, d, 0))
  pool = 25 if d == 4 else 75
  cfg = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, pool, 25, 75_000)
  dev.eagle_run(cfg, acq, 1, 7, prior=x)
  torch.cuda.synchronize()
