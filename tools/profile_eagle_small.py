"""One short persistent-Eagle run (N=50, D=4, pool 25, batch 25) for `ncu --set full -k regex:k_eagle_persistent64`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vizier_b200 import gp, _lib
from vizier_b200.multi_gpu import trust_radius
rng = np.random.default_rng(0)
n, d = 50, 4
x = rng.uniform(size=(n, d)); y = rng.normal(size=n)
dev = gp.DeviceGP(0)
dev.fit(x, y, gp.GPHyperParams(1.0, np.full(d, 0.5), 1e-3))
acq = gp.Acquisition(1.8, True, trust_radius(n, d, 0))
cfg = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, 25, 25, 25 * 300)
dev.eagle_run(cfg, acq, 1, 7, prior=x)
torch.cuda.synchronize()
