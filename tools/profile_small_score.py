"""25-candidate score at N=1000, D=20 (small-pool kernels) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vizier_b200 import gp
from vizier_b200.multi_gpu import trust_radius
x, y, th = bench.make_problem()
dev = gp.DeviceGP(0)
dev.fit(x, y, gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2']))
acq = gp.Acquisition(1.8, True, trust_radius(1000, 20, 0))
xs = torch.from_numpy(np.random.default_rng(0).uniform(size=(25, 20))).cuda()
for _ in range(5):
  dev.score(xs, acq)
dev.synchronize()
