"""One short default Eagle run at N=1000, D=20 (pool 75, batch 25: the cooperative grid kernel) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vizier_b200 import gp, _lib
from vizier_b200.multi_gpu import trust_radius
x, y, th = bench.make_problem()
dev = gp.DeviceGP(0)
dev.fit(x, y, gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2']))
acq = gp.Acquisition(1.8, True, trust_radius(1000, 20, 0))
cfg = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, 75, 25, 25 * 200)
for _ in range(2):
  dev.eagle_run(cfg, acq, 1, 7, prior=x)
dev.synchronize()
print('done')
