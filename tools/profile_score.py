"""Launches the C2 scoring kernel a few times (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vizier_b200 import gp
from vizier_b200.acquisitions import trust_radius

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = gp.DeviceGP(0)
x, y, th = bench.make_problem()
dev.fit(x, y, gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2']))
xs = dev.random_pool(bench.M_POOL, bench.DIM, seed=1)
acq = gp.Acquisition(1.8, True, trust_radius(bench.N_TRIALS, bench.DIM, 0))
for _ in range(n_launch):
  out = dev.score(xs, acq)
dev.synchronize()
print('done', float(out['score'][0]))
