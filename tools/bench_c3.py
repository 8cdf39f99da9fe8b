"""BASELINE config C3: Eagle with pool = batch = 1000, 200 iterations, against the C2 posterior."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vizier_b200 import gp, _lib
from vizier_b200.multi_gpu import trust_radius
x, y, th = bench.make_problem()
dev = gp.DeviceGP(0)
dev.fit(x, y, gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2']))
acq = gp.Acquisition(1.8, True, trust_radius(1000, 20, 0))
cfg = _lib.EagleConfig(0.45, 1.5, 0.008, 0.16, 7e-5, 0.7, 0.5, 0.96, 1000, 1000, 200_000)
dev.eagle_run(cfg, acq, 1, 7, prior=x)
torch.cuda.synchronize(); t0 = time.perf_counter()
dev.eagle_run(cfg, acq, 1, 7, prior=x)
torch.cuda.synchronize(); t = time.perf_counter() - t0
print(json.dumps({'C3_eagle_P1000_B1000_200it_s': t, 'us_per_iteration': 1e6 * t / 200, 'small_tiles': os.environ.get('VZGP_SMALL_TILES', 'default')}))
