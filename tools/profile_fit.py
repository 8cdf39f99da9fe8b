"""Runs a few fits / nll_grad calls at N=1000 (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from vizier_b200 import gp
dev = gp.DeviceGP(0)
x, y, th = bench.make_problem()
p = gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2'])
xt = torch.from_numpy(x).cuda(); yt = torch.from_numpy(y).cuda()
for _ in range(3):
  dev.fit(xt, yt, p)
for _ in range(3):
  dev.loss_and_grad(xt, yt, p)
print('done')
