"""Per-tile error map of the dataflow factorisation outputs (debugging aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.linalg as sla, torch
from oracle import gp_oracle as go
from vizier_b200 import gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(n)
x = rng.uniform(size=(n, 6))
a = go.kernel_matrix(go.GPParams(1.0, np.full(6, 0.6), 1e-2), x)
wl = np.linalg.cholesky(a); wi = sla.solve_triangular(wl, np.eye(n), lower=True); wk = wi.T @ wi
dev = gp.DeviceGP(0)
for rep in range(reps):
  outs, bad = dev.factor_inverse(a); dev.synchronize()
  for name, got, want in zip(('L', 'Linv', 'Kinv'), [o.cpu().numpy() for o in outs], (wl, wi, np.tril(wk))):
    got = np.tril(got)
    nb = (n + 63) // 64
    badt = []
    for i in range(nb):
      for j in range(i + 1):
        e = np.max(np.abs(got[i*64:(i+1)*64, j*64:(j+1)*64] - want[i*64:(i+1)*64, j*64:(j+1)*64]))
        if not e < 1e-8 * max(1.0, np.max(np.abs(want))):
          badt.append((i, j, float(e)))
    print(f'rep {rep} {name}: bad tiles {len(badt)}', badt[:12], flush=True)
