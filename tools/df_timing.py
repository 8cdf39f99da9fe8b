"""Chain-CTA phase split of the dataflow factorisation (needs a library built with
`make -C vizier_b200/csrc EXTRA=-DVZ_DF_TIMING`): clock64 cycles per panel step for
wait(S) | load+T GEMM+store | syrk+assemble | potf2_inv_64 | store+release."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vizier_b200 import gp, _lib
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 20
rng = np.random.default_rng(0)
x = rng.uniform(size=(n, d)); y = rng.normal(size=n)
dev = gp.DeviceGP(0)
p = gp.GPHyperParams(1.0, np.full(d, 1.0), 1e-2)
for _ in range(3):
  dev.fit(x, y, p)
lib = _lib.load()
buf = (C.c_longlong * 512)()
lib.vzgp_debug_df_timing.restype = C.c_int
assert lib.vzgp_debug_df_timing(buf, 512) == 0
t = np.array(buf[:], dtype=np.int64).reshape(64, 8)
nb = (n + 63) // 64
rows = []
for j in range(nb):
  s = t[j]
  rows.append({'step': j, 'wait_S': int(s[1] - s[0]) if j else 0, 'T_gemm': int(s[2] - s[1]) if j else 0,
               'syrk': int(s[3] - s[2]) if j else int(s[3] - s[0]), 'potf2_inv': int(s[4] - s[3]), 'store': int(s[5] - s[4]),
               'total': int(s[5] - s[0])})
tot = {k: int(np.sum([r[k] for r in rows])) for k in rows[0] if k != 'step'}
print(json.dumps({'n': n, 'nb': nb, 'per_step_cycles': rows[:4] + rows[-2:], 'sum_cycles': tot,
                  'chain_cycles': int(t[nb - 1][5] - t[0][0])}, indent=1))

if hasattr(lib, 'vzgp_debug_la_timing'):
  b2 = (C.c_longlong * 64)()
  lib.vzgp_debug_la_timing.restype = C.c_int
  if lib.vzgp_debug_la_timing(b2) == 0:
    u = np.array(b2[:], dtype=np.int64)
    base = u[0]
    names = {0: 'D start', 1: 'D end', 2: 'UPD passed', 3: 'P end'}
    for mb in range(4):
      print('warp0 step', mb, {names[k]: int(u[8 * mb + k] - base) for k in range(4) if u[8 * mb + k]})
    print('end before final sync', int(u[40] - base), 'after', int(u[41] - base))
    print('hw6 inverse done', [int(v - base) for v in u[42:46]], 'hw0 panel done', [int(v - base) for v in u[46:50]])
    print('hw3 after H', [int(v - base) for v in u[50:54]], 'hw3 before UPD', [int(v - base) for v in u[54:58]])
