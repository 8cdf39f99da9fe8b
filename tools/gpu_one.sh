#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ard_fit or retry_sem or posterior" 2>&1 | tail -40 | tee gpurun_out/one.log
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from vizier_b200 import gp
import bench
dev=gp.DeviceGP(0)
for n in (64,128,256,512,1024,2048):
    rng=np.random.default_rng(0); x=rng.uniform(size=(n,20)); y=rng.normal(size=n)
    p=gp.GPHyperParams(1.0, np.full(20,1.0), 1e-2)
    xt=torch.from_numpy(x).cuda(); yt=torch.from_numpy(y).cuda()
    dev.fit(xt,yt,p); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(20): dev.fit(xt,yt,p)
    torch.cuda.synchronize(); t=(time.perf_counter()-t0)/20
    t0=time.perf_counter()
    for _ in range(20): dev.loss_and_grad(xt,yt,p)
    torch.cuda.synchronize(); t2=(time.perf_counter()-t0)/20
    print(f'N={n}: fit {t*1e3:.3f} ms  nll_grad {t2*1e3:.3f} ms')
PY
