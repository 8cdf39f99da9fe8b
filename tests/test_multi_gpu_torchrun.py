"""N > 1 on real GPUs: launches tools/multi_gpu_check.py under torchrun with one rank per GPU (fused
NVLink exchange over CUDA IPC, in-library NCCL fallback, torch.distributed variant - all against a host
merge).  Skipped on single-GPU boxes; the host logic is covered on CPU by test_multi_gpu_cpu.py and the
fused kernel's protocol by test_gpu_exchange.py."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip('torch')
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_transports_under_torchrun():
  n = min(torch.cuda.device_count(), 8)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
         '--master-port', '29631', os.path.join(ROOT, 'tools', 'multi_gpu_check.py')]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert r.returncode == 0 and 'MULTI_GPU_CHECK OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
