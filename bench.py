#!/usr/bin/env python
"""Benchmark of the GP-Bandit scoring hot path (BASELINE.json metric).

metric   : GP-UCB candidates scored/sec (suggest() acquisition pass)
workload : C2 -- GP posterior mu/var + UCB (+trust region) over M=100k candidates,
           N=1000 trials, D=20, fp64, per GPU (weak scaling: every rank scores its own
           M-candidate shard of one global Philox pool, then one NCCL all-gather picks
           the global arg-max).
step     : one pass of the fused scoring kernel over the rank's M candidates + device top-1.
value    : candidates/s with candidates resident in HBM (CUDA events on the launching stream,
           max over ranks).
e2e      : same pass through the C-ABI call with HOST buffers (pinned): H2D of the M x D
           candidates and D2H of the M scores inside the timed region.
roofline : the scoring kernel is FP64-pipe bound (2N^2/2 flops per candidate against 8(D+1) bytes),
           so `achieved` is algorithmic TFLOP/s against the FP64 peak measured on this GPU by
           tools/fp64_peak.cu (profiles/fp64_peak_r01.json); the HBM view is reported beside it.
cpu_baseline / --impl reference: the NumPy/SciPy oracle (a port: the reference's JAX/TFP stack is
           not installable here) on the host cores, on a bounded candidate sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRIALS, DIM, M_POOL = 1000, 20, 100_000
SEED = 0


def make_problem():
  rng = np.random.default_rng(SEED)
  x = rng.uniform(size=(N_TRIALS, DIM))
  y = -np.sum((x - 0.3) ** 2, axis=1) + 0.05 * rng.normal(size=N_TRIALS)
  ls2 = 0.5 * (1 + np.arange(DIM) / DIM)
  return x, y, dict(sf2=1.0, ls2=ls2, sn2=1e-3)


def algorithmic_flops_per_candidate(n, d):
  # triangular contraction N^2 (N^2/2 FMA) + kernel row N*(3D+25) + mean 2N  
  return n * n + n * (3 * d + 25) + 2 * n


def algorithmic_bytes_per_candidate(d):
  return 8 * d + 8


def fp64_peak_tflops():
  p = os.path.join(ROOT, 'profiles', 'fp64_peak_r01.json')
  if os.path.exists(p):
    try:
      j = json.load(open(p))
      return float(max(j['dfma_tflops'], j['dmma_m8n8k4_tflops'])), 'measured (tools/fp64_peak.cu, profiles/fp64_peak_r01.json)'
    except Exception:  # pylint: disable=broad-except
      pass
  return 37.0, 'nominal B200 FP64 (no measurement found)'


def hbm_peak_gbs():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    try:
      return float(json.load(open(p))['hbm_gbs']), 'measured'
    except Exception:  # pylint: disable=broad-except
      pass
  return 6650.0, 'fallback'


class ClockSampler:
  """nvidia-smi clocks/throttle reasons during the timed region."""

  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.gpu = gpu_index
    self.lines = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:  # pylint: disable=broad-except
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if not self.proc:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:  # pylint: disable=broad-except
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    for ln in self.lines:
      f = [s.strip() for s in ln.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2]))
      except ValueError:
        continue
      for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'samples': len(sm), 'reasons': sorted(reasons)}


def cpu_oracle_rate(sample, reps):
  """Oracle (NumPy/SciPy) candidates/s on a bounded sample of the C2 workload, using every host
  core: candidates are split into 1024-row chunks scored concurrently by a thread pool (NumPy
  releases the GIL), with BLAS pinned to one thread per chunk to avoid oversubscription."""
  from concurrent.futures import ThreadPoolExecutor
  from oracle import gp_oracle as go
  x, y, th = make_problem()
  params = go.GPParams(th['sf2'], th['ls2'], th['sn2'])
  pred = go.precompute_predictive(params, x, y)
  rng = np.random.default_rng(1)
  xs = rng.uniform(size=(sample, DIM))
  cores = os.cpu_count() or 1
  chunks = [xs[i:i + 1024] for i in range(0, sample, 1024)]

  def work(c):
    return go.score_with_aux(pred, c)[0]

  try:
    from threadpoolctl import threadpool_limits
    limiter = threadpool_limits(limits=1, user_api='blas')
  except Exception:  # pylint: disable=broad-except
    limiter = None
  with ThreadPoolExecutor(max_workers=cores) as ex:
    list(ex.map(work, chunks[:cores]))  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
      list(ex.map(work, chunks))
    dt = (time.perf_counter() - t0) / reps
  if limiter is not None:
    limiter.restore_original_limits()
  return sample / dt, dt, cores


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  sample = 16384
  rates = []
  for _ in range(args.warmup):
    cpu_oracle_rate(2048, 1)
  t_all = 0.0
  for _ in range(args.steps):
    r, dt, cores = cpu_oracle_rate(sample, 1)
    rates.append(r); t_all += dt
  v = sample * args.steps / t_all
  line = {
      'impl': 'reference', 'metric': 'GP-UCB candidates scored/sec', 'value': v, 'unit': 'candidates/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_all / args.steps,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': f'C2: GP posterior mu/var + UCB, N={N_TRIALS}, D={DIM}, M={M_POOL} (CPU arm: bounded sample of {sample} candidates per step)'},
      'cpu_baseline': {'value': v, 'unit': 'candidates/s', 'cores': cores, 'kind': 'port',
                       'sample': f'{sample} candidates/step, NumPy/SciPy oracle (reference JAX/TFP build unavailable)'},
      'e2e': {'value': v, 'unit': 'candidates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line), flush=True)


def run_gpu(args):
  import torch
  from vizier_b200 import gp

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  dist = None
  if world > 1:
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  torch.cuda.set_device(local)
  dev = gp.DeviceGP(local)
  x, y, th = make_problem()
  params = gp.GPHyperParams(th['sf2'], th['ls2'], th['sn2'])
  dev.fit(x, y, params)  # every rank recomputes the (deterministic) factorisation: no broadcast needed
  from vizier_b200.multi_gpu import trust_radius, TopkExchange
  acq = gp.Acquisition(1.8, True, trust_radius(N_TRIALS, DIM, 0))

  # rotating candidate pools: 10 x 16 MB = 160 MB > 126 MB L2, so no step re-reads inputs from L2
  n_pools = 10
  pools = [dev.random_pool(M_POOL, DIM, seed=SEED + 17, index_base=(rank * n_pools + i) * M_POOL)
           for i in range(n_pools)]
  outs = [{'score': torch.empty(M_POOL, dtype=torch.float64, device=dev.device)} for _ in range(2)]
  stream = dev.stream
  exchange = TopkExchange(dist, dev, DIM, 1)
  last = {}

  def step(i):
    # one suggest over this rank's shard, all on the handle's stream with no host synchronisation:
    # fused score -> device top-1 -> pack [score, global index, x] -> NCCL all-gather (N > 1) ->
    # deterministic merge kernel -> async D2H of the winner.  The host reads step i-1's winner while
    # step i runs, so the read is inside the timed loop without stalling the GPU.
    slot = i % 2
    exchange.step(slot, pools[i % n_pools], acq, index_base=(rank * n_pools + i % n_pools) * M_POOL,
                  score_out=outs[slot]['score'])
    if 'slot' in last:
      last['winner'] = exchange.result(last['slot'])
    last['slot'] = slot

  for i in range(args.warmup):
    step(i)
  dev.synchronize()
  # ---- kernel-only duration of the dominant kernel (events on the launching stream) ----
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
  if dist is not None:
    dist.barrier()
  torch.cuda.synchronize()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  l0 = dev.launch_count
  t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
  t_start.record(stream)
  for i in range(args.steps):
    step(i)
  t_end.record(stream)
  winner = exchange.result(last['slot'])
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  launches = dev.launch_count - l0
  total_ms = t_start.elapsed_time(t_end)
  # latency of ONE synchronous suggest (enqueue -> winner on the host), median of 5
  lat = []
  for i in range(5):
    t0 = time.perf_counter()
    exchange.step(0, pools[i % n_pools], acq, index_base=(rank * n_pools + i % n_pools) * M_POOL, score_out=outs[0]['score'])
    w_idx, w_val, _ = exchange.result(0)
    lat.append(1e3 * (time.perf_counter() - t0))
  suggest_latency_ms = float(np.median(lat))
  ranks_agree = True
  if dist is not None:   # every rank must have merged the same global winner
    mine = torch.tensor([float(w_idx[0]), float(w_val[0])], dtype=torch.float64, device=dev.device)
    allw = torch.empty((world, 2), dtype=torch.float64, device=dev.device)
    dist.all_gather_into_tensor(allw, mine)
    ranks_agree = bool((allw == allw[0]).all().item())
  # duration of the dominant kernel alone: CUDA events on the launching stream around each launch
  for i in range(args.steps):
    ev[i][0].record(stream)
    dev.score(pools[i % n_pools], acq, out=outs[i % 2])
    ev[i][1].record(stream)
  torch.cuda.synchronize()
  kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))

  # ---- e2e through the host-buffer C-ABI call ----
  host_x = [torch.empty((M_POOL, DIM), dtype=torch.float64).pin_memory() for _ in range(2)]
  host_s = torch.empty(M_POOL, dtype=torch.float64).pin_memory()
  for i in range(2):
    host_x[i].copy_(pools[i].cpu())
  for i in range(max(1, args.warmup)):
    dev.score_host(host_x[i % 2], acq, score_out=host_s)
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  t0 = time.perf_counter()
  for i in range(args.steps):
    dev.score_host(host_x[i % 2], acq, score_out=host_s)
  torch.cuda.synchronize()
  e2e_s = time.perf_counter() - t0
  clocks = sampler.stop() if rank == 0 else None

  if dist is not None:
    t = torch.tensor([total_ms, kern_ms, e2e_s * 1e3], dtype=torch.float64, device=dev.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_ms = [float(v) for v in t.cpu()]
    e2e_s = e2e_ms / 1e3
  if rank != 0:
    if dist is not None:
      dist.destroy_process_group()
    return

  cand_total = M_POOL * world * args.steps
  value = cand_total / (total_ms * 1e-3)
  flops = algorithmic_flops_per_candidate(N_TRIALS, DIM) * M_POOL
  peak, peak_src = fp64_peak_tflops()
  achieved = flops / (kern_ms * 1e-3) * 1e-12
  hbm_peak, hbm_src = hbm_peak_gbs()
  hbm_ach = algorithmic_bytes_per_candidate(DIM) * M_POOL / (kern_ms * 1e-3) * 1e-9
  traffic = None
  tp = os.path.join(ROOT, 'profiles', 'score_kernel_ncu_r01_latest.json')
  if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum of one k_score launch (ncu --set full)
    try:
      j = json.load(open(tp))
      def _b(k):
        v, u = float(j[k]['value']), j[k]['unit'].lower()
        return v * {'gbyte': 1e9, 'mbyte': 1e6, 'kbyte': 1e3, 'byte': 1.0}[u]
      traffic = _b('dram__bytes_read.sum') + _b('dram__bytes_write.sum')
    except Exception:  # pylint: disable=broad-except
      pass
  cpu_v, cpu_dt, cores = (None, None, None)
  if world == 1:
    reps = 3
    cpu_v, cpu_dt, cores = cpu_oracle_rate(16384, reps)
  line = {
      'metric': 'GP-UCB candidates scored/sec', 'value': value, 'unit': 'candidates/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': f'C2: GP posterior mu/var + UCB + trust region + top-1, N={N_TRIALS}, D={DIM}, M={M_POOL} per GPU',
                 'l2': f'{n_pools} rotating candidate pools ({n_pools * M_POOL * DIM * 8 / 1e6:.0f} MB > 126 MB L2)',
                 'parallelism': f'candidate-pool shards x{world}, 1 NCCL all-gather for the global arg-max' if world > 1 else 'single GPU',
                 'suggest_latency_ms': suggest_latency_ms, 'ranks_agree': ranks_agree,
                 'pipelining': 'steps are enqueued without host sync; the host reads winner i-1 while step i runs'},
      'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                   'traffic': traffic, 'kernel': 'k_score', 'kernel_ms': kern_ms,
                   'note': 'fp64: tcgen05 has no f64 kind, the binding roof is the FP64 FMA/DMMA pipe; peak ' + peak_src,
                   'flops_per_candidate': algorithmic_flops_per_candidate(N_TRIALS, DIM),
                   'hbm': {'achieved': hbm_ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': hbm_ach / hbm_peak, 'peak_source': hbm_src}},
      'e2e': {'value': M_POOL * world * args.steps / e2e_s, 'unit': 'candidates/s',
              'h2d_bytes_per_step': M_POOL * DIM * 8 * world, 'd2h_bytes_per_step': M_POOL * 8 * world,
              'ms_per_step': 1e3 * e2e_s / args.steps},
      'gpu_launches': int(launches),
      'clocks': clocks,
  }
  if cpu_v is not None:
    line['cpu_baseline'] = {'value': cpu_v, 'unit': 'candidates/s', 'cores': cores, 'kind': 'port',
                            'sample': f'16384 candidates x 3 reps ({cpu_dt:.2f} s each), NumPy/SciPy oracle on the host'}
  print(json.dumps(line), flush=True)
  if dist is not None:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_gpu(args)


if __name__ == '__main__':
  main()
