"""The fused peer-memory top-k exchange (csrc/exchange.cu) on one GPU: a world of one, and a loop-back
world of several "ranks" that are handles on different streams of the same device (their kernels are
co-resident, so the flag protocol is exercised for real; the NVLink mapping itself is covered by
tests/test_multi_gpu_torchrun.py on boxes with >= 2 GPUs)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason='no CUDA device')]


def _rows(rng, count, width, base, ties=False):
  r = rng.normal(size=(count, width))
  if ties:
    r[:, 0] = np.round(r[:, 0])          # many equal scores: the index tie-break decides
  r[:, 1] = base + rng.permutation(1000)[:count]
  return r


def test_single_rank_exchange_is_the_merge():
  from vizier_b200 import gp, multi_gpu
  dev = gp.DeviceGP(0)
  count, width = 5, 9
  (ex,) = multi_gpu.PeerExchange.local_group([dev], count, width)
  rng = np.random.default_rng(0)
  rows = _rows(rng, count, width, 0)
  rows[2, 0] = np.nan
  pay = torch.from_numpy(rows).cuda()
  out = torch.empty_like(pay)
  host = torch.empty((count, width), dtype=torch.float64).pin_memory()
  dev.stream.wait_stream(torch.cuda.current_stream())
  ex.allgather_topk(pay, out, host)
  assert ex.status() == 0
  wi, wv, wx = multi_gpu.merge_topk(rows[:, 1].astype(np.int64), rows[:, 0], rows[:, 2:], count)
  np.testing.assert_array_equal(host.numpy()[:, 1].astype(np.int64), wi)
  np.testing.assert_array_equal(host.numpy()[:, 2:], wx)
  np.testing.assert_array_equal(out.cpu().numpy(), host.numpy())


@pytest.mark.parametrize('world,count,width,ties', [(2, 1, 22, False), (4, 3, 7, True), (8, 8, 52, False)])
def test_loopback_world_matches_host_merge(world, count, width, ties):
  from vizier_b200 import gp, multi_gpu
  devs = [gp.DeviceGP(0) for _ in range(world)]
  group = multi_gpu.PeerExchange.local_group(devs, count, width)
  rng = np.random.default_rng(world)
  outs = [torch.empty((count, width), dtype=torch.float64, device='cuda') for _ in range(world)]
  hosts = [torch.empty((count, width), dtype=torch.float64).pin_memory() for _ in range(world)]
  for step in range(7):                      # crosses the two-slot wrap several times
    rows = [_rows(rng, count, width, 1000 * r, ties) for r in range(world)]
    pays = [torch.from_numpy(x).cuda() for x in rows]
    torch.cuda.synchronize()
    order = rng.permutation(world)           # ranks are enqueued in a different order every step
    for r in order:
      group[r].allgather_topk(pays[r], outs[r], hosts[r])
    for g in group:
      assert g.status() == 0
    allr = np.concatenate(rows)
    wi, wv, wx = multi_gpu.merge_topk(allr[:, 1].astype(np.int64), allr[:, 0], allr[:, 2:], count)
    for r in range(world):
      got = hosts[r].numpy()
      np.testing.assert_array_equal(got[:, 1].astype(np.int64), wi)
      np.testing.assert_array_equal(got[:, 0], wv)
      np.testing.assert_array_equal(got[:, 2:], wx)
      np.testing.assert_array_equal(outs[r].cpu().numpy(), got)


def test_missing_peer_times_out_instead_of_hanging(monkeypatch):
  from vizier_b200 import gp, multi_gpu
  monkeypatch.setenv('VZGP_EXCHANGE_TIMEOUT_MS', '50')
  devs = [gp.DeviceGP(0) for _ in range(2)]
  group = multi_gpu.PeerExchange.local_group(devs, 2, 6)
  pay = torch.zeros((2, 6), dtype=torch.float64, device='cuda')
  out = torch.empty_like(pay)
  torch.cuda.synchronize()
  group[0].allgather_topk(pay, out)          # rank 1 never calls
  assert group[0].status() == 1
  got = out.cpu().numpy()
  assert np.all(np.isneginf(got[:, 0])) and np.all(got[:, 1] == -1)


def test_topk_exchange_world_of_one_pipeline():
  """TopkExchange (world 1, fused transport) reproduces score + top-k + gather on the host."""
  from oracle import gp_oracle as go
  from vizier_b200 import gp, multi_gpu
  rng = np.random.default_rng(3)
  n, d, m = 200, 6, 3000
  x = rng.uniform(size=(n, d)); y = rng.normal(size=n)
  ls2 = np.full(d, 0.7)
  dev = gp.DeviceGP(0)
  dev.fit(x, y, gp.GPHyperParams(1.0, ls2, 1e-2))
  acq = gp.Acquisition(1.8, True, go.trust_radius(n, d, 0))
  ex = multi_gpu.TopkExchange(None, dev, d, 3, slots=4)
  xs = dev.random_pool(m, d, seed=5)
  for i in range(6):
    ex.step(i % 4, xs, acq, index_base=1000)
  idx, val, feat = ex.result(5 % 4)
  want, _ = go.score_with_aux(go.precompute_predictive(go.GPParams(1.0, ls2, 1e-2), x, y), xs.cpu().numpy())
  top = go.top_k(want, 3)
  np.testing.assert_array_equal(idx, top + 1000)
  np.testing.assert_allclose(val, want[top], atol=1e-10)
  np.testing.assert_array_equal(feat, xs.cpu().numpy()[top])
  pinned = xs.cpu().pin_memory(); hs = torch.empty(m, dtype=torch.float64).pin_memory()
  idx2, val2, _ = ex.suggest_host(pinned, acq, 1000, hs)
  np.testing.assert_array_equal(idx2, idx)
  np.testing.assert_allclose(hs.numpy(), want, atol=1e-10)
