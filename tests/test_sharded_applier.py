"""World-size-2 `gloo` tests (CPU) of the N>1 host logic: env partition, per-env seeds and the
learner's gradient exchange (reduce-scatter -> one-double all-reduce -> shard update -> all-gather,
rmsprop_applier.py of this package).  The two K6 device kernels are replaced by a numpy stand-in
(the oracle's RMSProp restatement) through the applier's `_ops` seam; on a GPU box the same
plumbing runs over NCCL with the real kernels (tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_env_shard_partitions_exactly():
  from unreal_b200.train.sharding import env_seeds, env_shard
  for total in (0, 1, 7, 4096, 65536, 65537):
    for world in (1, 2, 3, 4, 8):
      spans = [env_shard(total, world, r) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      sizes = [hi - lo for lo, hi in spans]
      assert max(sizes) - min(sizes) <= 1
  assert env_shard(65536, 8, 3) == (24576, 32768)
  # seeds depend on the global env index only
  assert env_seeds(0xA3C, 24576, 24579) == [0xA3C + 24576, 0xA3C + 24577, 0xA3C + 24578]
  a = env_seeds(5, *env_shard(10, 2, 0)) + env_seeds(5, *env_shard(10, 2, 1))
  assert a == env_seeds(5, 0, 10)


class _CpuOps(object):
  """numpy stand-in for unreal_grad_sumsq / unreal_rmsprop_update (test only)."""

  @staticmethod
  def grad_sumsq(grad, out):
    out += float((grad.double() ** 2).sum())
    return out

  @staticmethod
  def rmsprop_update(var, rms, mom, grad, sumsq, lr, decay, momentum, eps, clip_norm, grad_scale=1.0, grad_norm=None):
    from oracle import unreal_oracle as O
    g = grad.numpy().astype(np.float32) * np.float32(grad_scale)
    norm = np.sqrt(float(sumsq)) * grad_scale
    if clip_norm > 0:
      g = g * np.float32(clip_norm / max(norm, clip_norm))
    m0 = np.zeros_like(g) if mom is None else mom.numpy()
    v, r, m = O.rmsprop_apply(var.numpy(), rms.numpy(), m0, g, lr, decay, momentum, eps)
    var.copy_(torch.from_numpy(np.asarray(v, np.float32)))
    rms.copy_(torch.from_numpy(np.asarray(r, np.float32)))
    if grad_norm is not None:
      grad_norm.fill_(norm)
    return grad_norm


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from unreal_b200.train.rmsprop_applier import RMSPropApplier
    rs = np.random.RandomState(0)
    p = 64 * 5
    var0 = rs.randn(p).astype(np.float32)
    grads = [rs.randn(p).astype(np.float32) * 30 for _ in range(world)]     # large: clipping is active
    ap = RMSPropApplier(0.01, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
    ap._ops = _CpuOps
    flat = torch.from_numpy(var0.copy())
    for step in range(2):
      norm = ap.apply_flat_to(flat, torch.from_numpy(grads[rank] * (step + 1)), 0.01)
    q.put((rank, flat.numpy().copy(), float(norm)))
  finally:
    dist.destroy_process_group()


def test_sharded_rmsprop_equals_single_process_on_mean_gradient():
  from oracle import unreal_oracle as O
  world, port = 2, 29641
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for pr in procs:
    pr.start()
  results = [q.get(timeout=60) for _ in range(world)]
  for pr in procs:
    pr.join(60)
    assert pr.exitcode == 0
  rs = np.random.RandomState(0)
  p = 64 * 5
  var = rs.randn(p).astype(np.float32)
  grads = [rs.randn(p).astype(np.float32) * 30 for _ in range(world)]
  rms = np.ones(p, np.float32)
  for step in range(2):
    g = sum(grads) * (step + 1) / world                                     # mean of the ranks' gradients
    norm = np.sqrt(float((g.astype(np.float64) ** 2).sum()))
    gc = g * np.float32(40.0 / max(norm, 40.0))
    var, rms, _ = O.rmsprop_apply(var, rms, np.zeros_like(var), gc, 0.01, 0.99, 0.0, 0.1)
  for rank, got, got_norm in results:
    assert np.allclose(got, var, rtol=1e-5, atol=1e-6), rank      # every rank holds the full updated vector
    assert abs(got_norm - norm) <= 1e-4 * norm
