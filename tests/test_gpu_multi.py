"""Two-GPU (NCCL) checks of the only exchange step on the path, the learner gradient
(SURVEY.md 8e): sharded clip + RMSProp equals the single-GPU kernel on the mean gradient, and two
data-parallel learners stay bit-identical.  Skipped on a one-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "tests"))
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  import torch.distributed as dist
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  try:
    from unreal_b200.train.rmsprop_applier import RMSPropApplier
    from unreal_b200.model.model import UnrealModel
    from test_gpu_model import _feed, _to
    # (1) sharded applier on a synthetic flat buffer
    rs = np.random.RandomState(0)
    p = 1898880
    var0 = rs.randn(p).astype(np.float32)
    grads = [rs.randn(p).astype(np.float32) for _ in range(world)]
    ap = RMSPropApplier(0.01, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
    flat = torch.from_numpy(var0.copy()).to(dev)
    norm = ap.apply_flat_to(flat, torch.from_numpy(grads[rank]).to(dev), 0.01)
    res = dict(rank=rank, flat=flat.cpu().numpy(), norm=float(norm))
    # (2) two data-parallel learners: same init, different env data, one update
    m = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0,
                    0.0, num_envs=2, seed=7)
    ap2 = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
    feed = _to(_feed(4, 2, 3, seed=20 + rank), dev)
    _, _, g = m.loss_and_grads(feed, 1.0 / 2)
    res["grad"] = g.detach().cpu().numpy().copy()
    res["before"] = m.flat.detach().cpu().numpy().copy()
    m.update(feed, 7e-4, ap2)
    res["after"] = m.flat.detach().cpu().numpy().copy()
    # (3) the whole agent under NCCL: CUDA-graph replay of the update (graph A fwd+bwd, eager exchange + K6, graph B shadow
    # refresh -- Trainer._update) against eager launches, same seeds, a few iterations
    from unreal_b200.train.trainer import Trainer
    from unreal_b200.environment.environment import Environment
    from unreal_b200.train.sharding import env_seeds, env_shard
    params = []
    for graphs in (False, True):
      Environment.action_size = -1
      n_env = 4
      lo, hi = env_shard(n_env * world, world, rank)
      net = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0,
                        0.0, num_envs=n_env, seed=3)
      ap3 = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
      tr = Trainer(rank, net, 7e-4, None, ap3, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, 40,
                   10 ** 9, str(dev), {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
                   num_envs=n_env, seeds=env_seeds(0xA3C, lo, hi), use_graphs=graphs, obs_cells=True)
      tr.prepare()
      while not tr.experience.is_full():
        tr.process(None, 0)
      steps = 0
      for it in range(5):
        d, _ = tr.process(None, it * 1000000)
        steps += d
      torch.cuda.synchronize()
      params.append(net.flat.detach().cpu().numpy().copy())
      res["agent_graph_%d" % int(graphs)] = dict(steps=steps, ugraph=tr._ugraph is not None, split=tr._ugraph_b is not None,
                                                  norm=float(tr.last_losses["grad_norm"]))
      tr.stop()
    res["agent_params"] = params
    torch.cuda.synchronize()
    q.put(res)
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_learner_exchange():
  import torch.multiprocessing as mp
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, 29655, q)) for r in range(world)]
  for pr in procs:
    pr.start()
  results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r["rank"])
  for pr in procs:
    pr.join(60)
    assert pr.exitcode == 0
  # (1) against the single-GPU kernel on the mean gradient
  from unreal_b200 import kernels as K
  dev = torch.device("cuda", 0)
  rs = np.random.RandomState(0)
  p = 1898880
  var = torch.from_numpy(rs.randn(p).astype(np.float32)).to(dev)
  grads = [rs.randn(p).astype(np.float32) for _ in range(world)]
  g = torch.from_numpy(sum(grads) / world).to(dev)
  rms = torch.ones(p, device=dev)
  K.rmsprop_update(var, rms, None, g, K.grad_sumsq(g), 0.01, 0.99, 0.0, 0.1, 40.0)
  want = var.cpu().numpy()
  for r in results:
    assert np.allclose(r["flat"], want, rtol=1e-5, atol=1e-6)
    assert abs(r["norm"] - float(torch.sqrt((g.double() ** 2).sum()).cpu())) <= 1e-3
  # (2) both learners applied the same (mean) gradient: identical parameters afterwards
  assert np.array_equal(results[0]["after"], results[1]["after"])
  assert np.array_equal(results[0]["before"], results[1]["before"])
  gm = (results[0]["grad"].astype(np.float64) + results[1]["grad"].astype(np.float64)) / world
  norm = np.sqrt((gm ** 2).sum())
  gc = gm * (40.0 / max(norm, 40.0))
  ms = 1.0 + (gc * gc - 1.0) * 0.01
  want2 = results[0]["before"] - 7e-4 * gc / np.sqrt(ms + 0.1)
  assert np.allclose(results[0]["after"], want2, rtol=1e-5, atol=1e-7)
  # (3) graph-replayed learner == eager learner on every rank, identical parameters across ranks
  for r in results:
    assert r["agent_graph_1"]["ugraph"] and r["agent_graph_1"]["split"] and not r["agent_graph_0"]["ugraph"]
    assert r["agent_graph_0"]["steps"] == r["agent_graph_1"]["steps"]
    eager, graphed = r["agent_params"]
    assert np.allclose(eager, graphed, rtol=1e-3, atol=1e-5)
  assert np.array_equal(results[0]["agent_params"][1], results[1]["agent_params"][1])
