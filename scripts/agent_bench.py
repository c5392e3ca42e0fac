"""Full-agent throughput (BASELINE config 3, one GPU's slice): N maze envs, per-env replay ring,
PC/VR/RP sampling, UnrealModel forward/backward on tcgen05, fused clip+RMSProp.
Prints JSON lines with per-phase device times."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  H = int(sys.argv[2]) if len(sys.argv) > 2 else 200
  iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
  from unreal_b200.environment.environment import Environment
  from unreal_b200.model.model import UnrealModel
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  from unreal_b200.train.trainer import Trainer
  dev = torch.device("cuda", 0)
  net = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0,
                    0.0, num_envs=n, seed=0)
  applier = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr = Trainer(0, net, 7e-4, None, applier, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, H,
               10 ** 8, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
               num_envs=n, seeds=np.arange(n) + 11, use_graphs=os.environ.get("AGENT_GRAPHS", "1") != "0",
               obs_s2d=os.environ.get("AGENT_S2D", "1") != "0")
  tr.prepare()
  t0 = time.perf_counter()
  while not tr.experience.is_full():
    tr.process(None, 0)
  torch.cuda.synchronize()
  fill_s = time.perf_counter() - t0
  for _ in range(2):
    tr.process(None, 0)
  torch.cuda.synchronize()
  # phase timing: rollout+targets (process minus update) vs update
  orig_update = net.update
  ev = []

  def timed_update(feed, lr, ap):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig_update(feed, lr, ap); e1.record(); ev.append((e0, e1))
    return out

  net.update = timed_update
  s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize(); w0 = time.perf_counter(); s0.record()
  steps = 0
  for _ in range(iters):
    d, _ = tr.process(None, 0)
    steps += d
  s1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - w0
  dev_ms = s0.elapsed_time(s1)
  upd_ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev) if ev else None     # graph replay never re-enters net.update
  print(json.dumps(dict(envs=n, history=H, iters=iters, fill_s=fill_s, wall_ms_per_update=wall / iters * 1e3,
                        device_ms_per_update=dev_ms / iters, model_update_ms=upd_ms,
                        env_steps_per_s=steps / wall, losses={k: float(v) for k, v in tr.last_losses.items()},
                        mem_gb=torch.cuda.max_memory_allocated() / 1e9)))


if __name__ == "__main__":
  main()
