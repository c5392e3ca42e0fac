"""Learning check of the whole B200 agent: N maze envs, UNREAL (A3C-LSTM + PC + VR + RP), shared
RMSProp.  Logs finished episodes and mean episode score (+1 goal, -1 per wall hit) per window; a
random policy scores about -25 over a ~110-step episode, the optimal path is 11 steps with score +1.

    python scripts/train_maze.py [envs] [seconds] [lr] [out.jsonl|-] [cells|s2d|f32] [graphs 0|1] [seed]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unreal_b200.environment.environment import Environment
from unreal_b200.model.model import UnrealModel
from unreal_b200.train.rmsprop_applier import RMSPropApplier
from unreal_b200.train.trainer import Trainer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
lr = float(sys.argv[3]) if len(sys.argv) > 3 else 7.0710678e-4
out = open(sys.argv[4], "w") if len(sys.argv) > 4 and sys.argv[4] != "-" else sys.stdout
obs = sys.argv[5] if len(sys.argv) > 5 else "cells"      # cells: render fused into conv1; s2d: K1-rendered x'' planes; f32
graphs = bool(int(sys.argv[6])) if len(sys.argv) > 6 else True
seed0 = int(sys.argv[7]) if len(sys.argv) > 7 else 0xA3C
dev = torch.device("cuda", 0)
net = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                  num_envs=n, seed=0)
applier = RMSPropApplier(lr, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
max_t = 10 ** 10
tr = Trainer(0, net, lr, None, applier, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, 2000, max_t,
             "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0, num_envs=n,
             seeds=np.arange(n) + seed0, use_graphs=graphs, obs_cells=(obs == 'cells'), obs_s2d=(obs == 's2d'))
tr.prepare()
while not tr.experience.is_full():
  tr.process(None, 0)
torch.cuda.synchronize()
t0 = time.perf_counter()
global_t, updates, last = 0, 0, time.perf_counter()
prev = tr.episode_stats.clone()
while time.perf_counter() - t0 < seconds:
  d, _ = tr.process(None, global_t)
  global_t += d * n
  updates += 1
  if updates % 100 == 0:
    cur = tr.episode_stats.clone()
    ep, sc = (cur - prev).tolist()
    prev = cur
    now = time.perf_counter()
    rec = dict(updates=updates, env_steps=global_t, wall_s=now - t0, episodes=ep, mean_score=(sc / ep if ep else None),
               steps_per_episode=(100 * 20 * n / ep if ep else None), env_steps_per_s=100 * 20 * n / (now - last),
               total_loss_per_env=float(tr.last_losses["total"]) / n, grad_norm=float(tr.last_losses["grad_norm"]))
    last = now
    out.write(json.dumps(rec) + "\n"); out.flush()
tr.stop()
