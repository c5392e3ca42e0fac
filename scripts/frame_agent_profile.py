"""Kernel-time breakdown of the framed UNREAL agent (synthetic indoor-shaped u8 frames, FrameTrainer), eager launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from unreal_b200.environment.environment import Environment
from unreal_b200.model.model import UnrealModel
from unreal_b200.train.rmsprop_applier import RMSPropApplier
from unreal_b200.train.trainer import Trainer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hist = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/frame_agent_profile.txt"
dev = torch.device("cuda", 0)
Environment.action_size = -1
net = UnrealModel(3, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                  num_envs=n, seed=0)
applier = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
tr = Trainer(0, net, 7e-4, None, applier, 'synthetic', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, hist,
             10 ** 8, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0, num_envs=n,
             seeds=np.arange(n) + 11, env_args={'producer': 'table', 'seed': 0}, use_graphs=False)
tr.prepare()
while not tr.experience.is_full():
  tr.process(None, 0)
for _ in range(2):
  tr.process(None, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
  for _ in range(2):
    tr.process(None, 0)
  torch.cuda.synchronize()
with open(out, "w") as f:
  f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=50, max_name_column_width=90))
print("written", out)
