"""Kernel-time breakdown of the configs[4] update (A3C-LSTM on u8 indoor-shaped frames, 1024 envs x T = 20), eager launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from unreal_b200.model.model import UnrealModel
from unreal_b200.train.rmsprop_applier import RMSPropApplier

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/a3c_profile.txt"
T, A, G = 20, 3, 2
m = UnrealModel(A, G, -1, True, False, False, False, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                num_envs=N, seed=0)
ap = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
g = torch.Generator(device=dev).manual_seed(0)
img = torch.randint(0, 256, (T, N, 84, 84, 3), dtype=torch.uint8, device=dev, generator=g)
lar = torch.zeros(T, N, A + 1 + G, device=dev)
lar.scatter_(2, torch.randint(0, A, (T, N, 1), device=dev, generator=g), 1.0)
a = torch.zeros(T, N, A, device=dev).scatter_(2, torch.randint(0, A, (T, N, 1), device=dev, generator=g), 1.0)
feed = {"base": dict(images=img, lar=lar, a=a, adv=torch.randn(T, N, device=dev, generator=g),
                     R=torch.randn(T, N, device=dev, generator=g), mask=torch.ones(T, N, device=dev),
                     c0=torch.zeros(N, 256, device=dev), h0=torch.zeros(N, 256, device=dev))}
lr = torch.full((1,), 7e-4, device=dev)
for _ in range(3):
  m.update(feed, lr, ap)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
  for _ in range(4):
    m.update(feed, lr, ap)
  torch.cuda.synchronize()
with open(out, "w") as f:
  f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
print("written", out)
