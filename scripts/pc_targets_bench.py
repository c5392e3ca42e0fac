"""Trainer._process_pc's target kernels at the agent's size (T = 20, N = 8192): maze_pixel_change + pc_targets against the one-pass
maze_pc_targets.  us per call."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import kernels as K
dev = torch.device("cuda", 0)
T, N = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device=dev).manual_seed(0)
p0 = torch.randint(0, 7, (T, N, 2), device=dev, generator=g, dtype=torch.int32)
p1 = (p0 + torch.randint(-1, 2, (T, N, 2), device=dev, generator=g, dtype=torch.int32)).clamp_(0, 6)
boot = torch.rand(N, 20, 20, device=dev, generator=g)
ln = torch.full((N,), T, device=dev, dtype=torch.int32)
pc = torch.empty(T, N, 20, 20, device=dev); out = torch.empty(T, N, 20, 20, device=dev)
def timed(fn, reps=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fn()
  e.record(); torch.cuda.synchronize()
  return round(a.elapsed_time(e) * 1e3 / reps, 1)
res = {"T": T, "N": N, "unit": "us"}
res["maze_pixel_change"] = timed(lambda: K.maze_pixel_change(p0.view(-1, 2), p1.view(-1, 2), out=pc.view(-1, 20, 20)))
res["pc_targets"] = timed(lambda: K.pc_targets(pc, None, ln, boot, 0.9, out=out))
res["maze_pc_targets"] = timed(lambda: K.maze_pc_targets(p0, p1, ln, boot, 0.9, out=out))
res["maze_pc_targets_gbs"] = round(T * N * 1600 / res["maze_pc_targets"] / 1e3)
print(json.dumps(res))
