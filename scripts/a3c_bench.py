"""BASELINE configs[4] slice per GPU: MINOS-shaped synthetic observations (uint8 84x84 RGB; the
reference model consumes no depth channel), 3 pointgoal actions, goal vector G = 2,
A3C-LSTM forward/backward + fused clip+RMSProp on 1024 envs x T = 20 = 20 480 samples per update.
Reports samples/s and the model's algorithmic FLOP rate (18.5 MFLOP/sample)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unreal_b200 import _lib
from unreal_b200.model.model import UnrealModel
from unreal_b200.train.rmsprop_applier import RMSPropApplier

import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:      # configs[4] exactly: `torchrun --nproc-per-node 8 scripts/a3c_bench.py 1024` = batch 8192 x T over 8 GPUs,
  dist.init_process_group("nccl", device_id=dev)     # gradients exchanged by the sharded RMSProp (reduce-scatter / all-gather)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T, A, G = 20, 3, 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m = UnrealModel(A, G, -1, True, False, False, False, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                num_envs=N, seed=0)
ap = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
g = torch.Generator(device=dev).manual_seed(0)
img = torch.randint(0, 256, (T, N, 84, 84, 3), dtype=torch.uint8, device=dev, generator=g)
lar = torch.zeros(T, N, A + 1 + G, device=dev)
lar.scatter_(2, torch.randint(0, A, (T, N, 1), device=dev, generator=g), 1.0)
lar[..., A + 1:] = torch.rand(T, N, G, device=dev, generator=g)
a = torch.zeros(T, N, A, device=dev).scatter_(2, torch.randint(0, A, (T, N, 1), device=dev, generator=g), 1.0)
feed = {"base": dict(images=img, lar=lar, a=a, adv=torch.randn(T, N, device=dev, generator=g),
                     R=torch.randn(T, N, device=dev, generator=g), mask=torch.ones(T, N, device=dev),
                     c0=torch.zeros(N, 256, device=dev), h0=torch.zeros(N, 256, device=dev))}
graph = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
if world > 1:
  graph = False      # capturing the NCCL reduce-scatter / all-gather into the graph deadlocked here (2 GPUs, NCCL 2.28.9): eager
lr = torch.full((1,), 7e-4, device=dev)      # device scalar: K6 reads it when it runs, so a graph follows the anneal
for _ in range(3):
  out = m.update(feed, lr, ap)
torch.cuda.synchronize()
if graph:                                    # the whole update (fwd, bwd, clip + RMSProp, shadow refresh) as ONE CUDA graph
  g = torch.cuda.CUDAGraph()
  with _lib.graph_capture(g):
    out = m.update(feed, lr, ap)
  step = g.replay
else:
  def step():
    global out
    out = m.update(feed, lr, ap)
for _ in range(2):
  step()
if world > 1:
  dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
  step()
e1.record()
if world > 1:
  dist.barrier()
torch.cuda.synchronize()
t_ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
if world > 1:
  dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
ms = float(t_ms)
samples = N * T * world
if world > 1 and dist.get_rank() != 0:
  dist.destroy_process_group()
  sys.exit(0)
print(json.dumps(dict(n_gpus=world, cuda_graph=graph, workload="configs[4] slice: %d envs x T=20, u8 84x84x3 frames, A=3, G=2, A3C-LSTM fwd/bwd + RMSProp" % N,
                      ms_per_update=ms, samples_per_s=samples / (ms * 1e-3), model_tflops=samples * 18.5e6 / (ms * 1e-3) / 1e12,
                      frac_of_sustained_bf16_peak_per_gpu=samples * 18.5e6 / (ms * 1e-3) / 1e12 / 1393.1 / world,
                      finite=bool(torch.isfinite(out["total"])), params=m.num_parameters)))
if world > 1:
  dist.destroy_process_group()
