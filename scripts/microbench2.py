"""Per-kernel timings of the small kernels: K1 u8 warp-per-env variants, K1 x'' render, K5 replay ring,
RNG action choice, K6 RMSProp.  JSON lines to stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unreal_b200 import _lib, kernels as K
from microbench import timeit, timeit_graph, report

dev = "cuda:0"
_lib.require_device()
N, T = 4096, 20
st = K.MazeState(N, dev)
g = torch.Generator(device=dev).manual_seed(0)
acts = torch.randint(0, 4, (T, N), device=dev, dtype=torch.int32, generator=g)
pc = torch.empty(T, N, 20, 20, device=dev); rew = torch.empty(T, N, device=dev)
term = torch.empty(T, N, dtype=torch.uint8, device=dev)
obs8 = torch.empty(T, N, 84, 84, 3, dtype=torch.uint8, device=dev)
for w in (2, 4, 8, 16):
  _lib.set_tunable("maze_warps_per_cta", w)

  def rollout():
    for t in range(T):
      K.maze_step(st, acts[t], obs=obs8[t], pc=pc[t], reward=rew[t], terminal=term[t], auto_reset=True)
  tm, tn = timeit_graph(rollout, reps=1, iters=10)
  report("maze_step u8 warp-per-env", N * (21168 + 1600 + 41), tm / T, tn / T, warps_per_cta=w, n=N)
_lib.set_tunable("maze_warps_per_cta", 4)
del obs8
obsb = torch.empty(T, N, 6, 441, 8, dtype=torch.bfloat16, device=dev)


def rollout_b():
  for t in range(T):
    K.maze_step(st, acts[t], obs=obsb[t], pc=pc[t], reward=rew[t], terminal=term[t], auto_reset=True)
tm, tn = timeit_graph(rollout_b, reps=1, iters=10)
report("maze_step x'' (bf16 planes)", N * (42336 + 1600 + 41), tm / T, tn / T, n=N)
del obsb

# K5: replay ring at config-3 size
n_envs, H = 8192, 2000
ring = K.ReplayRing(n_envs, H, dev)
ms = K.MazeState(n_envs, dev)
rec = torch.zeros(n_envs, dtype=torch.int64, device=dev)
a8 = torch.randint(0, 4, (n_envs,), device=dev, dtype=torch.int32, generator=g)
for _ in range(H + 10):
  K.maze_step(ms, a8, frame_rec=rec, auto_reset=True)
  ring.add(rec)
streams = K.MtStreams(np.arange(n_envs) + 1, dev)
tm, tn = timeit_graph(lambda: ring.add(rec))
report("replay_add", n_envs * 16, tm, tn, envs=n_envs, history=H)
tm, tn = timeit(lambda: ring.sample_sequence(streams, 21), iters=20)
report("replay_sample_sequence(21)", n_envs * (21 * 8 * 2 + 16), tm, tn, envs=n_envs, history=H)
tm, tn = timeit(lambda: ring.sample_rp(streams), iters=20)
report("replay_sample_rp (rank-select over 1997 records)", n_envs * (H * 8 + 48), tm, tn, envs=n_envs, history=H)
pi = torch.softmax(torch.randn(n_envs, 4, device=dev), -1)
tm, tn = timeit(lambda: streams.choose_action(pi), iters=20)
report("choose_action (MT19937 choice)", n_envs * (16 + 4 + 8), tm, tn, envs=n_envs)

# K6
for P in (1898880, 64 * 1024 * 1024):
  var = torch.randn(P, device=dev); rms = torch.ones(P, device=dev); grad = torch.randn(P, device=dev)
  ss = torch.zeros(1, dtype=torch.float64, device=dev)

  def upd():
    ss.zero_()
    K.grad_sumsq(grad, ss)
    K.rmsprop_update(var, rms, None, grad, ss, 7e-4, 0.99, 0.0, 0.1, 40.0)
  tm, tn = timeit_graph(upd, reps=4)
  report("grad_sumsq + rmsprop_update", P * 24, tm, tn, params=P)
