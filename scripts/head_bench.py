"""Policy / value head kernels (unreal_a3c_head_loss / _bwd) at the agent's two sizes: the acting step (8192 rows, pi and v
out) and the update (163 840 rows, losses + gradients).  CUDA-event timed as graph replays.

    python scripts/head_bench.py > profiles/r2_head_bench.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unreal_b200 import kernels as K

dev = torch.device("cuda", 0)
A = 4


def timed(fn, reps=50):
  fn(); torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    for _ in range(10):
      fn()
  for _ in range(3):
    g.replay()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    g.replay()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) * 1e3 / reps / 10


gen = torch.Generator(device=dev).manual_seed(0)
wp = torch.randn(256, A, device=dev, generator=gen) * 0.05; bp = torch.zeros(A, device=dev)
wv1 = torch.randn(256, device=dev, generator=gen) * 0.05; bv = torch.zeros(1, device=dev)
for m in (8192, 163840):
  # several buffers in rotation so that the rows come from HBM like in the update (168 MB each at 163 840 rows)
  hs = [torch.randn(m, 256, device=dev, generator=gen) for _ in range(3)]
  act = torch.randint(0, A, (m,), device=dev, dtype=torch.int32, generator=gen)
  adv = torch.randn(m, device=dev, generator=gen); R = torch.randn(m, device=dev, generator=gen)
  mask = torch.ones(m, device=dev)
  it = [0]

  def acting():
    it[0] += 1
    K.a3c_head(hs[it[0] % 3], wp, bp, wv1, bv, want_pi=True, want_v=True)

  def loss():
    it[0] += 1
    K.a3c_head(hs[it[0] % 3], wp, bp, wv1, bv, act, adv, R, mask, 0.001, 0.5, want_sums=True, want_grads=True)

  out = K.a3c_head(hs[0], wp, bp, wv1, bv, act, adv, R, mask, 0.001, 0.5, want_sums=True, want_grads=True)
  dz, dv = out["dz"], out["dv"]
  go2 = torch.ones(2, device=dev)

  def bwd():
    it[0] += 1
    K.a3c_head_bwd(hs[it[0] % 3], wp, wv1, dz, dv, go2)

  row = {"rows": m, "unit": "us", "acting_pi_v": round(timed(acting), 2), "loss_and_grad": round(timed(loss), 2),
         "bwd": round(timed(bwd), 2), "h_megabytes": round(m * 1024 / 1e6, 1)}
  print(json.dumps(row), flush=True)
