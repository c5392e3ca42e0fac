"""One LSTM step of the learner, forward and backward: the fused step kernels (csrc/lstm_tcgen05.cu) against the GEMM +
cell kernel pairs they replace, CUDA-event timed over a chain of T steps like the unroll (buffers rotate, so every
step's operands come from L2 / HBM as in the update).

    python scripts/lstm_step_bench.py [envs ...] > profiles/r2_lstm_step_bench.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unreal_b200 import _lib, kernels as K

dev = torch.device("cuda", 0)
T, KX, KC = 20, 264, 520


def timed(fn, reps=20):
  """the chain as ONE CUDA graph (what the update replays): device time, no host launch overhead"""
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    fn()
  fn = graph.replay
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) * 1e3 / reps / T     # us per step


for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 8192]:
  g = torch.Generator(device=dev).manual_seed(0)
  xh = torch.randn(T + 1, n, KC, device=dev, generator=g).to(torch.bfloat16)
  w = (torch.randn(KC, 1024, device=dev, generator=g) * 0.05).to(torch.bfloat16)
  wh = w[KX:]
  b = torch.zeros(1024, device=dev)
  gates = torch.empty(T, n, 1024, device=dev, dtype=torch.bfloat16)
  c_all = torch.zeros(T + 1, n, 256, device=dev)
  h_all = torch.empty(T, n, 256, device=dev)
  dh_all = torch.randn(T, n, 256, device=dev, generator=g) * 0.1
  dgates = torch.empty(T, n, 1024, device=dev, dtype=torch.bfloat16)
  dc = torch.zeros(n, 256, device=dev)

  def fwd_pair():
    for i in range(T):
      K.gemm_bf16(xh[i], w, out=gates[i], b_mn_major=True, bias=b)
      K.lstm_cell_fwd(gates[i], c_all[i], c_all[i + 1], h_all[i], xh[i + 1, :, KX:])

  def fwd_fused(tiled=False):
    for i in range(T):
      K.lstm_step_fwd(xh[i], w, b, c_all[i], c_all[i + 1], h_out=h_all[i], h16_out=xh[i + 1, :, KX:], acts=gates[i], tiled=tiled)

  def bwd_pair():
    dh_rec = None
    for i in range(T - 1, -1, -1):
      K.lstm_cell_bwd(gates[i], c_all[i], c_all[i + 1], dh_all[i], dc, dgates[i], dh_rec)
      dh_rec = K.gemm_bf16(dgates[i], wh, split_k=4 if n <= 2048 else 1)

  def bwd_fused(tiled=False):
    for i in range(T - 1, -1, -1):
      K.lstm_step_bwd(dgates[i + 1] if i < T - 1 else None, wh, gates[i], c_all[i], c_all[i + 1], dh_all[i], dc, dgates[i], tiled=tiled)

  row = {"envs": n, "unit": "us per step"}
  row["fwd_gemm_plus_cell"] = round(timed(fwd_pair), 2)
  row["fwd_fused"] = round(timed(fwd_fused), 2)
  row["bwd_cell_plus_gemm"] = round(timed(bwd_pair), 2)
  row["bwd_fused"] = round(timed(bwd_fused), 2)
  if n % 32 == 0:       # the tiled layout is a permutation of the same buffers here
    row["fwd_fused_tiled"] = round(timed(lambda: fwd_fused(True)), 2)
    row["bwd_fused_tiled"] = round(timed(lambda: bwd_fused(True)), 2)
  print(json.dumps(row), flush=True)
