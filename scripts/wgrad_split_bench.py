"""Filter-gradient GEMMs of the agent update (dW = X^T dY, K = samples) over split-K and the CTA-pair switch.

    python scripts/wgrad_split_bench.py > profiles/r2_wgrad_split_bench.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unreal_b200 import _lib, kernels as K

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)


def timed(fn, reps=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) * 1e3 / reps


for name, m, n, s in (("lstm [520,S]x[S,1024]", 520, 1024, 163840), ("pc_fc1 [256,S]x[S,2592]", 256, 2592, 163840),
                      ("lstm at 1024 envs", 520, 1024, 20480), ("fc1 [2592,S]x[S,256]", 2592, 256, 20480)):
  x = torch.randn(s, m, device=dev, generator=gen).to(torch.bfloat16)
  dy = torch.randn(s, n, device=dev, generator=gen).to(torch.bfloat16)
  for two in (0, 1):
    _lib.set_tunable("gemm_2sm", two)
    for sk in (2, 3, 4, 5, 6, 7, 8, 10, 12, 16):
      us = timed(lambda: K.gemm_bf16(x, dy, a_mn_major=True, b_mn_major=True, split_k=sk))
      print(json.dumps({"shape": name, "samples": s, "gemm_2sm": two, "split_k": sk, "us": round(us, 1),
                        "tflops": round(2.0 * m * n * s / us / 1e6, 1)}), flush=True)
  _lib.set_tunable("gemm_2sm", -1)
