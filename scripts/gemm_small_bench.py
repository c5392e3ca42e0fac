"""Latency of the small sequential GEMMs of the LSTM (one per step) under different tile widths / K splits:
forward step [N, 520] x [520, 1024] (+ bias), backward recurrent gradient [N, 1024] x [256, 1024]^T."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import kernels as K, _lib
dev = torch.device("cuda", 0)


def bench(fn, iters=50):
  for _ in range(5):
    fn()
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    for _ in range(iters):
      fn()
  g.replay(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e3


for n_envs in (1024, 2048, 8192):
  dg = torch.randn(n_envs, 1024, device=dev).to(torch.bfloat16)
  wcat = torch.randn(520, 1024, device=dev).to(torch.bfloat16)
  wh = wcat[264:]
  xh = torch.randn(n_envs, 520, device=dev).to(torch.bfloat16)
  bias = torch.randn(1024, device=dev)
  gates = torch.zeros(n_envs, 1024, device=dev)
  out = torch.empty(n_envs, 256, device=dev)
  for bn in (0, 64, 128, 256):
    _lib.set_tunable("gemm_bn", bn)
    t = bench(lambda: K.gemm_bf16(xh, wcat, out=gates, b_mn_major=True, bias=bias))
    print(json.dumps(dict(shape="lstm fwd step [N,520]x[520,1024]+b", envs=n_envs, bn=bn, us=round(t, 2))))
    for sk in (1, 2, 4):
      o = torch.zeros(n_envs, 256, device=dev) if sk > 1 else out
      t = bench(lambda: K.gemm_bf16(dg, wh, out=o, split_k=sk))
      print(json.dumps(dict(shape="lstm bwd dh [N,1024]x[256,1024]^T", envs=n_envs, bn=bn, split_k=sk, us=round(t, 2))))
_lib.set_tunable("gemm_bn", 0)
