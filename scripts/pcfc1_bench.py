"""pc_fc1 forward [S,256]x[256,2592] (+bias, ReLU, bf16 out) and dgrad [S,2592]x[2592,256]^T over the GEMM tunables."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import _lib, kernels as K
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
x = torch.randn(S, 256, device=dev, generator=g).to(torch.bfloat16)
w = (torch.randn(256, 2592, device=dev, generator=g) * 0.05).to(torch.bfloat16)
b = torch.zeros(2592, device=dev)
dy = torch.randn(S, 2592, device=dev, generator=g).to(torch.bfloat16)
out = torch.empty(S, 2592, device=dev, dtype=torch.bfloat16)
def timed(fn, reps=10):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fn()
  e.record(); torch.cuda.synchronize()
  return a.elapsed_time(e) * 1e3 / reps
for two in (0, 1):
  for bn in (0, 128):
    _lib.set_tunable("gemm_deep_epilogue", two); _lib.set_tunable("gemm_bn", bn)
    f = timed(lambda: K.gemm_bf16(x, w, out=out, b_mn_major=True, bias=b, relu=True))
    d = timed(lambda: K.gemm_bf16(dy, w, out_dtype=torch.float32))
    print(json.dumps({"S": S, "deep_epilogue": two, "gemm_bn": bn, "fwd_us": round(f, 1), "fwd_tflops": round(2.0 * S * 256 * 2592 / f / 1e6),
                      "dgrad_us": round(d, 1)}), flush=True)
