"""K7 GEMM micro-benchmark on the model's shapes (config 2/3 per-GPU sizes): TFLOP/s of
unreal_gemm_bf16 next to torch.matmul (cuBLAS) on the same operands.  Writes JSON lines."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import kernels as K

dev = torch.device("cuda", 0)
SHAPES = [
    # name, m, n, k, a_mn, b_mn, split_k
    ("fc1 fwd  [S,2592]x[2592,256]", 81920, 256, 2592, False, True, 1),
    ("lstm-in  [S,264]x[264,1024]", 81920, 1024, 264, False, True, 1),
    ("lstm-rec [N,256]x[256,1024]", 4096, 1024, 256, False, True, 1),
    ("pc_fc1   [S,256]x[256,2592]", 81920, 2592, 256, False, True, 1),
    ("fc1 dgrad [S,256]x[2592,256]^T", 81920, 2592, 256, False, False, 1),
    ("fc1 wgrad X^T dY", 2592, 256, 81920, True, True, 16),
    ("pc_fc1 wgrad", 256, 2592, 81920, True, True, 16),
    ("conv1 im2col [S*400,192]x[16,192]^T", 4096 * 400, 16, 192, False, False, 1),
    ("conv2 im2col [S*81,256]x[32,256]^T", 81920 * 81 // 4, 32, 256, False, False, 1),
    ("square 8192", 8192, 8192, 8192, False, False, 1),
    ("square 4096", 4096, 4096, 4096, False, False, 1),
]


def bench(fn, iters=None):
  iters = iters or ITERS
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
ONLY = os.environ.get("GEMM_ONLY")          # substring filter, e.g. GEMM_ONLY="fc1 fwd"
ITERS = int(os.environ.get("GEMM_ITERS", "20"))
for name, m, n, k, a_mn, b_mn, sk in SHAPES:
  if ONLY and ONLY not in name:
    continue
  a = torch.randn((k, m) if a_mn else (m, k), device=dev).to(torch.bfloat16)
  b = torch.randn((k, n) if b_mn else (n, k), device=dev).to(torch.bfloat16)
  c = torch.zeros(m, n, device=dev, dtype=torch.float32 if sk > 1 else torch.bfloat16)
  t = bench(lambda: K.gemm_bf16(a, b, out=c, a_mn_major=a_mn, b_mn_major=b_mn, split_k=sk))
  A = a.t() if a_mn else a
  B = b if b_mn else b.t()
  cb = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
  tb = bench(lambda: torch.matmul(A, B, out=cb))
  flops = 2.0 * m * n * k
  byts = 2.0 * (m * k + n * k) + c.element_size() * m * n
  rec = dict(shape=name, m=m, n=n, k=k, us=t * 1e6, tflops=flops / t / 1e12, gbs=byts / t / 1e9,
             cublas_us=tb * 1e6, cublas_tflops=flops / tb / 1e12)
  out.write(json.dumps(rec) + "\n")
  out.flush()
