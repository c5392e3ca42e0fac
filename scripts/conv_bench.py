"""Encoder convolutions: fused TMA-im2col kernels vs explicit im2col + GEMM, GB/s against the
algorithmic bytes (frame in + activation out)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import kernels as K

dev = torch.device("cuda", 0)


def bench(fn, iters=20):
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


S = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
w1 = (torch.randn(8, 8, 3, 16, device=dev) * 0.05).to(torch.bfloat16); b1 = torch.zeros(16, device=dev)
w2 = (torch.randn(4, 4, 16, 32, device=dev) * 0.05).to(torch.bfloat16); b2 = torch.zeros(32, device=dev)
t1, t2 = K.conv1_w_planes(w1), K.conv_taps(w2, 2)
for dt in (torch.float32, torch.uint8):
  x = torch.rand(S, 84, 84, 3, device=dev) if dt == torch.float32 else torch.randint(0, 256, (S, 84, 84, 3), device=dev, dtype=torch.uint8)
  xs = torch.empty(S, 6, 441, 8, dtype=torch.bfloat16, device=dev)
  h1 = torch.empty(S, 20, 20, 16, dtype=torch.bfloat16, device=dev)
  h2 = torch.empty(S, 9, 9, 32, dtype=torch.bfloat16, device=dev)
  esz = x.element_size()
  recs = []
  t = bench(lambda: K.s2d_frames(x, xs)); recs.append(("s2d_frames", t, S * (21168 * esz + 42336)))
  t = bench(lambda: K.conv_fwd(xs, 1, t1, b1, h1)); recs.append(("conv1 fused (x' -> h1)", t, S * (42336 + 12800)))
  t = bench(lambda: K.conv_fwd(h1, 2, t2, b2, h2)); recs.append(("conv2 fused (h1 -> h2)", t, S * (12800 + 5184)))
  t = bench(lambda: (K.s2d_frames(x, xs), K.conv_fwd(xs, 1, t1, b1, h1), K.conv_fwd(h1, 2, t2, b2, h2)))
  recs.append(("encoder fused total, algorithmic = frame in + h2 out", t, S * (21168 * esz + 5184)))
  if dt == torch.float32:
    cols1 = torch.empty(S * 400, 192, dtype=torch.bfloat16, device=dev)
    cols2 = torch.empty(S * 81, 256, dtype=torch.bfloat16, device=dev)
    t = bench(lambda: (K.im2col(x, 8, 8, 4, out=cols1),
                       K.gemm_bf16(cols1, w1.view(192, 16), out=h1.view(S * 400, 16), b_mn_major=True, bias=b1, relu=True),
                       K.im2col(h1, 4, 4, 2, out=cols2),
                       K.gemm_bf16(cols2, w2.view(256, 32), out=h2.view(S * 81, 32), b_mn_major=True, bias=b2, relu=True)))
    recs.append(("encoder via explicit im2col + GEMM", t, S * (21168 * esz + 5184)))
    del cols1, cols2
  for name, t, byts in recs:
    out.write(json.dumps(dict(kernel=name, frames=S, dtype=str(dt), us=t * 1e6, gbs=byts / t / 1e9,
                              frac_of_measured_hbm=byts / t / 1e9 / 6535.7, ns_per_frame=t / S * 1e9)) + "\n")
  out.flush()
