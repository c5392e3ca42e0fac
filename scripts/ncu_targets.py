"""Small single-kernel workloads for `ncu --set full` captures (one GPU, a handful of launches).

    python scripts/ncu_targets.py k2u8 | k2f32 | k1u8w | k1f32w | k1u8 | k4 | rp | conv2bwd | gemm2sm

Each target runs its kernel three times on the benchmark's shapes; select the kernel with `-k regex:...`.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unreal_b200 import kernels as K

dev = torch.device("cuda", 0)
which = sys.argv[1]
g = torch.Generator(device=dev).manual_seed(0)
if which in ("k2u8", "k2f32"):
  S, L = 6000, 20
  if which == "k2u8":
    frames = torch.randint(0, 256, (S, L + 1, 84, 84, 3), dtype=torch.uint8, device=dev, generator=g)
  else:
    S = 1500
    frames = torch.rand(S, L + 1, 84, 84, 3, device=dev, generator=g)
  out = torch.empty(S, L, 20, 20, device=dev)
  for _ in range(3):
    K.pixel_change_stream(frames, out)
elif which in ("k1u8w", "k1f32w", "k1u8", "k1f32"):
  from unreal_b200.train.rollout import RolloutTargets
  dt = torch.uint8 if "u8" in which else torch.float32
  eng = RolloutTargets(4096, 20, 0.99, 0.9, dt, dev, auto_reset=True, use_graphs=False, window_kernel=which.endswith("w"))
  eng.actions.copy_(torch.randint(0, 4, (20, 4096), device=dev, dtype=torch.int32, generator=g))
  for _ in range(3):
    eng.run_device()
elif which == "k4":
  pc = torch.rand(20, 4096, 20, 20, device=dev, generator=g)
  boot = torch.rand(4096, 20, 20, device=dev, generator=g)
  term = torch.zeros(20, 4096, dtype=torch.uint8, device=dev)
  out = torch.empty_like(pc)
  for _ in range(3):
    K.pc_targets(pc, term, None, boot, 0.9, out)
elif which == "gemm2sm":
  for (m, n, k) in ((81920, 256, 2592), (8192, 8192, 8192)):
    a = torch.randn(m, k, device=dev, generator=g).to(torch.bfloat16)
    b = torch.randn(k, n, device=dev, generator=g).to(torch.bfloat16)
    for _ in range(3):
      K.gemm_bf16(a, b, b_mn_major=True, out_dtype=torch.bfloat16)
  x = torch.randn(81920, 2592, device=dev, generator=g).to(torch.bfloat16)
  dy = torch.randn(81920, 256, device=dev, generator=g).to(torch.bfloat16)
  for _ in range(3):
    K.gemm_bf16(x, dy, a_mn_major=True, b_mn_major=True, split_k=16)
else:
  raise SystemExit("unknown target " + which)
torch.cuda.synchronize()
print("done", which)
