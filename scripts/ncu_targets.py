"""Small single-kernel workloads for `ncu --set full` captures (one GPU, a handful of launches).

    python scripts/ncu_targets.py k2u8 | k2f32 | k1u8w | k1f32w | k1u8 | k4 | rp | conv2bwd | gemm2sm | lstm [envs] | lstmfused [envs] | pcfc1 | pcplanes | pctargets

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
elif which == "lstm":
  # one forward and one backward LSTM step of the learner at `envs` rows: step GEMM over [x, h], cell, recurrent dh GEMM, cell backward
  n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
  xh = torch.randn(n, 520, device=dev, generator=g).to(torch.bfloat16)
  w = (torch.randn(520, 1024, device=dev, generator=g) * 0.05).to(torch.bfloat16)
  b = torch.zeros(1024, device=dev)
  c0 = torch.randn(n, 256, device=dev, generator=g)
  dh = torch.randn(n, 256, device=dev, generator=g)
  for _ in range(2):
    gates = torch.empty(n, 1024, device=dev, dtype=torch.bfloat16)
    c1 = torch.empty(n, 256, device=dev); h = torch.empty(n, 256, device=dev)
    h16 = torch.empty(n, 256, device=dev, dtype=torch.bfloat16)
    K.gemm_bf16(xh, w, out=gates, b_mn_major=True, bias=b)
    K.lstm_cell_fwd(gates, c0, c1, h, h16)
    dc = torch.zeros(n, 256, device=dev)
    dg = torch.empty(n, 1024, device=dev, dtype=torch.bfloat16)
    K.lstm_cell_bwd(gates, c0, c1, dh, dc, dg, None)
    K.gemm_bf16(dg, w[264:])
elif which == "lstmfused":
  # the fused step kernels (csrc/lstm_tcgen05.cu) in the unroll's tiled mode: two forward and two backward steps
  n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
  xh = torch.randn(3, n, 520, device=dev, generator=g).to(torch.bfloat16)
  w = (torch.randn(520, 1024, device=dev, generator=g) * 0.05).to(torch.bfloat16)
  b = torch.zeros(1024, device=dev)
  c_all = torch.zeros(3, n, 256, device=dev)
  h_all = torch.empty(2, n, 256, device=dev)
  gates = torch.empty(2, n, 1024, device=dev, dtype=torch.bfloat16)
  dh = torch.randn(2, n, 256, device=dev, generator=g) * 0.1
  dc = torch.zeros(n, 256, device=dev)
  dg = torch.empty(2, n, 1024, device=dev, dtype=torch.bfloat16)
  for _ in range(2):
    for i in range(2):
      K.lstm_step_fwd(xh[i], w, b, c_all[i], c_all[i + 1], h_out=h_all[i], h16_out=xh[i + 1, :, 264:], acts=gates[i], tiled=True)
    K.lstm_step_bwd(None, w[264:], gates[1], c_all[1], c_all[2], dh[1], dc, dg[1], tiled=True)
    K.lstm_step_bwd(dg[1], w[264:], gates[0], c_all[0], c_all[1], dh[0], dc, dg[0], tiled=True)
elif which == "pcfc1":
  # pc_fc1 forward at the agent's batch: [163840,256] x [256,2592] + bias, ReLU, bf16 out (short K, 849 MB of output)
  S = 163840
  x = torch.randn(S, 256, device=dev, generator=g).to(torch.bfloat16)
  w = (torch.randn(256, 2592, device=dev, generator=g) * 0.05).to(torch.bfloat16)
  b = torch.zeros(2592, device=dev)
  out = torch.empty(S, 2592, device=dev, dtype=torch.bfloat16)
  for _ in range(3):
    K.gemm_bf16(x, w, out=out, b_mn_major=True, bias=b, relu=True)
elif which == "pcplanes":
  # the pixel-control tower behind pc_fc1 at half the agent's batch: fused deconv + loss (plane-major gradient out), the backward
  # convolution with pc_fc1's ReLU / bias gradient, the deconv filters' gradient -- one launch each
  from unreal_b200.model.model import UnrealModel
  S, A = 81920, 4
  m = UnrealModel(A, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0, num_envs=4, seed=0)
  hp = torch.relu(torch.randn(S, 2592, device=dev, generator=g)).to(torch.bfloat16)
  act = torch.randint(0, A, (S,), device=dev, generator=g, dtype=torch.int32)
  tgt = torch.rand(S, 400, device=dev, generator=g)
  msk = torch.ones(S, device=dev)
  sc = torch.tensor([0.5], device=dev)
  for _ in range(2):
    loss, dyp, db8 = K.pc_deconv_loss(hp, m.pc_taps, m.pc_b8, act, tgt, msk, A, 0.05, planes=True)
    K.pc_planes_conv(dyp, m.pc_w_planes, hp, scale=sc)
    K.pc_planes_wgrad(dyp, hp)
elif which == "pctargets":
  # Trainer._process_pc's one-pass target kernel at the agent's size (T = 20, N = 8192)
  T, N = 20, 8192
  p0 = torch.randint(0, 7, (T, N, 2), device=dev, generator=g, dtype=torch.int32)
  p1 = (p0 + torch.randint(-1, 2, (T, N, 2), device=dev, generator=g, dtype=torch.int32)).clamp_(0, 6)
  boot = torch.rand(N, 20, 20, device=dev, generator=g)
  ln = torch.full((N,), T, device=dev, dtype=torch.int32)
  out = torch.empty(T, N, 20, 20, device=dev)
  for _ in range(3):
    K.maze_pc_targets(p0, p1, ln, boot, 0.9, out=out)
else:
  raise SystemExit("unknown target " + which)
torch.cuda.synchronize()
print("done", which)
