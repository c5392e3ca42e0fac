"""Small fixed workload for ncu (session-3 kernels, 4096 frames / samples, one launch each after a warm-up pass):
conv2 forward on 128-byte overlapping-row boxes, conv2 dgrad fused with conv1's ReLU gradient, conv2 wgrad (M = 64),
conv1 forward / wgrad reading x'' and in render-fused (maze cell) mode, the pixel-control deconv forward, and the
framed-ring gather."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import kernels as K
dev = torch.device("cuda", 0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device=dev).manual_seed(0)
w1 = (torch.randn(8, 8, 3, 16, device=dev, generator=g) * 0.05).to(torch.bfloat16); b1 = torch.zeros(16, device=dev)
w2 = (torch.randn(4, 4, 16, 32, device=dev, generator=g) * 0.05).to(torch.bfloat16); b2 = torch.zeros(32, device=dev)
w8 = (torch.randn(4, 4, 8, 32, device=dev, generator=g) * 0.05).to(torch.bfloat16); b8 = torch.zeros(8, device=dev)
t1, t2, d2, t8 = K.conv1_w_planes(w1), K.conv_taps(w2, 2), K.conv2_dgrad_taps(w2), K.pc_deconv_taps(w8)
pos = torch.stack((torch.randint(0, 2, (S,), device=dev, generator=g), torch.randint(0, 7, (S,), device=dev, generator=g)), 1).to(torch.int32)
xpp = K.maze_render(pos, dtype=torch.bfloat16)
dy2 = (torch.randn(S * 81, 32, device=dev, generator=g) * 0.1).to(torch.bfloat16)
ring = K.ReplayRing(256, 64, dev)
payload = torch.randint(0, 256, (256, 64, 84, 84, 3), dtype=torch.uint8, device=dev, generator=g)
start = torch.randint(0, 40, (256,), dtype=torch.int32, device=dev, generator=g)
length = torch.full((256,), 21, dtype=torch.int32, device=dev)
act_pc = torch.randint(0, 4, (S,), device=dev, dtype=torch.int32, generator=g)
tgt_pc = torch.rand(S, 400, device=dev, generator=g)
for _ in range(2):
  h1 = K.conv_fwd(xpp, 1, t1, b1)
  h1m = K.conv1_fwd_maze(pos, t1, b1)
  h2 = K.conv_fwd(h1, 2, t2, b2)
  planes, db1 = K.conv2_dgrad_relu(dy2, d2, h1, pitch21=True)
  dw2 = K.conv2_wgrad(h1, dy2)
  dw1 = K.conv1_wgrad(xpp, planes)
  dw1m = K.conv1_wgrad_maze(pos, planes)
  y8 = K.pc_deconv_fwd(h2, t8, b8)
  pl = K.pc_deconv_loss(h2, t8, b8, act_pc, tgt_pc, torch.ones(S, device=dev), 4, 0.05)
  qm = K.pc_deconv_qmax(h2, t8, b8, 4)
  fr = ring.gather(payload, start, length, 21)
torch.cuda.synchronize()
assert torch.equal(h1, h1m)
print("ok")
