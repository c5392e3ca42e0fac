"""Small fixed workload for ncu: s2d, conv1 fwd, conv2 fwd, conv1 wgrad, K2 stream on 4096 frames."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import kernels as K
dev = torch.device("cuda", 0)
S = 4096
x = torch.rand(S, 84, 84, 3, device=dev)
w1 = (torch.randn(8, 8, 3, 16, device=dev) * 0.05).to(torch.bfloat16); b1 = torch.zeros(16, device=dev)
w2 = (torch.randn(4, 4, 16, 32, device=dev) * 0.05).to(torch.bfloat16); b2 = torch.zeros(32, device=dev)
t1, t2 = K.conv1_w_planes(w1), K.conv_taps(w2, 2)
dy = (torch.randn(S * 400, 16, device=dev) * 0.1).to(torch.bfloat16)
frames = torch.rand(200, 21, 84, 84, 3, device=dev)
for _ in range(3):
  xs = K.s2d_frames(x)
  h1 = K.conv_fwd(xs, 1, t1, b1)
  h2 = K.conv_fwd(h1, 2, t2, b2)
  dyp, db = K.relu_grad(dy, h1.view(S * 400, 16), planes=True)
  dw = K.conv1_wgrad(xs, dyp)
  pc = K.pixel_change_stream(frames)
torch.cuda.synchronize()
print("ok")
