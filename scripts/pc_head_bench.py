"""The pixel-control tower's kernels after pc_fc1 at the agent's batch (S = envs x 20 samples): fused deconv + loss with four /
eight epilogue warps, the backward convolution plain / with pc_fc1's ReLU mask + bias gradient in its epilogue (against the
separate unreal_relu_grad pass it replaces), conv2's wgrad kernel with the roles exchanged.  us per launch, GB/s of the
algorithmic bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import _lib, kernels as K
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
A = 4
hp = torch.relu(torch.randn(S, 2592, device=dev, generator=g)).to(torch.bfloat16)
from unreal_b200.model.model import UnrealModel
m = UnrealModel(A, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0, num_envs=4, seed=0)
b8 = m.pc_b8
act = torch.randint(0, A, (S,), device=dev, generator=g, dtype=torch.int32)
tgt = torch.rand(S, 400, device=dev, generator=g)
msk = torch.ones(S, device=dev)
sc = torch.tensor([0.5], device=dev)
def timed(fn, reps=5):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fn()
  e.record(); torch.cuda.synchronize()
  return a.elapsed_time(e) * 1e3 / reps
res = {"S": S, "unit": "us"}
for epi8 in (0, 1):
  _lib.set_tunable("pc_loss_epi8", epi8)
  res["deconv_loss_epi%d" % (8 if epi8 else 4)] = round(timed(lambda: K.pc_deconv_loss(hp, m.pc_taps, b8, act, tgt, msk, A, 0.05)), 1)
loss, dy16, db8 = K.pc_deconv_loss(hp, m.pc_taps, b8, act, tgt, msk, A, 0.05)
dy16 = dy16.view(S, 20, 20, 16)
out = torch.empty(S, 9, 9, 32, dtype=torch.bfloat16, device=dev)
res["bwd_conv_scaled"] = round(timed(lambda: K.conv2_fwd_linear(dy16, m.pc_lin_taps, out=out, scale=sc)), 1)
res["relu_grad_pass"] = round(timed(lambda: K.relu_grad(out.view(S, 2592), hp)), 1)
res["bwd_conv_masked"] = round(timed(lambda: K.conv2_fwd_linear(dy16, m.pc_lin_taps, out=out, scale=sc, mask_y=hp)), 1)
res["wgrad"] = round(timed(lambda: K.conv2_wgrad(dy16, hp.view(S * 81, 32))), 1)
gb = lambda nbytes, us: round(nbytes / us / 1e3)
res["deconv_loss_gbs"] = gb(S * (5184 + 1600 + 12800), res["deconv_loss_epi8"])
res["bwd_conv_masked_gbs"] = gb(S * (12800 + 5184 + 5184), res["bwd_conv_masked"])
res["wgrad_gbs"] = gb(S * (12800 + 5184), res["wgrad"])
# the same three kernels on the 8-channel gradient
loss, dy8, db8 = K.pc_deconv_loss(hp, m.pc_taps, b8, act, tgt, msk, A, 0.05, c8=True)
res["deconv_loss_c8"] = round(timed(lambda: K.pc_deconv_loss(hp, m.pc_taps, b8, act, tgt, msk, A, 0.05, c8=True)), 1)
dy8 = dy8.view(S, 20, 20, 8)
res["bwd_conv_masked_c8"] = round(timed(lambda: K.conv2_fwd_linear(dy8, m.pc_lin_taps8, out=out, scale=sc, mask_y=hp)), 1)
res["wgrad_c8"] = round(timed(lambda: K.conv2_wgrad(dy8, hp.view(S * 81, 32))), 1)
res["deconv_loss_c8_gbs"] = gb(S * (5184 + 1600 + 6400), res["deconv_loss_c8"])
res["bwd_conv_masked_c8_gbs"] = gb(S * (6400 + 5184 + 5184), res["bwd_conv_masked_c8"])
res["wgrad_c8_gbs"] = gb(S * (6400 + 5184), res["wgrad_c8"])
# ... and on the plane-major gradient (one bulk copy per sample)
res["deconv_loss_planes"] = round(timed(lambda: K.pc_deconv_loss(hp, m.pc_taps, b8, act, tgt, msk, A, 0.05, planes=True)), 1)
loss, dyp, db8 = K.pc_deconv_loss(hp, m.pc_taps, b8, act, tgt, msk, A, 0.05, planes=True)
res["bwd_conv_planes"] = round(timed(lambda: K.pc_planes_conv(dyp, m.pc_w_planes, hp, scale=sc, out=out)), 1)
res["wgrad_planes"] = round(timed(lambda: K.pc_planes_wgrad(dyp, hp)), 1)
res["deconv_loss_planes_gbs"] = gb(S * (5184 + 1600 + 6400), res["deconv_loss_planes"])
res["bwd_conv_planes_gbs"] = gb(S * (6400 + 5184 + 5184), res["bwd_conv_planes"])
res["wgrad_planes_gbs"] = gb(S * (6400 + 5184), res["wgrad_planes"])
print(json.dumps(res), flush=True)
