"""Small invocations of the round-2 kernels for `compute-sanitizer --tool memcheck` (odd sizes, tails, optional outputs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unreal_b200 import kernels as K

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
# K1 window kernels, env counts off the CTA / warp granularity
for dt in (torch.float32, torch.uint8):
  for n, t in ((5, 7), (67, 20), (1, 32)):
    st = K.MazeState(n, dev)
    acts = torch.randint(0, 4, (t, n), device=dev, dtype=torch.int32, generator=g)
    obs = torch.empty(t, n, 84, 84, 3, dtype=dt, device=dev)
    pc = torch.empty(t, n, 20, 20, device=dev)
    rec = torch.empty(t, n, dtype=torch.int64, device=dev)
    K.maze_window(st, acts, obs=obs, pc=pc, frame_rec=rec, auto_reset=True)
    K.maze_window(st, acts, auto_reset=False)
# K2 u8 (fewer sequences than SMs, several frames) + self-check + subsample
f8 = torch.randint(0, 256, (3, 5, 84, 84, 3), dtype=torch.uint8, device=dev, generator=g)
K.pixel_change_stream(f8)
K.pixel_change(f8[:, 1].contiguous(), f8[:, 0].contiguous())
K.subsample(torch.rand(2, 80, 80, device=dev, generator=g), 4)
K.s2d_frames(f8[:, 0].contiguous())
# cell tables
pos = torch.stack((torch.randint(0, 7, (1001,), device=dev, generator=g), torch.randint(0, 7, (1001,), device=dev, generator=g)), 1).int().contiguous()
tab = torch.randn(49, 256, device=dev, generator=g).to(torch.bfloat16)
K.cell_gather(tab, pos)
wide = torch.zeros(1001, 520, dtype=torch.bfloat16, device=dev)
K.cell_gather(tab, pos, out=wide[:, :256])
K.cell_segment_sum(torch.randn(1001, 256, device=dev, generator=g).to(torch.bfloat16), pos)
K.cell_segment_sum(torch.randn(1001, 256, device=dev, generator=g), pos)
# RP loss, rows_select
lg = torch.randn(77, 8, device=dev, generator=g)
c = torch.nn.functional.one_hot(torch.randint(0, 3, (77,), device=dev, generator=g), 3).float()
K.rp_loss(lg, torch.zeros(3, device=dev), c, want_p=True, want_loss=True, want_grad=True)
out = torch.zeros(9, 84, 84, 3, dtype=torch.uint8, device=dev)
K.rows_select(out, f8[:, 0].contiguous().repeat(3, 1, 1, 1), None, (torch.arange(9, device=dev) % 2).to(torch.uint8))
K.rows_select(out, f8.view(15, 84, 84, 3), torch.randint(0, 15, (9,), device=dev, generator=g), None)
# replay sample_rp on a small ring
ring = K.ReplayRing(5, 40, dev)
streams = K.MtStreams(np.arange(5) + 1, dev)
st = K.MazeState(5, dev)
rec = torch.empty(5, dtype=torch.int64, device=dev)
for i in range(60):
  K.maze_step(st, torch.randint(0, 4, (5,), device=dev, dtype=torch.int32, generator=g), frame_rec=rec, auto_reset=True)
  ring.add(rec)
ring.sample_rp(streams); ring.sample_sequence(streams, 21)
# LSTM cells with bf16 gates, fused PC head
n = 37
gates = torch.randn(n, 1024, device=dev, generator=g).to(torch.bfloat16)
cprev = torch.randn(n, 256, device=dev, generator=g)
cout = torch.empty(n, 256, device=dev); hout = torch.empty(n, 256, device=dev)
xh = torch.zeros(n, 520, dtype=torch.bfloat16, device=dev)
K.lstm_cell_fwd(gates, cprev, cout, hout, xh[:, 264:])
K.lstm_cell_act(gates, cout, hout, torch.empty(n, 256, device=dev), (torch.arange(n, device=dev) % 3 != 0).to(torch.uint8))
dg = torch.empty(n, 1024, dtype=torch.bfloat16, device=dev)
K.lstm_cell_bwd(gates, cprev, cout, hout, torch.zeros(n, 256, device=dev), dg, None)
from unreal_b200.model.model import UnrealModel
m = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0, num_envs=3, seed=1)
hp = torch.relu(torch.randn(7, 2592, device=dev, generator=g)).to(torch.bfloat16)
K.pc_deconv_loss(hp, m.pc_taps, m.pc_b8, torch.randint(0, 4, (7,), device=dev, dtype=torch.int32, generator=g),
                 torch.rand(7, 400, device=dev, generator=g), torch.ones(7, device=dev), 4, 0.05)
K.pc_deconv_qmax(hp, m.pc_taps, m.pc_b8, 4)
torch.cuda.synchronize()
print("sanitize targets done")
