"""BASELINE configs[3]: pixel-control target stress -- synthetic 84x84x3 u8 frame stream, 20x20
cells, n = 20, gamma_pc = 0.9, 1 M frames per batch: K2 (stream pixel change, every frame read once)
+ K4 (PC Q-target scan).  Frames/s and GB/s against the algorithmic bytes of SURVEY 8(d)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unreal_b200 import kernels as K

dev = torch.device("cuda", 0)
dtype = torch.uint8 if (len(sys.argv) < 2 or sys.argv[1] == "u8") else torch.float32
SEQ = int(sys.argv[2]) if len(sys.argv) > 2 else 50000        # 50 000 sequences x 20 (+1) frames = 1 M frames
out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
L = 20
g = torch.Generator(device=dev).manual_seed(0)
frames = torch.empty(SEQ, L + 1, 84, 84, 3, dtype=dtype, device=dev)
for i in range(0, SEQ, 5000):     # fill in chunks (randint of the whole 22 GB tensor at once needs 8x the memory)
  chunk = frames[i:i + 5000]
  if dtype == torch.uint8:
    chunk.copy_(torch.randint(0, 256, chunk.shape, dtype=torch.uint8, device=dev, generator=g))
  else:
    chunk.copy_(torch.rand(chunk.shape, device=dev, generator=g))
boot = torch.rand(SEQ, 20, 20, device=dev, generator=g)
pc = torch.empty(SEQ, L, 20, 20, device=dev)
pc_t = torch.empty(L, SEQ, 20, 20, device=dev)      # K4 wants time-major
tgt = torch.empty(L, SEQ, 20, 20, device=dev)


def bench(fn, iters=5):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


n_frames = SEQ * L
esz = frames.element_size()
t2 = bench(lambda: K.pixel_change_stream(frames, pc))
b2 = SEQ * (L + 1) * 21168 * esz + n_frames * 1600
pc_t.copy_(pc.transpose(0, 1))
t4 = bench(lambda: K.pc_targets(pc_t, None, None, boot, 0.9, tgt))
b4 = n_frames * 3200 + SEQ * 1600
for name, t, b in (("K2 pixel_change_stream", t2, b2), ("K4 pc_targets", t4, b4)):
  out.write(json.dumps(dict(kernel=name, dtype=str(dtype), sequences=SEQ, frames=n_frames, ms=t * 1e3,
                            frames_per_s=n_frames / t, gbs=b / t / 1e9, frac_of_measured_hbm=b / t / 1e9 / 6535.7)) + "\n")
out.write(json.dumps(dict(kernel="K2+K4 (configs[3] pass)", dtype=str(dtype), frames=n_frames, ms=(t2 + t4) * 1e3,
                          frames_per_s=n_frames / (t2 + t4), gbs=(b2 + b4) / (t2 + t4) / 1e9,
                          frac_of_measured_hbm=(b2 + b4) / (t2 + t4) / 1e9 / 6535.7)) + "\n")
