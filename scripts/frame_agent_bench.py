"""Framed path on one GPU: (1) payload-ring kernels against the HBM roofline -- ring
store of one step's frames and the gather of PC / VR sequences; (2) the full UNREAL agent on synthetic
indoor-shaped frames (uint8 84x84x3, 3 actions; BASELINE configs[4]'s observation shape) driven by
FrameTrainer: rollout through the frame env adapter (K2 pixel change per step), framed ring, PC / VR / RP
sampling with frame gathers, UnrealModel fwd/bwd, fused RMSProp.

  python scripts/frame_agent_bench.py [envs] [history] [updates] [graphs 0|1]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

PEAK = 6535.7
try:
  PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
  pass


def timed(fn, iters=20, warm=3):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  hist = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
  updates = int(sys.argv[3]) if len(sys.argv) > 3 else 6
  graphs = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
  from unreal_b200.environment.environment import Environment
  from unreal_b200.model.model import UnrealModel
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  from unreal_b200.train.trainer import Trainer
  dev = torch.device("cuda", 0)
  Environment.action_size = -1
  net = UnrealModel(3, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0,
                    0.0, num_envs=n, seed=0)
  applier = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr = Trainer(0, net, 7e-4, None, applier, 'synthetic', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9,
               hist, 10 ** 8, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
               num_envs=n, seeds=np.arange(n) + 11, env_args={'producer': 'table', 'seed': 0}, use_graphs=graphs)
  tr.prepare()
  ex = tr.experience
  t0 = time.time()
  # warm the ring with the env adapter alone (the policy forward of _fill_experience is not what is measured)
  env = tr.environment
  act = torch.zeros(n, dtype=torch.int32, device=dev)
  while not ex.is_full():
    for _ in range(50):
      prev = env.last_state['image']
      lr = env.last_reward.clone()
      act.random_(0, 3)
      _, r, _, pc = env.process(act)
      ex.add_frames(env.frame_rec, frame=prev, pixel_change=pc, reward=r, last_reward=lr)
  torch.cuda.synchronize()
  fill_s = time.time() - t0
  out = dict(workload="framed path: %d envs x %d-frame ring of uint8 84x84x3 frames (%.1f GB frames + %.1f GB maps)" %
             (n, hist, ex.frames.numel() / 1e9, ex.pc.numel() * 4 / 1e9), ring_fill_s=fill_s, peak_gbs=PEAK)

  # (1) payload kernels
  L = 21
  fb = 84 * 84 * 3
  slot = ex.ring.add_slots(torch.zeros(n, dtype=torch.int64, device=dev))      # invalid records: slots -1
  slot.random_(0, hist)
  src = env.last_state['image']
  s = timed(lambda: ex.ring.store(ex.frames, src, slot))
  out["ring_store_frames"] = dict(us=s * 1e6, gbs=2 * n * fb / s / 1e9, frac=2 * n * fb / s / 1e9 / PEAK)
  start, length, _ = ex.sample_sequence(L)
  dst = torch.empty(L, n, 84, 84, 3, dtype=torch.uint8, device=dev)
  s = timed(lambda: ex.ring.gather(ex.frames, start, length, L, True, out=dst))
  moved = float(length.sum()) * fb * 2 + float((L - length).sum()) * fb
  out["gather_frames_seq21"] = dict(us=s * 1e6, gbs=moved / s / 1e9, frac=moved / s / 1e9 / PEAK,
                                    mean_len=float(length.float().mean()))
  dpc = torch.empty(L, n, 20, 20, dtype=torch.float32, device=dev)
  s = timed(lambda: ex.ring.gather(ex.pc, start, length, L, True, out=dpc))
  moved = float(length.sum()) * 1600 * 2 + float((L - length).sum()) * 1600
  out["gather_pc_seq21"] = dict(us=s * 1e6, gbs=moved / s / 1e9, frac=moved / s / 1e9 / PEAK)

  # (2) the full agent
  tr._ring_full = True
  for _ in range(2):
    tr.process(None, 0)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  steps = 0
  e0.record()
  for _ in range(updates):
    d, _ = tr.process(None, 0)
    steps += n * 20
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / updates
  out["agent"] = dict(cuda_graph_data_phase=graphs, ms_per_update=ms, env_steps_per_s=n * 20 / (ms * 1e-3), updates=updates,
                      finite=bool(torch.isfinite(tr.last_losses["total"])), grad_norm=float(tr.last_losses["grad_norm"]),
                      peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
  print(json.dumps(out))


if __name__ == "__main__":
  main()
