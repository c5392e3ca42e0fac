"""Per-kernel timings with CUDA events (not the headline bench; used to A/B variants).
Writes JSON lines to stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from unreal_b200 import _lib, kernels as K

PEAK = 6535.7  # GB/s, MEASURED_PEAKS.json hbm_gbs


def timeit(fn, iters=20, warm=3):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
  ev[0].record()
  for i in range(iters):
    fn()
    ev[i + 1].record()
  torch.cuda.synchronize()
  ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
  return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3


def timeit_graph(fn, reps=20, iters=10):
  """Launch-overhead-free timing: `reps` calls captured in one CUDA graph."""
  fn()
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    for _ in range(reps):
      fn()
  tm, tn = timeit(g.replay, iters=iters, warm=2)
  return tm / reps, tn / reps


def report(name, bytes_, t_med, t_min, **kw):
  print(json.dumps(dict(kernel=name, us_median=round(t_med * 1e6, 2), us_min=round(t_min * 1e6, 2),
                        gbs_median=round(bytes_ / t_med / 1e9, 1), frac_of_measured_peak=round(bytes_ / t_med / 1e9 / PEAK, 3),
                        **kw)), flush=True)


def main():
  _lib.require_device()
  dev = "cuda:0"
  N, T = 4096, 20
  st = K.MazeState(N, dev)
  g = torch.Generator(device=dev).manual_seed(0)
  acts = torch.randint(0, 4, (T, N), device=dev, dtype=torch.int32, generator=g)
  pc = torch.empty(T, N, 20, 20, device=dev)
  rew = torch.empty(T, N, device=dev)
  term = torch.empty(T, N, dtype=torch.uint8, device=dev)
  for dt, name, fb in ((torch.float32, "f32", 84672), (torch.uint8, "u8", 21168)):
    obs = torch.empty(T, N, 84, 84, 3, dtype=dt, device=dev)
    variants = ((0, 256), (0, 512), (1, 0), (2, 0))
    for v, th in variants:
      _lib.set_tunable("maze_render_variant", v)
      if th:
        _lib.set_tunable("maze_cta_threads", th)

      def rollout():
        for t in range(T):
          K.maze_step(st, acts[t], obs=obs[t], pc=pc[t], reward=rew[t], terminal=term[t], auto_reset=True)
      tm, tn = timeit_graph(rollout, reps=1, iters=10)
      report("maze_step(graph of %d)" % T, N * (fb + 1600 + 41), tm / T, tn / T, obs=name, variant=v, threads=th, n=N)
    del obs
  _lib.set_tunable("maze_render_variant", -1)
  v = torch.randn(T, N, device=dev)
  boot = torch.randn(N, device=dev)
  R = torch.empty(T, N, device=dev); adv = torch.empty(T, N, device=dev)
  tm, tn = timeit_graph(lambda: K.nstep_returns(rew, v, term, boot, 0.99, R, adv))
  report("nstep_returns(graph)", N * 344, tm, tn, n=N, t=T)
  qb = torch.rand(N, 20, 20, device=dev)
  tgt = torch.empty_like(pc)
  tm, tn = timeit_graph(lambda: K.pc_targets(pc, term, None, qb, 0.9, tgt), reps=4)
  report("pc_targets(graph)", N * 65600, tm, tn, n=N, t=T)
  # a plain device copy of the same size as one f32 K1 launch, for calibration
  a = torch.empty(N * 84672 // 4, device=dev); b = torch.empty_like(a)
  tm, tn = timeit(lambda: b.copy_(a))
  report("torch_copy_calibration", 2 * a.numel() * 4, tm, tn)
  tm, tn = timeit(lambda: b.fill_(1.0))
  report("torch_fill_calibration", a.numel() * 4, tm, tn)


if __name__ == "__main__":
  sys.exit(main())
