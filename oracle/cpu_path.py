"""CPU timing leg of the benchmark: the oracle's restatement of the reference path
(maze step + render + pixel-change, then n-step returns and PC targets), run on host cores.

TEST/BENCH INFRASTRUCTURE (see oracle/__init__.py).  Used only by bench.py's `cpu_baseline`
leg and by `bench.py --impl reference`; it is the thing timed there, never a product path.
"""
import os
import time

import numpy as np

from . import unreal_oracle as O


def rollout_pass(env, rs, t_len, gamma=0.99, gamma_pc=0.9):
  """One env, one pass: T x process(action) (maze_environment.py:98-128, incl. render and
  _calc_pixel_change), reset on terminal, then the reverse scans of trainer.py:313-324 and
  :359-361 over what was collected.  Returns env-steps done."""
  rewards, values, pcs, terms = [], [], [], []
  for _ in range(t_len):
    a = int(rs.randint(4))
    _, r, term, pc = env.process(a)
    rewards.append(r); values.append(np.float32(0.1)); pcs.append(pc); terms.append(term)
    if term:
      env.reset()
  # returns: segmented like the batched kernel (a terminal cuts the bootstrap)
  R = np.float32(0.5)
  for i in range(t_len - 1, -1, -1):
    if terms[i]:
      R = np.float32(0.0)
    R = rewards[i] + gamma * R
    _adv = R - values[i]
  pc_R = np.full((20, 20), 0.5, np.float32)
  for i in range(t_len - 1, -1, -1):
    if terms[i]:
      pc_R = np.zeros((20, 20), np.float32)
    pc_R = pcs[i] + gamma_pc * pc_R
  return t_len


def _worker(args):
  seed, envs, t_len, passes = args
  rs = np.random.RandomState(seed)
  es = [O.MazeOracle() for _ in range(envs)]
  t0 = time.perf_counter()
  steps = 0
  for _ in range(passes):
    for e in es:
      steps += rollout_pass(e, rs, t_len)
  return steps, time.perf_counter() - t0


def run_parallel(procs, envs_per_proc, t_len, passes, pool=None):
  """-> (total env-steps, wall seconds) over `procs` worker processes."""
  import multiprocessing as mp
  args = [(1000 + i, envs_per_proc, t_len, passes) for i in range(procs)]
  t0 = time.perf_counter()
  if procs == 1:
    out = [_worker(args[0])]
  else:
    own = pool is None
    if own:
      pool = mp.get_context("fork").Pool(procs)
    try:
      out = pool.map(_worker, args)
    finally:
      if own:
        pool.close(); pool.join()
  wall = time.perf_counter() - t0
  return sum(s for s, _ in out), wall


def host_cores():
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count() or 1
