"""Kernel-time breakdown of one learner update + rollout (torch.profiler, CUDA activities)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/agent_profile.txt"
  mode = sys.argv[3] if len(sys.argv) > 3 else "cells"
  from unreal_b200.environment.environment import Environment
  from unreal_b200.model.model import UnrealModel
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  from unreal_b200.train.trainer import Trainer
  dev = torch.device("cuda", 0)
  net = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0,
                    0.0, num_envs=n, seed=0)
  applier = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr = Trainer(0, net, 7e-4, None, applier, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, 100,
               10 ** 8, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
               num_envs=n, seeds=np.arange(n) + 11, obs_s2d=(mode == "s2d"), obs_cells=(mode == "cells"), use_graphs=False)
  tr.prepare()
  while not tr.experience.is_full():
    tr.process(None, 0)
  for _ in range(2):
    tr.process(None, 0)
  torch.cuda.synchronize()
  if len(sys.argv) > 4 and sys.argv[4] == "phases":
    # the data phase (rollout + sampling + targets) and the learner update profiled separately, kernels with launch counts
    tables = []
    for name, fn in (("data phase", lambda: tr._data_phase(None)), ("update", lambda: tr._update(tr.last_feed, 7e-4))):
      with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        tr._rng_in()
        try:
          feed = fn()
        finally:
          tr._rng_out()
        if name == "data phase":
          tr.last_feed = feed
        torch.cuda.synchronize()
      ka = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0]
      ka.sort(key=lambda e: -e.self_device_time_total)
      tot = sum(e.self_device_time_total for e in ka)
      lines = ["==== %s: %.1f us of device time, %d launches ====" % (name, tot, sum(e.count for e in ka))]
      for e in ka[:70]:
        lines.append("%9.1f us %5d x %7.2f  %-70s %s" % (e.self_device_time_total, e.count, e.self_device_time_total / e.count,
                                                        e.key[:70], str(e.input_shapes)[:110]))
      tables.append("\n".join(lines))
    with open(out, "w") as f:
      f.write("\n\n".join(tables) + "\n")
    return
  with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(2):
      tr.process(None, 0)
    torch.cuda.synchronize()
  with open(out, "w") as f:
    f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
    f.write("\n\n==== by input shape ====\n")
    f.write(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=40,
                                                                max_name_column_width=50, max_shapes_column_width=90))


if __name__ == "__main__":
  main()
