"""Learning check of the whole B200 agent: N maze envs, UNREAL (A3C-LSTM + PC + VR + RP), shared RMSProp.
Logs finished episodes and mean episode score (+1 goal, -1 per wall hit) per 100 updates; a random policy needs
~1000-2300 steps per episode (score about -330 from wall hits), the optimal path of this map is exactly 20 moves
(score +1).  One run per seed; the last line is the summary the judge asked for:

    {"learning_check": {"seeds": [...], "solved": k, "criterion": "..."}}

    python scripts/train_maze.py --envs 1024 --updates 2500 --lr 1e-2 --seeds 0,1,2,3 [--warmup 300] [--entropy-beta 0.001]
                                 [--obs cells|s2d|f32] [--out file.jsonl]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unreal_b200.environment.environment import Environment
from unreal_b200.model.model import UnrealModel
from unreal_b200.train.rmsprop_applier import RMSPropApplier
from unreal_b200.train.trainer import Trainer


def run(args, seed, out):
  n, dev = args.envs, torch.device("cuda", 0)
  Environment.action_size = -1
  net = UnrealModel(4, 0, -1, True, True, True, True, 0.05, args.entropy_beta, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0,
                    0.0, num_envs=n, seed=seed)
  if args.grad_sum:
    net.grad_scale = 1.0
  applier = RMSPropApplier(args.lr, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=args.clip)
  # the Trainer anneals lr0 * (max_t - global_t) / max_t (trainer.py:140-144): a warm-up is the same formula run
  # backwards, so it is expressed through the global_t handed to process()
  max_t = 10 ** 12
  tr = Trainer(0, net, args.lr, None, applier, 'maze', '', True, True, True, True, 0.05, args.entropy_beta, 20, 20, 0.99, 0.9,
               args.history, max_t, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
               num_envs=n, seeds=np.arange(n) + 1000 * seed + 0xA3C, use_graphs=True, obs_cells=(args.obs == 'cells'),
               obs_s2d=(args.obs == 's2d'))
  tr.prepare()
  while not tr.experience.is_full():
    tr.process(None, 0)
  torch.cuda.synchronize()
  t0 = last = time.perf_counter()
  env_steps, prev = 0, tr.episode_stats.clone()
  final = None
  for u in range(1, args.updates + 1):
    frac = min(1.0, (0.1 + 0.9 * u / args.warmup)) if args.warmup > 0 else 1.0      # lr = frac * lr0
    d, _ = tr.process(None, int((1.0 - frac) * max_t))
    env_steps += d
    if u % args.log_every == 0:
      cur = tr.episode_stats.clone()
      ep, sc = (cur - prev).tolist()
      prev = cur
      now = time.perf_counter()
      final = dict(seed=seed, updates=u, env_steps=env_steps, wall_s=now - t0, episodes=ep,
                   mean_score=(sc / ep if ep else None), steps_per_episode=(args.log_every * 20 * n / ep if ep else None),
                   env_steps_per_s=args.log_every * 20 * n / (now - last), lr=frac * args.lr,
                   total_loss_per_env=float(tr.last_losses["total"]) / n, grad_norm=float(tr.last_losses["grad_norm"]))
      last = now
      out.write(json.dumps(final) + "\n"); out.flush()
      if final["mean_score"] is not None and final["mean_score"] >= 0.95 and final["steps_per_episode"] <= 22 and u >= 300:
        break
  tr.stop()
  solved = bool(final and final["mean_score"] is not None and final["mean_score"] >= 0.9 and final["steps_per_episode"] <= 25)
  return solved, final


def main():
  p = argparse.ArgumentParser()
  p.add_argument("--envs", type=int, default=1024)
  p.add_argument("--updates", type=int, default=2500)
  p.add_argument("--lr", type=float, default=1e-2)
  p.add_argument("--warmup", type=int, default=0, help="updates over which lr ramps linearly from lr/10 to lr")
  p.add_argument("--entropy-beta", type=float, default=0.001)
  p.add_argument("--clip", type=float, default=40.0)
  p.add_argument("--grad-sum", action="store_true",
                 help="apply the SUM of the envs' gradients (N reference workers each applying its own, to first order) instead "
                      "of their mean; give --clip per update, e.g. 40 * envs")
  p.add_argument("--history", type=int, default=2000)
  p.add_argument("--log-every", type=int, default=100)
  p.add_argument("--seeds", default="0")
  p.add_argument("--obs", default="cells", choices=["cells", "s2d", "f32"])
  p.add_argument("--out", default="-")
  args = p.parse_args()
  out = sys.stdout if args.out == "-" else open(args.out, "w")
  seeds = [int(s) for s in args.seeds.split(",")]
  results = []
  for s in seeds:
    solved, final = run(args, s, out)
    results.append((s, solved, final))
    torch.cuda.empty_cache()
  out.write(json.dumps({"learning_check": {
      "seeds": seeds, "solved": sum(1 for _, ok, _ in results if ok),
      "criterion": "mean episode score >= 0.9 at <= 25 steps per episode (optimum: +1 in 20 moves) in the last logging window",
      "config": {"envs": args.envs, "max_updates": args.updates, "lr": args.lr, "warmup_updates": args.warmup,
                 "entropy_beta": args.entropy_beta, "clip": args.clip, "gradient": "sum over envs" if args.grad_sum else "mean over envs",
                 "obs": args.obs},
      "per_seed": [{"seed": s, "solved": ok, "updates": (f or {}).get("updates"), "mean_score": (f or {}).get("mean_score"),
                    "steps_per_episode": (f or {}).get("steps_per_episode")} for s, ok, f in results]}}) + "\n")
  out.flush()


if __name__ == "__main__":
  main()
