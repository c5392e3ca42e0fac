"""Checkpoint of a whole batched agent.

The reference checkpoints only the TF variables through tf.train.Saver (main.py:356, :469-519):
the RMSProp slots travel with them, but the replay memory, the RandomState and the env / LSTM states
are lost on restart and the buffer is re-filled.  Here everything that determines the next update is
saved: parameters, `rms` / `momentum` slots, the 8-byte-per-frame rings (plus the frame / map / reward
payload rings of a framed agent -- 23 KB per frame, so that is a large file at full size), the per-env
MT19937 streams, maze positions (or current frames and producer state) with last action / reward, LSTM
state and the step counters -- so a restored agent
continues with the very same actions and replay samples.  Format: one `torch.save` dict of CPU tensors.
Variables are also stored by their reference names (`W_base_conv1` ...) in TF layouts, which is what
a converter from / to a TF-1 checkpoint (`net_-1/...` variable names) needs.
"""
import torch

from .. import _lib

FORMAT = "unreal_b200.checkpoint.v1"


def _cpu(t):
  return t.detach().to("cpu").clone()


def state_dict(trainer, global_t=0):
  net, ap, env = trainer.local_network, trainer.grad_applier, trainer.environment
  torch.cuda.synchronize(trainer.device)
  d = {"format": FORMAT, "global_t": int(global_t), "local_t": int(trainer.local_t), "num_envs": trainer.num_envs,
       "history": trainer.experience_history_size,
       "flat": _cpu(net.flat), "variables": {k: _cpu(v) for k, v in net.named_vars().items()},
       "lstm": tuple(_cpu(s) for s in net.base_lstm_state_out),
       "ring": {k: _cpu(v) for k, v in trainer.experience.ring.export_state().items()},
       "ring_full": bool(trainer._ring_full),
       "rng": {"mt": _cpu(trainer.streams.mt), "pos": _cpu(trainer.streams.pos)},
       "episode_reward": _cpu(trainer.episode_reward)}
  if hasattr(env, "export_state"):       # generic-frame env (frame_environment.py): frames are state, not a function of it
    d["env"] = {k: (_cpu(v) if isinstance(v, torch.Tensor) else v) for k, v in env.export_state().items()}
    d["payload"] = {k: _cpu(v) for k, v in trainer.experience.payload_state().items()}
  else:
    d["env"] = {"pos": _cpu(env.state.pos), "last_action": _cpu(env.state.last_action),
                "last_reward": _cpu(env.state.last_reward)}
  if ap is not None and getattr(ap, "_vars", None) is not None:
    d["rmsprop"] = {"rms": _cpu(ap._rms), "momentum": _cpu(ap._mom), "shard_lo": int(ap._lo), "shard": int(ap._shard)}
  return d


def save(path, trainer, global_t=0):
  torch.save(state_dict(trainer, global_t), path)


def load_state_dict(trainer, d):
  if d.get("format") != FORMAT:
    raise _lib.UnrealError("not an unreal_b200 checkpoint (format %r)" % (d.get("format"),))
  if d["num_envs"] != trainer.num_envs or d["history"] != trainer.experience_history_size:
    raise _lib.UnrealError("checkpoint is for %d envs / history %d" % (d["num_envs"], d["history"]))
  net, ap, env, dev = trainer.local_network, trainer.grad_applier, trainer.environment, trainer.device
  with torch.no_grad():
    net.flat.copy_(d["flat"].to(dev))
  net.refresh_shadow()
  net.base_lstm_state_out = tuple(s.to(dev) for s in d["lstm"])      # copied into the persistent state buffers
  trainer.experience.ring.import_state(d["ring"])
  trainer._ring_full = bool(d["ring_full"])
  trainer.streams.mt.copy_(d["rng"]["mt"].to(dev)); trainer.streams.pos.copy_(d["rng"]["pos"].to(dev))
  if hasattr(env, "import_state"):
    if "payload" not in d:
      raise _lib.UnrealError("checkpoint of a compact-record (maze) agent cannot restore a framed agent")
    env.import_state(d["env"])
    trainer.experience.load_payload_state(d["payload"])
  else:
    env.state.pos.copy_(d["env"]["pos"].to(dev)); env.state.last_action.copy_(d["env"]["last_action"].to(dev))
    env.state.last_reward.copy_(d["env"]["last_reward"].to(dev))
    from .. import kernels as K
    K.maze_render(env.state.pos, env._obs)            # frames are a function of the positions
    env.last_state = {'image': env._obs}
  trainer.episode_reward.copy_(d["episode_reward"].to(dev))
  trainer.local_t = int(d["local_t"])
  if "rmsprop" in d and ap is not None:
    ap.bind_flat(net.flat)
    if int(d["rmsprop"]["shard"]) != int(ap._shard) or int(d["rmsprop"]["shard_lo"]) != int(ap._lo):
      raise _lib.UnrealError("RMSProp slots were saved for a different world size / rank")
    ap._rms.copy_(d["rmsprop"]["rms"].to(dev)); ap._mom.copy_(d["rmsprop"]["momentum"].to(dev))
  torch.cuda.synchronize(dev)
  return int(d["global_t"])


def load(path, trainer):
  return load_state_dict(trainer, torch.load(path, map_location="cpu", weights_only=False))
