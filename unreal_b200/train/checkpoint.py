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


def _world(applier):
  """(world_size, rank) of the learner group: the applier's when it has one, else torch.distributed's."""
  if applier is not None and hasattr(applier, "_world"):
    return applier._world()
  import torch.distributed as dist
  if dist.is_available() and dist.is_initialized():
    return dist.get_world_size(), dist.get_rank()
  return 1, 0


def rank_path(path, world, rank):
  """A sharded agent is `world` files: each rank owns its envs, rings, RNG streams, LSTM state and its 1/world slice of
  the RMSProp slots, so each rank writes (and later reads) its own file.  world == 1: `path` itself."""
  return path if world == 1 else "%s.rank%dof%d" % (path, rank, world)


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
  world, rank = _world(ap)
  d["world"], d["rank"] = world, rank
  if ap is not None and getattr(ap, "_vars", None) is not None:
    d["rmsprop"] = {"rms": _cpu(ap._rms), "momentum": _cpu(ap._mom), "shard_lo": int(ap._lo), "shard": int(ap._shard)}
  return d


def save(path, trainer, global_t=0):
  """Write the checkpoint; returns the file actually written (`rank_path`: one file per rank when sharded)."""
  world, rank = _world(trainer.grad_applier)
  out = rank_path(path, world, rank)
  torch.save(state_dict(trainer, global_t), out)
  return out


_TENSOR_DTYPES = (torch.float32, torch.float64, torch.int32, torch.int64, torch.uint8, torch.bfloat16)
_KEYS = {"format", "global_t", "local_t", "num_envs", "history", "flat", "variables", "lstm", "ring", "ring_full", "rng",
         "episode_reward", "env", "payload", "rmsprop", "world", "rank"}


def _check_tree(x, where):
  """Only tensors of the dtypes state_dict() writes, plain containers and scalars."""
  if isinstance(x, torch.Tensor):
    if x.dtype not in _TENSOR_DTYPES:
      raise _lib.UnrealError("checkpoint entry %s has unexpected dtype %s" % (where, x.dtype))
  elif isinstance(x, dict):
    for k, v in x.items():
      if not isinstance(k, str):
        raise _lib.UnrealError("checkpoint entry %s has a non-string key" % where)
      _check_tree(v, "%s.%s" % (where, k))
  elif isinstance(x, (tuple, list)):
    for i, v in enumerate(x):
      _check_tree(v, "%s[%d]" % (where, i))
  elif not isinstance(x, (int, float, bool, str, type(None))):
    raise _lib.UnrealError("checkpoint entry %s has unexpected type %s" % (where, type(x).__name__))


def _want(cond, msg):
  if not cond:
    raise _lib.UnrealError("checkpoint rejected: " + msg)


def validate(trainer, d):
  """Everything load_state_dict() relies on, checked BEFORE the trainer is touched: a rejected checkpoint leaves
  the agent exactly as it was."""
  _want(isinstance(d, dict) and d.get("format") == FORMAT, "not an unreal_b200 checkpoint (format %r)" % (
      d.get("format") if isinstance(d, dict) else None,))
  unknown = set(d) - _KEYS
  _want(not unknown, "unknown keys %s" % sorted(unknown))
  _check_tree(d, "checkpoint")
  net, ap, env = trainer.local_network, trainer.grad_applier, trainer.environment
  _want(d["num_envs"] == trainer.num_envs and d["history"] == trainer.experience_history_size,
        "it is for %d envs / history %d, the trainer has %d / %d" % (d["num_envs"], d["history"], trainer.num_envs,
                                                                     trainer.experience_history_size))
  world, rank = _world(ap)
  _want((d.get("world", 1), d.get("rank", 0)) == (world, rank),
        "written by rank %s of %s, this is rank %d of %d" % (d.get("rank", 0), d.get("world", 1), rank, world))
  _want(tuple(d["flat"].shape) == tuple(net.flat.shape) and d["flat"].dtype == torch.float32,
        "parameter buffer %s does not match the network's %s" % (tuple(d["flat"].shape), tuple(net.flat.shape)))
  _want(len(d["lstm"]) == 2 and all(tuple(s.shape) == (trainer.num_envs, 256) for s in d["lstm"]), "LSTM state shape")
  ring = d["ring"]
  _want(set(ring) == {"rec", "top", "count", "n_pos", "n_neg"} and
        tuple(ring["rec"].shape) == (trainer.num_envs, trainer.experience_history_size) and
        all(tuple(ring[k].shape) == (trainer.num_envs,) for k in ("top", "count", "n_pos", "n_neg")), "replay ring shape")
  _want(tuple(d["rng"]["mt"].shape) == tuple(trainer.streams.mt.shape) and
        tuple(d["rng"]["pos"].shape) == tuple(trainer.streams.pos.shape), "RNG stream shape")
  _want(tuple(d["episode_reward"].shape) == (trainer.num_envs,), "episode_reward shape")
  framed = hasattr(env, "import_state")
  _want(framed == ("payload" in d), "a %s checkpoint cannot restore a %s agent" % (
      ("framed", "compact-record (maze)") if "payload" in d else ("compact-record (maze)", "framed")))
  if framed:
    for k, v in trainer.experience.payload_state().items():
      _want(k in d["payload"] and tuple(d["payload"][k].shape) == tuple(v.shape), "payload ring %s shape" % k)
  else:
    e = d["env"]
    _want(tuple(e["pos"].shape) == (trainer.num_envs, 2) and tuple(e["last_action"].shape) == (trainer.num_envs,) and
          tuple(e["last_reward"].shape) == (trainer.num_envs,), "maze state shape")
  if "rmsprop" in d and ap is not None:
    ap.bind_flat(net.flat)                     # allocates the (still untouched-by-the-checkpoint) slots
    r = d["rmsprop"]
    _want(int(r["shard"]) == int(ap._shard) and int(r["shard_lo"]) == int(ap._lo),
          "RMSProp slots were saved for a different world size / rank")
    _want(tuple(r["rms"].shape) == tuple(ap._rms.shape) and tuple(r["momentum"].shape) == tuple(ap._mom.shape),
          "RMSProp slot shape")


def load_state_dict(trainer, d):
  validate(trainer, d)
  net, ap, env, dev = trainer.local_network, trainer.grad_applier, trainer.environment, trainer.device
  with torch.no_grad():
    net.flat.copy_(d["flat"].to(dev))
  net.refresh_shadow()
  net.base_lstm_state_out = tuple(s.to(dev) for s in d["lstm"])      # copied into the persistent state buffers
  trainer.experience.ring.import_state(d["ring"])
  trainer._ring_full = bool(d["ring_full"])
  trainer._fill_active = (1 - trainer.experience.ring.state()["full"]).to(torch.uint8)   # warm-up mask follows the rings
  trainer.streams.mt.copy_(d["rng"]["mt"].to(dev)); trainer.streams.pos.copy_(d["rng"]["pos"].to(dev))
  if hasattr(env, "import_state"):
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
    ap._rms.copy_(d["rmsprop"]["rms"].to(dev)); ap._mom.copy_(d["rmsprop"]["momentum"].to(dev))
  torch.cuda.synchronize(dev)
  return int(d["global_t"])


def load(path, trainer):
  """Read this rank's file.  `weights_only=True`: the file is tensors, plain containers and scalars, so nothing in it
  is ever unpickled into code; its keys, dtypes and shapes are validated before the trainer is modified."""
  world, rank = _world(trainer.grad_applier)
  return load_state_dict(trainer, torch.load(rank_path(path, world, rank), map_location="cpu", weights_only=True))
