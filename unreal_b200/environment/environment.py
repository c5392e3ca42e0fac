"""Environment base class and factory: drop-in for environment/environment.py of the
reference (same static methods, same instance contract), backed by libunreal_b200.

Only the env types on the hot path are constructed here: 'maze' (environment/maze_environment.py)
and 'synthetic' (frames of the MINOS observation shape, environment/indoor_environment.py:63-139,
because MINOS scenes are unavailable offline).  'lab', 'gym' and 'indoor' need external
simulators and raise.  The log-dir statics of the reference (environment.py:18-27) are file
logging for MINOS and are out of scope.
"""
import numpy as np
import torch

from .. import _lib
from .. import kernels as K


class Environment(object):
  # cached action size (environment.py:13, :46-47)
  action_size = -1
  # env types added by the integrator: env_type -> (factory, action_size, objective_size)
  _registry = {}

  @staticmethod
  def register(env_type, factory, action_size, objective_size=0):
    """Make `create_environment(env_type, ...)` build `factory(env_name, env_args, termination_time, thread_index)`
    -- typically a BatchedFrameEnvironment over host simulators (lab / gym / indoor) -- and let
    get_action_size / get_objective_size answer for it, the way environment.py:30-72 dispatches on env_type."""
    if env_type in ('maze', 'synthetic'):
      raise _lib.UnrealError("env_type %r is built in" % (env_type,))
    Environment._registry[env_type] = (factory, int(action_size), int(objective_size))

  @staticmethod
  def create_environment(env_type, env_name, termination_time=50.0, env_args=None, thread_index=0):
    """environment.py:30-43.  env_args for the batched extension: {'num_envs': N,
    'device': 'cuda:0', 'obs_dtype': torch.float32|torch.uint8, 'auto_reset': bool}."""
    env_args = env_args or {}
    if env_type == 'maze':
      from . import maze_environment
      if 'num_envs' in env_args:
        return maze_environment.BatchedMazeEnvironment(**env_args)
      return maze_environment.MazeEnvironment(device=env_args.get('device', 'cuda:0'))
    if env_type == 'synthetic':
      from . import synthetic_environment
      return synthetic_environment.SyntheticIndoorEnvironment(env_name, env_args, termination_time, thread_index)
    if env_type in Environment._registry:
      return Environment._registry[env_type][0](env_name, env_args, termination_time, thread_index)
    raise _lib.UnrealError(
        "env_type %r needs an external simulator (deepmind_lab / gym / MINOS) that is outside the "
        "B200 hot path; use 'maze' or 'synthetic', or Environment.register() a frame producer for it" % (env_type,))

  @staticmethod
  def get_action_size(env_type, env_name):
    """environment.py:45-66 (the size is cached in the class attribute)."""
    if Environment.action_size >= 0:
      return Environment.action_size
    if env_type == 'maze':
      from . import maze_environment
      Environment.action_size = maze_environment.MazeEnvironment.get_action_size()
    elif env_type == 'synthetic':
      from . import synthetic_environment
      Environment.action_size = synthetic_environment.SyntheticIndoorEnvironment.get_action_size(env_name)
    elif env_type in Environment._registry:
      Environment.action_size = Environment._registry[env_type][1]
    else:
      raise _lib.UnrealError("env_type %r is outside the B200 hot path" % (env_type,))
    return Environment.action_size

  @staticmethod
  def get_objective_size(env_type, env_name):
    """environment.py:68-72."""
    if env_type == 'synthetic':
      from . import synthetic_environment
      return synthetic_environment.SyntheticIndoorEnvironment.get_objective_size(env_name)
    if env_type in Environment._registry:
      return Environment._registry[env_type][2]
    return 0

  def __init__(self):
    pass

  def process(self, action):
    pass

  def reset(self):
    pass

  def stop(self):
    pass

  def is_all_scheduled_episodes_done(self):
    return False

  # The two helpers below keep the reference's numpy-in / numpy-out contract but compute on
  # the device (K2); they fail loudly without one.
  def _subsample(self, a, average_width):
    """environment.py:88-91: mean over average_width x average_width blocks (columns, then rows)."""
    a = np.asarray(a)
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda().unsqueeze(0)
    out = K.subsample(t, int(average_width))[0].cpu().numpy()      # unreal_subsample
    return out.astype(np.float64) if a.dtype == np.float64 else out

  def _calc_pixel_change(self, state, last_state):
    """environment.py:93-99 through unreal_pixel_change (K2)."""
    try:
      cur = torch.as_tensor(np.ascontiguousarray(state, dtype=np.float32)).cuda().unsqueeze(0)
      prev = torch.as_tensor(np.ascontiguousarray(last_state, dtype=np.float32)).cuda().unsqueeze(0)
      out = K.pixel_change(cur, prev)[0].cpu().numpy()
      return out.astype(np.asarray(state).dtype) if np.asarray(state).dtype == np.float64 else out
    except Exception as e:
      print(str(e))
      raise Exception("Exception inside calc_pixel_change")   # environment.py:100-102
