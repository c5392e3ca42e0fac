"""MazeEnvironment: drop-in for environment/maze_environment.py, computed by K1 on the device.

`MazeEnvironment()` is the reference's scalar object (python int action in, numpy out) so the
reference's own callers and tests run against it unchanged; `BatchedMazeEnvironment` is the
additive batched extension that steps N mazes per launch and keeps everything
in HBM.
"""
import numpy as np
import torch

from .. import _lib
from .. import kernels as K
from . import environment


class BatchedMazeEnvironment(environment.Environment):
  """N mazes stepped by one kernel launch.  process(actions[N] int32 cuda) ->
  ({'image': [N,84,84,3]}, reward[N] f32, terminal[N] u8, pixel_change[N,20,20] f32)."""

  @staticmethod
  def get_action_size():
    return 4

  def __init__(self, num_envs, device='cuda:0', obs_dtype=torch.float32, auto_reset=False, map_data=None):
    environment.Environment.__init__(self)
    _lib.require_device()
    self.num_envs = int(num_envs)
    self.device = torch.device(device)
    self.obs_dtype = obs_dtype
    self.auto_reset = auto_reset
    with torch.cuda.device(self.device):
      K.maze_set_map(map_data)
      (self._start_pos, self._goal_pos, walls) = K.maze_layout()
      self._walls = np.frombuffer(walls, np.uint8).reshape(7, 7).astype(bool)
      self.state = K.MazeState(self.num_envs, self.device)
      n = self.num_envs
      self._obs = torch.empty(n, *K.obs_shape(obs_dtype), dtype=obs_dtype, device=self.device)
      self._pc = torch.empty(n, 20, 20, dtype=torch.float32, device=self.device)
      self._reward = torch.empty(n, dtype=torch.float32, device=self.device)
      self._terminal = torch.empty(n, dtype=torch.uint8, device=self.device)
      self.frame_rec = torch.zeros(n, dtype=torch.int64, device=self.device)
      self.reset()

  # device views of the reference's attributes
  @property
  def last_action(self):
    return self.state.last_action

  @property
  def last_reward(self):
    return self.state.last_reward

  def reset(self, mask=None):
    """maze_environment.py:50-55 for every env (or those with mask != 0)."""
    with torch.cuda.device(self.device):
      K.maze_reset(self.state, mask)
      K.maze_render(self.state.pos, self._obs)
    self.last_state = {'image': self._obs}

  def process(self, action, active=None, out_obs=None, out_pc=None, out_reward=None, out_terminal=None):
    """maze_environment.py:98-128 for all envs.  The returned tensors are owned by the env
    and overwritten by the next call unless out_* buffers are supplied.  out_pc=False: no pixel-change map is
    written (returned as None) -- the record ring re-derives it from the two cells when a frame is replayed, so the
    batched Trainer has no reader for the 1.6 KB per env and step."""
    obs = self._obs if out_obs is None else out_obs
    pc = None if out_pc is False else (self._pc if out_pc is None else out_pc)
    reward = self._reward if out_reward is None else out_reward
    terminal = self._terminal if out_terminal is None else out_terminal
    with torch.cuda.device(self.device):
      K.maze_step(self.state, action, obs=obs, pc=pc, reward=reward, terminal=terminal,
                  frame_rec=self.frame_rec, active=active, auto_reset=self.auto_reset)
    state = {'image': obs}
    self.last_state = state
    return state, reward, terminal, pc


class MazeEnvironment(environment.Environment):
  """The reference's scalar maze (one env), same attributes and return types:
  process(action) -> (image float64 [84,84,3], int reward, bool terminal, pixel_change [20,20])."""

  @staticmethod
  def get_action_size():
    return 4   # maze_environment.py:11-13

  def __init__(self, device='cuda:0'):
    environment.Environment.__init__(self)
    self._map_data = "--+---G" "--+-+++" "S-+---+" "--+++--" "--+-+--" "--+----" "-----++"  # :17-25
    self._b = BatchedMazeEnvironment(1, device=device, auto_reset=False, map_data=self._map_data)
    self._start_pos = self._b._start_pos
    self._goal_pos = self._b._goal_pos
    self._action = torch.zeros(1, dtype=torch.int32, device=self._b.device)
    self.reset()

  @property
  def x(self):
    return int(self._b.state.pos[0, 0])

  @property
  def y(self):
    return int(self._b.state.pos[0, 1])

  def _get_pixel(self, x, y):
    return self._map_data[y * 7 + x]       # :62-64

  def _is_wall(self, x, y):
    return self._get_pixel(x, y) == '+'    # :66-67

  def _get_current_image(self):
    """:93-96 (rendered on the device)."""
    return K.maze_render(self._b.state.pos)[0].cpu().numpy().astype(np.float64)

  def reset(self):
    self._b.reset()
    self.last_state = {'image': self._b._obs[0].cpu().numpy().astype(np.float64)}
    self.last_action = 0
    self.last_reward = 0

  def process(self, action):
    a = int(action)
    self._action.fill_(a if 0 <= a < 2 ** 31 else -1)
    _, r, t, pc = self._b.process(self._action)
    image = self._b._obs[0].cpu().numpy().astype(np.float64)
    reward = int(r.item())
    terminal = bool(t.item())
    pixel_change = pc[0].cpu().numpy().astype(np.float64)
    self.last_state = {'image': image}     # :125 (a dict, as in the fork)
    self.last_action = action
    self.last_reward = reward
    return image, reward, terminal, pixel_change
