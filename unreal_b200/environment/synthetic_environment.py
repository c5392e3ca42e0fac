"""Synthetic stand-in for IndoorEnvironment (environment/indoor_environment.py): frames of the
MINOS observation shape from a seeded device generator, because MINOS scenes are unavailable
offline (BASELINE.json north_star).  Observation contract of indoor_environment.py:63-139:
state = {'image': f32 [H,W,3] in [0,1], 'objective': f32 [G]}, 3 actions (:16-20), pixel change
from the generic kernel K2 (frames have no closed form).  Batched over num_envs.
"""
import torch

from .. import _lib
from .. import kernels as K
from . import environment


class SyntheticIndoorEnvironment(environment.Environment):
  ACTIONS = 3           # indoor_environment.py:16-20
  OBJECTIVE_SIZE = 0    # reference default when sim_config has none (:27-29)

  @staticmethod
  def get_action_size(env_name=None):
    return SyntheticIndoorEnvironment.ACTIONS

  @staticmethod
  def get_objective_size(env_name=None):
    return SyntheticIndoorEnvironment.OBJECTIVE_SIZE

  def __init__(self, env_name='', env_args=None, termination_time=50.0, thread_index=0):
    environment.Environment.__init__(self)
    _lib.require_device()
    a = env_args or {}
    self.num_envs = int(a.get('num_envs', 1))
    self.device = torch.device(a.get('device', 'cuda:0'))
    self.h = int(a.get('height', 84))
    self.w = int(a.get('width', 84))
    self.objective_size = int(a.get('objective_size', self.OBJECTIVE_SIZE))
    self.episode_len = int(a.get('episode_len', 200))
    self.gen = torch.Generator(device=self.device).manual_seed(int(a.get('seed', 0)) + thread_index)
    n = self.num_envs
    self._frames = [torch.empty(n, self.h, self.w, 3, dtype=torch.uint8, device=self.device) for _ in range(2)]
    self._cur = 0
    self._t = torch.zeros(n, dtype=torch.int32, device=self.device)
    self.reset()

  def _draw(self, out):
    out.copy_(torch.randint(0, 256, out.shape, dtype=torch.uint8, device=self.device, generator=self.gen))

  def _state(self, frame_u8):
    s = {'image': frame_u8}      # uint8; consumers divide by 255 like _preprocess_frame (:90-104)
    if self.objective_size:
      s['objective'] = torch.zeros(self.num_envs, self.objective_size, device=self.device)
    return s

  def reset(self):
    self._draw(self._frames[self._cur])
    self._t.zero_()
    self.last_state = self._state(self._frames[self._cur])
    self.last_action = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
    self.last_reward = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)

  def process(self, action, flag=1):
    prev = self._frames[self._cur]
    self._cur ^= 1
    cur = self._frames[self._cur]
    self._draw(cur)
    pc = K.pixel_change(cur, prev)
    self._t += 1
    terminal = (self._t >= self.episode_len).to(torch.uint8)
    reward = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
    self._t.mul_(1 - terminal.to(torch.int32))
    state = self._state(cur)
    self.last_state = state
    self.last_action = action
    self.last_reward = reward
    return state, reward, terminal, pc
