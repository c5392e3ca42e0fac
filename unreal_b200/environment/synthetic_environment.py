"""Synthetic stand-in for IndoorEnvironment (environment/indoor_environment.py): frames of the
MINOS observation shape from a seeded device generator, because MINOS scenes are unavailable
offline (BASELINE.json north_star).  Observation contract of indoor_environment.py:63-139:
state = {'image': [H,W,3] (uint8 here, /255 on load), 'objective': f32 [G]}, 3 actions (:16-20), pixel
change from the generic kernel K2 (frames have no closed form).  Batched over num_envs: a
`BatchedFrameEnvironment` (frame_environment.py) over a `RandomFrameProducer`, or over the
deterministic `TableFrameProducer` with env_args={'producer': 'table'} (parity tests).
"""
from . import frame_environment as FE


class SyntheticIndoorEnvironment(FE.BatchedFrameEnvironment):
  ACTIONS = 3           # indoor_environment.py:16-20
  OBJECTIVE_SIZE = 0    # reference default when sim_config has none (:27-29)

  @staticmethod
  def get_action_size(env_name=None):
    return SyntheticIndoorEnvironment.ACTIONS

  @staticmethod
  def get_objective_size(env_name=None):
    return SyntheticIndoorEnvironment.OBJECTIVE_SIZE

  def __init__(self, env_name='', env_args=None, termination_time=50.0, thread_index=0):
    a = dict(env_args or {})
    n = int(a.get('num_envs', 1))
    device = a.get('device', 'cuda:0')
    shape = (int(a.get('height', 84)), int(a.get('width', 84)), 3)
    seed = int(a.get('seed', 0)) + thread_index
    producer = a.get('producer', 'random')
    if producer == 'table':
      producer = FE.TableFrameProducer(n, device, self.ACTIONS, seed=seed, first_env=int(a.get('first_env', 0)),
                                       frame_shape=shape)
    elif producer == 'random':
      producer = FE.RandomFrameProducer(n, device, self.ACTIONS, seed=seed, episode_len=int(a.get('episode_len', 200)),
                                        frame_shape=shape,
                                        objective_size=int(a.get('objective_size', self.OBJECTIVE_SIZE)))
    FE.BatchedFrameEnvironment.__init__(self, producer, device)
    self.h, self.w = shape[:2]
