"""Environment factory (drop-in for environment/environment.py:30-72): built-in types, the integrator's registry
for simulator-backed env types and the reference's error behaviour.  CPU only: nothing is
constructed on a device here."""
import pytest


def test_registered_env_type_is_dispatched_and_sized():
  from unreal_b200 import _lib
  from unreal_b200.environment.environment import Environment
  calls = []

  def factory(env_name, env_args, termination_time, thread_index):
    calls.append((env_name, dict(env_args), termination_time, thread_index))
    return "the-env"

  try:
    Environment.register('indoor', factory, action_size=3, objective_size=2)
    Environment.action_size = -1
    assert Environment.get_action_size('indoor', 'pointgoal') == 3
    assert Environment.get_objective_size('indoor', 'pointgoal') == 2
    # the reference caches the action size in a class attribute (environment.py:13, :46-47): so does the drop-in
    assert Environment.get_action_size('maze', '') == 3
    env = Environment.create_environment('indoor', 'pointgoal', 25.0, {'num_envs': 4, 'device': 'cuda:0'}, thread_index=2)
    assert env == "the-env" and calls == [('pointgoal', {'num_envs': 4, 'device': 'cuda:0'}, 25.0, 2)]
    with pytest.raises(_lib.UnrealError):
      Environment.register('maze', factory, 4)
    with pytest.raises(_lib.UnrealError, match="register"):
      Environment.create_environment('lab', 'nav_maze_static_01')
  finally:
    Environment._registry.pop('indoor', None)
    Environment.action_size = -1
  assert Environment.get_objective_size('maze', '') == 0
