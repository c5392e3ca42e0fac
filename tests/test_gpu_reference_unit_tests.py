"""The reference's OWN unit tests, un-staled against the fork's current signatures and run
against the scalar drop-in classes -- python ints / numpy in and out, one env, exactly the objects a caller
of the reference would hold:

  environment/environment_test.py:36-54      -> test_environment_maze  (+ the reference's golden trajectory)
  train/experience_test.py:16-36             -> test_experience_process
  train/rmsprop_applier_test.py:9-53         -> test_rmsprop_apply (through _apply_gradients / get_slot)
  model/model_test.py:8-59                   -> test_unreal_/pc_/vr_/rp_variable_size (20 / 18 / 12 / 14)

What had gone stale in the fork and how each test is un-staled is said beside it.  The assertions themselves are
the reference's.  One more test drives `minimize_local` (rmsprop_applier.py:95-106), which no reference test does.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import unreal_oracle as O

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- environment_test.py
def _check_environment(env_type, env_name):
  """environment_test.py:36-54.  Stale line: `environment.last_state.shape` (:45) -- `last_state` is a dict in the
  fork (maze_environment.py:53, :125), so the image inside it is what has the shape."""
  from unreal_b200.environment.environment import Environment
  environment = Environment.create_environment(env_type, env_name)
  for i in range(3):
    state, reward, terminal, pixel_change = environment.process(0)
    assert isinstance(state, np.ndarray) and isinstance(pixel_change, np.ndarray)
    assert isinstance(reward, int) and isinstance(terminal, bool)
    # Check shape
    assert state.shape == (84, 84, 3)
    assert environment.last_state['image'].shape == (84, 84, 3)
    assert pixel_change.shape == (20, 20)
    # state and pixel_change value range should be [0,1]
    assert np.amax(state) <= 1.0
    assert np.amin(state) >= 0.0
    assert np.amax(pixel_change) <= 1.0
    assert np.amin(pixel_change) >= 0.0
  environment.stop()


def test_environment_maze():
  _check_environment("maze", "")


def test_environment_lab_and_gym_fail_loudly():
  """test_lab skips without deepmind_lab (:13-24) and test_gym needs gym: both simulators are absent here and
  outside the hot path -- the factory must say so instead of returning something else."""
  from unreal_b200.environment.environment import Environment
  from unreal_b200 import _lib
  for env_type, env_name in (("lab", "nav_maze_static_01"), ("gym", "MontezumaRevenge-v0")):
    with pytest.raises(_lib.UnrealError):
      Environment.create_environment(env_type, env_name)


def test_scalar_maze_follows_the_reference_trajectory(golden_dir):
  """The scalar MazeEnvironment (python int action in, numpy out) on the reference's own seed-0xA3C trajectory
  (tests/golden/maze_golden.npz, written by the reference's MazeEnvironment): positions, rewards, terminals,
  float64 frames and pixel-change maps, with the caller-side reset of trainer.py:201-202."""
  from unreal_b200.environment.environment import Environment
  with np.load(os.path.join(golden_dir, "maze_golden.npz")) as z:
    g = {k: z[k] for k in z.files}
  env = Environment.create_environment("maze", "")
  assert env.last_action == 0 and env.last_reward == 0
  assert np.array_equal(env.last_state['image'].astype(np.uint8), g["initial_frame"])
  assert env.last_state['image'].dtype == np.float64           # maze_environment.py:31 `dtype=float`
  for i in range(900):
    a = int(g["actions"][i])
    image, reward, terminal, pc = env.process(a)
    assert (env.x, env.y, reward, int(terminal)) == (g["x"][i], g["y"][i], g["reward"][i], g["terminal"][i]), i
    assert env.last_action == a and env.last_reward == reward
    if i < 64:
      assert np.array_equal(pc.astype(np.float32), g["first_pc"][i].astype(np.float32)), i
      assert np.array_equal(image, O.maze_render(env.x, env.y)), i
    if terminal:
      env.reset()
  env.stop()


def test_calc_pixel_change_and_subsample_scalar(golden_dir):
  """Environment._calc_pixel_change / _subsample (environment.py:88-99) with numpy in / numpy out: the reference's
  maps of tests/golden/pixel_change_golden.npz (uint8/255 float32 frames, five geometries) bit for bit."""
  from unreal_b200.environment.environment import Environment
  env = Environment()
  with np.load(os.path.join(golden_dir, "pixel_change_golden.npz")) as z:
    g = {k: z[k] for k in z.files}
  for i in range(5):
    a = g["a%d" % i].astype(np.float32) / np.float32(255.0)
    b = g["b%d" % i].astype(np.float32) / np.float32(255.0)
    got = env._calc_pixel_change(a, b)
    assert got.dtype == np.float32 and np.array_equal(got, g["pc%d" % i]), i
  rs = np.random.RandomState(5)
  for h, w, width in ((80, 80, 4), (24, 36, 4), (30, 20, 5), (16, 16, 2)):
    m = rs.rand(h, w).astype(np.float32)
    assert np.array_equal(env._subsample(m, width), O.subsample(m, width)), (h, w, width)
  m64 = rs.rand(80, 80)
  got = env._subsample(m64, 4)
  assert got.dtype == np.float64 and np.allclose(got, O.subsample(m64, 4), rtol=1e-6)


# ----------------------------------------------------------------------------- experience_test.py
def _add_frame(experience, reward):
  from unreal_b200.train.experience import ExperienceFrame
  frame = ExperienceFrame(0, reward, 0, False, 0, 0, 0)
  experience.add_frame(frame)


def test_experience_process():
  """experience_test.py:16-36.  Stale line: `Experience(10)` (:17) -- the constructor now takes the shared
  RandomState (experience.py:49)."""
  from unreal_b200.train.experience import Experience
  experience = Experience(10, np.random.RandomState(0))
  for i in range(10):
    if i == 5:
      _add_frame(experience, 1)
    else:
      _add_frame(experience, 0)
  assert experience.is_full()
  assert experience._top_frame_index == 0
  _add_frame(experience, 0)
  assert experience._top_frame_index == 1
  rewards = []
  for i in range(100):
    frames = experience.sample_rp_sequence()
    assert len(frames) == 4
    rewards.append(frames[3].reward)
  # "Reward should be skewed here": about half of the draws end on the single positive frame
  assert 25 <= sum(1 for r in rewards if r > 0) <= 75


def test_experience_draws_equal_the_reference_ring():
  """The scalar Experience against the oracle ring (pinned to the reference's traces) on one shared RandomState
  stream each: same sampled frames, and the caller's RandomState ends in the same state."""
  from unreal_b200.train.experience import Experience, ExperienceFrame
  rs_a, rs_b = np.random.RandomState(0xA3C), np.random.RandomState(0xA3C)
  ex = Experience(40, rs_a)
  ring = O.RingOracle(40, rs_b)
  env_rs = np.random.RandomState(3)
  for i in range(130):
    reward = int(env_rs.choice([-1, 0, 0, 1]))
    terminal = bool(env_rs.rand() < 0.08)
    ex.add_frame(ExperienceFrame({'id': i}, reward, i % 4, terminal, None, 0, 0))
    ring.add({'id': i, 'reward': reward, 'terminal': terminal})
    assert ex._top_frame_index == ring.top
    if ex.is_full() and i % 3 == 0:
      assert [f.state['id'] for f in ex.sample_sequence(21)] == [f['id'] for f in ring.sample_sequence(21)]
      assert [f.state['id'] for f in ex.sample_rp_sequence()] == [f['id'] for f in ring.sample_rp_sequence()]
  assert rs_a.randint(0, 1 << 30) == rs_b.randint(0, 1 << 30)


# ----------------------------------------------------------------------------- rmsprop_applier_test.py
def test_rmsprop_apply():
  """rmsprop_applier_test.py:9-53.  Stale lines: the bare `import rmsprop_applier` (:6) and `.run()` on what is now
  a `(group_op, global_norm)` tuple (:27, :44; rmsprop_applier.py:129) -- here `_apply_gradients` applies eagerly and
  returns `(None, global_norm)`.  The expected values are computed by the reference test's own formula."""
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  dev = "cuda:0"
  var = torch.tensor([1.0, 2.0], device=dev)
  grad0 = torch.tensor([2.0, 4.0], device=dev)
  grad1 = torch.tensor([3.0, 6.0], device=dev)
  opt = RMSPropApplier(learning_rate=2.0, decay=0.9, momentum=0.0, epsilon=1.0)
  assert opt.get_slot(var, "rms") is None                      # no slots before the first apply (:116-117)

  # apply grad0
  op, norm = opt._apply_gradients([var], [grad0])
  assert op is None
  ms_x = 1.0
  ms_y = 1.0
  x = 1.0
  y = 2.0
  dx = 2.0
  dy = 4.0
  ms_x = ms_x + (dx * dx - ms_x) * (1.0 - 0.9)
  ms_y = ms_y + (dy * dy - ms_y) * (1.0 - 0.9)
  x = x - (2.0 * dx / math.sqrt(ms_x + 1.0))
  y = y - (2.0 * dy / math.sqrt(ms_y + 1.0))
  np.testing.assert_allclose(np.array([x, y]), var.cpu().numpy(), rtol=1e-6)
  np.testing.assert_allclose(float(norm), math.sqrt(dx * dx + dy * dy), rtol=1e-6)
  np.testing.assert_allclose(opt.get_slot(var, "rms").cpu().numpy(), [ms_x, ms_y], rtol=1e-6)      # :38-43 rms0 = 1
  # the `momentum` slot starts at 0 (:38-43) and, as in TF's ApplyRMSProp, holds the step just taken even with
  # momentum = 0:  mom = momentum * mom + lr * g / sqrt(ms + eps);  var -= mom
  np.testing.assert_allclose(opt.get_slot(var, "momentum").cpu().numpy(),
                             [2.0 * dx / math.sqrt(ms_x + 1.0), 2.0 * dy / math.sqrt(ms_y + 1.0)], rtol=1e-6)

  # apply grad1
  opt._apply_gradients([var], [grad1])
  dx = 3.0
  dy = 6.0
  ms_x = ms_x + (dx * dx - ms_x) * (1.0 - 0.9)
  ms_y = ms_y + (dy * dy - ms_y) * (1.0 - 0.9)
  x = x - (2.0 * dx / math.sqrt(ms_x + 1.0))
  y = y - (2.0 * dy / math.sqrt(ms_y + 1.0))
  np.testing.assert_allclose(np.array([x, y]), var.cpu().numpy(), rtol=1e-6)
  np.testing.assert_allclose(opt.get_slot(var, "rms").cpu().numpy(), [ms_x, ms_y], rtol=1e-6)
  assert opt.get_slot(var, "no_such_slot") is None


@pytest.mark.parametrize("momentum", [0.0, 0.9])
def test_minimize_local_applies_local_gradients_to_global_vars(momentum):
  """rmsprop_applier.py:95-106: gradients of the loss w.r.t. the LOCAL variables, clipped by their global norm
  (:121) and applied to the GLOBAL variables with shared slots -- several variables of different shapes, a gradient
  norm above the clip, two successive calls; against the oracle's per-variable restatement."""
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  dev = "cuda:0"
  rs = np.random.RandomState(11)
  shapes = [(8, 8, 3, 16), (16,), (37, 5), (1,), (1031,)]
  g_np = [rs.randn(*s).astype(np.float32) for s in shapes]
  global_vars = [torch.from_numpy(v.copy()).to(dev) for v in g_np]
  handles = list(global_vars)
  local_vars = [torch.from_numpy(rs.randn(*s).astype(np.float32)).to(dev).requires_grad_(True) for s in shapes]
  coef = [torch.from_numpy(rs.randn(*s).astype(np.float32) * 3.0).to(dev) for s in shapes]
  opt = RMSPropApplier(learning_rate=7e-4, decay=0.99, momentum=momentum, epsilon=0.1, clip_norm=40.0)
  rms = [np.ones(s, np.float32) for s in shapes]
  mom = [np.zeros(s, np.float32) for s in shapes]
  want = [v.copy() for v in g_np]
  for it in range(2):
    loss = sum((c * v).sum() + 0.5 * (v * v).sum() for c, v in zip(coef, local_vars))
    grads = [(c + v).detach().cpu().numpy() for c, v in zip(coef, local_vars)]
    op, norm = opt.minimize_local(loss, global_vars, local_vars, thread_index=0)
    assert op is None
    n = O.rmsprop_step(want, rms, mom, grads, 7e-4, 0.99, momentum, 0.1, 40.0, np.float32)
    assert n > 40.0                                              # the clip is active
    np.testing.assert_allclose(float(norm), n, rtol=1e-5)
    for v, w, h in zip(global_vars, want, handles):
      assert v is h                                              # the caller's tensors themselves were updated
      np.testing.assert_allclose(v.cpu().numpy(), w, rtol=1e-5, atol=1e-7)
    for v, r, m in zip(global_vars, rms, mom):
      np.testing.assert_allclose(opt.get_slot(v, "rms").cpu().numpy(), r, rtol=1e-5)
      np.testing.assert_allclose(opt.get_slot(v, "momentum").cpu().numpy(), m, rtol=1e-5, atol=1e-9)
    with torch.no_grad():                                        # the "sync" of trainer.py:457 for the next iteration
      for lv, gv in zip(local_vars, global_vars):
        lv.copy_(gv)
  # local variables were only read
  assert all(lv.grad is None for lv in local_vars)


# ----------------------------------------------------------------------------- model_test.py
def _check_model_var_size(use_pixel_change, use_value_replay, use_reward_prediction, var_size):
  """model_test.py:60-79.  Stale line: the 10-positional-argument constructor call (:68-77); the fork's constructor
  (model.py:49-66) also takes segnet_param_dict / image_shape / is_training / n_classes / segnet_lambda / dropout."""
  from unreal_b200.model.model import UnrealModel
  use_lstm = True
  model = UnrealModel(1, 0, -1, use_lstm, use_pixel_change, use_value_replay, use_reward_prediction, 1.0, 1.0, "/cpu:0",
                      {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0)
  variables = model.get_vars()
  assert len(variables) == var_size
  return variables


def test_unreal_variable_size():
  """all options ON: base conv=4, fc=2, lstm=2, policy_fc=2, value_fc=2; pc fc=2, deconv_v=2, deconv_a=2; rp fc=2."""
  v = _check_model_var_size(True, True, True, 20)
  # and their shapes are the reference graph's (model.py:283-284, :333, :362, :373, :418-421, :482; A = 1, G = 0)
  assert [tuple(x.shape) for x in v] == [
      (8, 8, 3, 16), (16,), (4, 4, 16, 32), (32,), (2592, 256), (256,), (256 + 1 + 1 + 256, 1024), (1024,), (256, 1), (1,),
      (256, 1), (1,), (256, 2592), (2592,), (4, 4, 1, 32), (1,), (4, 4, 1, 32), (1,), (7776, 3), (3,)]


def test_pc_variable_size():
  _check_model_var_size(True, False, False, 18)


def test_vr_variable_size():
  _check_model_var_size(False, True, False, 12)


def test_rp_variable_size():
  _check_model_var_size(False, False, True, 14)
