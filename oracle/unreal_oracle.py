"""numpy restatement of the UNREAL rollout-and-target hot path (the checker).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): imported only by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``.

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  Arithmetic follows the reference operation by operation
(same numpy reductions in the same order, float64 where the reference is
float64) so that it is also a fair stand-in for the reference's CPU cost.
Parity is pinned by ``tests/test_oracle_golden.py`` against fixtures produced
from the reference's own code by ``tests/golden/make_golden.py``.
"""
from collections import deque

import numpy as np

# --------------------------------------------------------------------------
# maze (environment/maze_environment.py)
# --------------------------------------------------------------------------

# maze_environment.py:18-25, row-major, index = y * 7 + x (:62-64)
MAZE_ROWS = ("--+---G",
             "--+-+++",
             "S-+---+",
             "--+++--",
             "--+-+--",
             "--+----",
             "-----++")
MAZE_MAP = "".join(MAZE_ROWS)
GRID = 7
CELL = 12
FRAME = GRID * CELL  # 84
PC_CELLS = 20

ACTION_DELTAS = ((0, -1), (0, 1), (-1, 0), (1, 0))  # UP, DOWN, LEFT, RIGHT (:99-108)


def maze_layout(map_data=MAZE_MAP):
  """walls[y, x] bool, start (x, y), goal (x, y)   (maze_environment.py:30-48)."""
  walls = np.zeros((GRID, GRID), dtype=bool)
  start = goal = (-1, -1)
  for y in range(GRID):
    for x in range(GRID):
      c = map_data[y * GRID + x]
      if c == '+':
        walls[y, x] = True
      elif c == 'S':
        start = (x, y)
      elif c == 'G':
        goal = (x, y)
  return walls, start, goal


WALLS, START, GOAL = maze_layout()


def maze_wall_image(dtype=np.float64, map_data=MAZE_MAP):
  """Walls drawn into channel 0 of an 84x84x3 zero image (maze_environment.py:30-41)."""
  walls, _, _ = maze_layout(map_data)
  image = np.zeros((FRAME, FRAME, 3), dtype=dtype)
  for y in range(GRID):
    for x in range(GRID):
      if walls[y, x]:
        image[CELL * y:CELL * (y + 1), CELL * x:CELL * (x + 1), 0] = 1.0
  return image


_WALL_IMAGE = maze_wall_image()


def maze_render(x, y, dtype=np.float64):
  """Wall image + 12x12 agent block in channel 1 (maze_environment.py:57-60, :93-96)."""
  image = np.array(_WALL_IMAGE, dtype=dtype)
  image[CELL * y:CELL * (y + 1), CELL * x:CELL * (x + 1), 1] = 1.0
  return image


def _clamp(n, lo, hi):
  # maze_environment.py:69-74
  if n < lo:
    return lo, True
  if n > hi:
    return hi, True
  return n, False


def maze_move(x, y, action, walls=WALLS):
  """-> (new_x, new_y, hit)   (maze_environment.py:76-91, :99-108).

  An action outside 0..3 is a zero displacement (no ``if`` fires), reward 0."""
  dx, dy = ACTION_DELTAS[action] if 0 <= action < 4 else (0, 0)
  nx, cx = _clamp(x + dx, 0, GRID - 1)
  ny, cy = _clamp(y + dy, 0, GRID - 1)
  hit_wall = False
  if walls[ny, nx]:
    nx, ny = x, y
    hit_wall = True
  return nx, ny, (cx or cy or hit_wall)


def maze_step(x, y, action, walls=WALLS, goal=GOAL):
  """-> (new_x, new_y, reward, terminal)   (maze_environment.py:110-122)."""
  nx, ny, hit = maze_move(x, y, action, walls)
  terminal = (nx == goal[0] and ny == goal[1])
  if terminal:
    reward = 1
  elif hit:
    reward = -1
  else:
    reward = 0
  return nx, ny, reward, terminal


def subsample(a, average_width):
  # environment.py:88-91
  s = a.shape
  sh = s[0] // average_width, average_width, s[1] // average_width, average_width
  return a.reshape(sh).mean(-1).mean(1)


def pixel_change(state, last_state):
  """abs-diff, crop 2, channel mean, 4x4 mean   (environment.py:93-99)."""
  d = np.absolute(state[2:-2, 2:-2, :] - last_state[2:-2, 2:-2, :])
  m = np.mean(d, 2)
  return subsample(m, 4)


def _overlap_weights(c):
  """Pixels shared by pool block i (rows 4i+2..4i+5) and maze cell c (12c..12c+11)."""
  w = np.zeros(PC_CELLS, dtype=np.int64)
  for i in range(PC_CELLS):
    lo = max(4 * i + 2, CELL * c)
    hi = min(4 * i + 5, CELL * c + CELL - 1)
    w[i] = max(0, hi - lo + 1)
  return w


def maze_pixel_change_closed_form(x0, y0, x1, y1):
  """Closed form of ``pixel_change(render(x1,y1), render(x0,y0))`` as float32.

  Not in the reference: it is the identity the fused CUDA kernel relies on
  (SURVEY.md 8a note).  ``tests/test_oracle_golden.py`` proves it equal to the
  literal computation for every (free cell, action) pair."""
  if x0 == x1 and y0 == y1:
    return np.zeros((PC_CELLS, PC_CELLS), dtype=np.float32)
  k = (np.outer(_overlap_weights(y0), _overlap_weights(x0)) +
       np.outer(_overlap_weights(y1), _overlap_weights(x1)))
  return k.astype(np.float32) / np.float32(48.0)


class MazeOracle(object):
  """Stateful single maze with the reference's ``process``/``reset`` contract
  (maze_environment.py:50-55, :98-128)."""

  def __init__(self):
    self.reset()

  def reset(self):
    self.x, self.y = START
    self.last_state = {'image': maze_render(self.x, self.y)}
    self.last_action = 0
    self.last_reward = 0

  def process(self, action):
    self.x, self.y, reward, terminal = maze_step(self.x, self.y, action)
    image = maze_render(self.x, self.y)
    pc = pixel_change(image, self.last_state['image'])
    self.last_state = {'image': image}
    self.last_action = action
    self.last_reward = reward
    return image, reward, terminal, pc


# --------------------------------------------------------------------------
# replay (train/experience.py)
# --------------------------------------------------------------------------

def concat_action_and_reward(action, action_size, reward, objective=None):
  """one-hot(action) ++ [reward] (++ objective)   (experience.py:34-46)."""
  v = np.zeros([action_size + 1])
  v[action] = 1.0
  v[-1] = float(reward)
  if objective is not None:
    return np.concatenate((v, objective))
  return v


class RingOracle(object):
  """Index-level restatement of ``Experience`` (experience.py:48-153).

  Frames are opaque records with ``.reward`` and ``.terminal``-like fields; here
  each frame is a dict so tests can carry any payload."""

  def __init__(self, history_size, random_state):
    self.history_size = history_size
    self.frames = deque(maxlen=history_size)
    self.pos_idx = deque()
    self.neg_idx = deque()
    self.top = 0
    self.random_state = random_state

  def is_full(self):
    return len(self.frames) >= self.history_size  # :96-97

  def add(self, frame):
    """experience.py:63-93.  Returns False when the frame was discarded."""
    if frame['terminal'] and len(self.frames) > 0 and self.frames[-1]['terminal']:
      return False
    index = self.top + len(self.frames)
    was_full = self.is_full()
    self.frames.append(frame)
    if index >= 3:
      (self.pos_idx if frame['reward'] > 0 else self.neg_idx).append(index)
    if was_full:
      self.top += 1
      cut = self.top + 3
      if len(self.pos_idx) > 0 and self.pos_idx[0] < cut:
        self.pos_idx.popleft()
      if len(self.neg_idx) > 0 and self.neg_idx[0] < cut:
        self.neg_idx.popleft()
    return True

  def sample_sequence_index(self, sequence_size):
    """-> (raw start position, number of frames)   (experience.py:100-118)."""
    start = self.random_state.randint(0, self.history_size - sequence_size - 1)
    if self.frames[start]['terminal']:
      start += 1
    n = 0
    for i in range(sequence_size):
      n += 1
      if self.frames[start + i]['terminal']:
        break
    return start, n

  def sample_sequence(self, sequence_size):
    start, n = self.sample_sequence_index(sequence_size)
    return [self.frames[start + i] for i in range(n)]

  def sample_rp_index(self):
    """-> raw position of the first of the 4 frames   (experience.py:121-153)."""
    from_neg = (self.random_state.randint(2) == 0)
    if len(self.pos_idx) == 0:
      from_neg = True
    elif len(self.neg_idx) == 0:
      from_neg = False
    lst = self.neg_idx if from_neg else self.pos_idx
    k = self.random_state.randint(len(lst))
    end = lst[k]
    return end - 3 - self.top

  def sample_rp_sequence(self):
    s = self.sample_rp_index()
    return [self.frames[s + i] for i in range(4)]


# --------------------------------------------------------------------------
# targets (train/trainer.py)
# --------------------------------------------------------------------------

def nstep_returns(rewards, values, bootstrap, gamma, dtype=np.float64):
  """One rollout, time order in / time order out   (trainer.py:298-324).

  ``bootstrap`` is 0 when the rollout ended in a terminal, else V(new_state).
  ``dtype`` float64 = numpy<2 promotion of the reference's era; float32 = what
  numpy>=2 computes when ``bootstrap`` is a float32 network output."""
  R = dtype(bootstrap)
  g = dtype(gamma)
  T = len(rewards)
  out_R = np.zeros(T, dtype=dtype)
  out_adv = np.zeros(T, dtype=dtype)
  for i in range(T - 1, -1, -1):
    R = dtype(rewards[i]) + g * R
    out_R[i] = R
    out_adv[i] = R - dtype(values[i])
  return out_R, out_adv


def nstep_returns_segmented(r, v, term, boot, gamma, dtype=np.float32):
  """Batched window form: r, v, term are [T, N]; boot [N].

  ``R_t = r_t + gamma * (term_t ? 0 : R_{t+1})`` with ``R_T = boot``: for a
  window that is one reference rollout (its only terminal, if any, is its last
  step) this is exactly ``nstep_returns`` (trainer.py:298-324)."""
  T, N = r.shape
  R = np.asarray(boot, dtype=dtype).copy()
  g = dtype(gamma)
  out_R = np.zeros((T, N), dtype=dtype)
  out_adv = np.zeros((T, N), dtype=dtype)
  for t in range(T - 1, -1, -1):
    R = np.where(term[t] != 0, dtype(0), R)
    R = r[t].astype(dtype) + g * R
    out_R[t] = R
    out_adv[t] = R - v[t].astype(dtype)
  return out_R, out_adv


def pc_targets(pixel_changes, bootstrap, gamma_pc):
  """``pc_R = pc + gamma_pc * pc_R`` backwards over the n-1 batch frames
  (trainer.py:352-372).  pixel_changes [n, 20, 20] in time order; bootstrap
  [20, 20] = max_a Q of the frame after the last one.  -> [n, 20, 20]."""
  R = np.asarray(bootstrap)
  out = [None] * len(pixel_changes)
  for i in range(len(pixel_changes) - 1, -1, -1):
    R = pixel_changes[i] + gamma_pc * R
    out[i] = R
  return np.stack(out) if out else np.zeros((0, PC_CELLS, PC_CELLS))


def vr_returns(rewards, bootstrap, gamma):
  """``vr_R = r + gamma * vr_R`` backwards   (trainer.py:394-403)."""
  R = bootstrap
  out = [None] * len(rewards)
  for i in range(len(rewards) - 1, -1, -1):
    R = rewards[i] + gamma * R
    out[i] = R
  return out


def rp_target(reward):
  """[zero, positive, negative] one-hot   (trainer.py:426-434)."""
  c = [0.0, 0.0, 0.0]
  if -1e-10 < reward < 1e-10:
    c[0] = 1.0
  elif reward > 0:
    c[1] = 1.0
  else:
    c[2] = 1.0
  return c


def anneal_learning_rate(initial_lr, global_t, max_global_t):
  """trainer.py:140-144."""
  lr = initial_lr * (max_global_t - global_t) / max_global_t
  return lr if lr >= 0.0 else 0.0


# --------------------------------------------------------------------------
# optimiser (train/rmsprop_applier.py + TensorFlow training_ops.apply_rms_prop)
# --------------------------------------------------------------------------

def global_norm(grads, dtype=np.float32):
  """``tf.global_norm``: sqrt(sum_i 2 * l2_loss(g_i)) = sqrt(sum of squares)."""
  s = dtype(0)
  for g in grads:
    s = s + np.sum(np.square(np.asarray(g, dtype=dtype)), dtype=dtype)
  return np.sqrt(s)


def clip_by_global_norm(grads, clip_norm, dtype=np.float32):
  """``tf.clip_by_global_norm`` as called at rmsprop_applier.py:121.  TensorFlow documents
  ``g * clip / max(norm, clip)`` and evaluates it as ``g * (clip * min(1/norm, 1/clip))``
  (tensorflow/python/ops/clip_ops.py); the latter is restated.  -> (clipped list, norm)."""
  norm = global_norm(grads, dtype)
  one = dtype(1.0)
  scale = dtype(clip_norm) * min(one / norm, one / dtype(clip_norm)) if norm > 0 else one
  return [np.asarray(g, dtype=dtype) * scale for g in grads], norm


def rmsprop_apply(var, rms, mom, grad, lr, decay, momentum, epsilon, dtype=np.float32):
  """TensorFlow ``ApplyRMSProp`` dense kernel, argument order of
  rmsprop_applier.py:83-93; arithmetic pinned by rmsprop_applier_test.py:31-51:
  ``ms += (g*g - ms) * (1 - decay); mom = momentum*mom + lr*g/sqrt(ms+eps);
  var -= mom``.  Updates in place, returns (var, rms, mom)."""
  g = np.asarray(grad, dtype=dtype)
  rms += (g * g - rms) * dtype(1.0 - decay)
  mom[...] = mom * dtype(momentum) + dtype(lr) * g / np.sqrt(rms + dtype(epsilon))
  var -= mom
  return var, rms, mom


def rmsprop_step(vars_, rms_, mom_, grads, lr, decay=0.9, momentum=0.0, epsilon=1e-10,
                 clip_norm=40.0, dtype=np.float32):
  """clip (rmsprop_applier.py:121) then per-variable apply (:124-128).
  Slots start at rms=1, momentum=0 (:38-43).  -> global grad norm."""
  clipped, norm = clip_by_global_norm(grads, clip_norm, dtype)
  for v, r, m, g in zip(vars_, rms_, mom_, clipped):
    rmsprop_apply(v, r, m, g, lr, decay, momentum, epsilon, dtype)
  return norm


# --------------------------------------------------------------------------
# one worker's rollout + replay targets (train/trainer.py:176-205, :218-436)
# --------------------------------------------------------------------------

def make_frame_table(seed, frame_shape=(84, 84, 3), k_frames=32):
  """The seeded table of uint8 frames behind TableFrameEnvOracle (same formula as the device producer's)."""
  return np.random.RandomState(seed).randint(0, 256, size=(k_frames,) + tuple(frame_shape)).astype(np.uint8)


class TableFrameEnvOracle(object):
  """A generic-frame env with the instance contract of the reference's lab / gym / indoor classes
  (environment/lab_environment.py:94-131, indoor_environment.py:63-139): `last_state['image']` is
  `uint8 / 255.0` as float32, `process(action) -> (state, reward, terminal, pixel_change)` with the
  pixel change from environment.py:93-99 on the float32 frames, `reset()` starts a new episode.
  Frames, float rewards and terminals are a hash of (env id, step counter, action) into a seeded table
  -- the numpy twin of unreal_b200/environment/frame_environment.py:TableFrameProducer, so the device
  path and this oracle see the same frame stream (test infrastructure for SURVEY.md 8f-4)."""
  K_FRAMES = 32
  MOD = 1000003

  def __init__(self, env_id, table):
    self.env_id = int(env_id)
    self.table = table                  # uint8 [K, H, W, 3]
    self.counter = 0
    self.reset()

  def _hash(self, a_plus_1):
    return (self.env_id * 7919 + self.counter * 104729 + a_plus_1 * 613) % self.MOD

  def _frame(self, h):
    return self.table[h % self.K_FRAMES].astype(np.float32) / np.float32(255.0)   # lab_environment.py:99-102

  def reset(self):
    self.counter += 1
    self.last_state = {'image': self._frame(self._hash(0))}
    self.last_action = 0
    self.last_reward = 0

  def process(self, action, flag=1):
    self.counter += 1
    h = self._hash(int(action) + 1)
    image = self._frame(h)
    reward = float(((h // self.K_FRAMES) % 5) - 2) * 0.5 if h % 3 == 0 else 0.0
    terminal = (h % 29 == 0)
    pc = pixel_change(image, self.last_state['image'])
    state = {'image': image}
    self.last_state = state
    self.last_action = action
    self.last_reward = reward
    return state, reward, terminal, pc


class RolloutOracle(object):
  """One reference ``Trainer`` worth of host logic on one maze: warm-up fill,
  on-policy rollout with n-step returns, and the three replay targets.

  ``net`` supplies ``run_base_policy_and_value / run_base_value / run_pc_q_max /
  run_vr_value / reset_state`` (model/model.py:630-728 call surface).  Frames are
  dicts ``{state, pos, reward, action, terminal, pixel_change, last_action,
  last_reward}`` mirroring ``ExperienceFrame`` (experience.py:10-18); ``pos`` is
  the agent cell of ``state`` and is carried only so tests can compare compactly.
  """

  def __init__(self, history_size, random_state, net, n_step_TD=20, local_t_max=20,
               gamma=0.99, gamma_pc=0.9, action_size=4, env=None):
    self.env = MazeOracle() if env is None else env   # any env with the Environment instance contract
    self.ring = RingOracle(history_size, random_state)
    self.random_state = random_state
    self.net = net
    self.n_step_TD = n_step_TD
    self.local_t_max = local_t_max
    self.gamma = gamma
    self.gamma_pc = gamma_pc
    self.action_size = action_size
    self.local_t = 0
    self.episode_reward = 0

  def choose_action(self, pi):
    return self.random_state.choice(len(pi), p=pi)  # trainer.py:147-148

  def _step(self, action):
    env = self.env
    prev_state, prev_pos = env.last_state, (getattr(env, 'x', 0), getattr(env, 'y', 0))
    last_action, last_reward = env.last_action, env.last_reward
    image, reward, terminal, pc = env.process(action)
    frame = dict(state=prev_state, pos=prev_pos, reward=reward, action=action, terminal=terminal,
                 pixel_change=pc, last_action=last_action, last_reward=last_reward)
    self.ring.add(frame)
    return image, frame

  def fill_step(self):
    """trainer.py:176-205."""
    env = self.env
    lar = concat_action_and_reward(env.last_action, self.action_size, env.last_reward)
    pi, _, _ = self.net.run_base_policy_and_value(None, env.last_state, lar)
    action = self.choose_action(pi)
    _, frame = self._step(action)
    if frame['terminal']:
      env.reset()
    if self.ring.is_full():
      env.reset()

  def process_base(self):
    """trainer.py:218-336 -> dict of time-ordered lists."""
    env = self.env
    states, pos, lars, actions, rewards, values = [], [], [], [], [], []
    terminal_end = False
    score = None                       # summary_dict['values']['score_input'] (trainer.py:281-283) -> process()'s episode_score
    image = frame = None
    for _ in range(self.n_step_TD):
      lar = concat_action_and_reward(env.last_action, self.action_size, env.last_reward)
      pi, v, _ = self.net.run_base_policy_and_value(None, env.last_state, lar, "")
      action = self.choose_action(pi)
      states.append(env.last_state); pos.append((getattr(env, 'x', 0), getattr(env, 'y', 0))); lars.append(lar)
      actions.append(action); values.append(v)
      image, frame = self._step(action)
      self.episode_reward += frame['reward']
      rewards.append(frame['reward'])
      self.local_t += 1
      if frame['terminal']:
        terminal_end = True
        score = self.episode_reward
        self.episode_reward = 0
        env.reset()
        self.net.reset_state()
        break
    R = 0.0
    if not terminal_end:
      R = self.net.run_base_value(
          None, image if isinstance(image, dict) else {'image': image},   # the maze returns a bare array (:125)
          concat_action_and_reward(frame['action'], self.action_size, frame['reward']))
    batch_R, batch_adv, batch_a = [], [], []
    for ai, ri, Vi in zip(actions[::-1], rewards[::-1], values[::-1]):
      R = ri + self.gamma * R
      batch_R.append(R); batch_adv.append(R - Vi)
      a = np.zeros([self.action_size]); a[ai] = 1.0
      batch_a.append(a)
    return dict(states=states, pos=pos, lar=lars, a=batch_a[::-1], adv=batch_adv[::-1],
                R=batch_R[::-1], terminal_end=terminal_end, score=score)

  def _lar(self, f):
    return concat_action_and_reward(f['last_action'], self.action_size, f['last_reward'])

  def process_pc(self):
    """trainer.py:339-380."""
    frames = self.ring.sample_sequence(self.local_t_max + 1)[::-1]
    pc_R = np.zeros([PC_CELLS, PC_CELLS], dtype=np.float32)
    if not frames[1]['terminal']:
      pc_R = self.net.run_pc_q_max(None, frames[0]['state'], self._lar(frames[0]))
    out = dict(states=[], pos=[], lar=[], a=[], R=[])
    for f in frames[1:]:
      pc_R = f['pixel_change'] + self.gamma_pc * pc_R
      a = np.zeros([self.action_size]); a[f['action']] = 1.0
      out['states'].append(f['state']); out['pos'].append(f['pos']); out['lar'].append(self._lar(f))
      out['a'].append(a); out['R'].append(pc_R)
    return {k: v[::-1] for k, v in out.items()}

  def process_vr(self):
    """trainer.py:383-412."""
    frames = self.ring.sample_sequence(self.local_t_max + 1)[::-1]
    vr_R = 0.0
    if not frames[1]['terminal']:
      vr_R = self.net.run_vr_value(None, frames[0]['state'], self._lar(frames[0]))
    out = dict(states=[], pos=[], lar=[], R=[])
    for f in frames[1:]:
      vr_R = f['reward'] + self.gamma * vr_R
      out['states'].append(f['state']); out['pos'].append(f['pos']); out['lar'].append(self._lar(f))
      out['R'].append(vr_R)
    return {k: v[::-1] for k, v in out.items()}

  def process_rp(self):
    """trainer.py:415-436."""
    frames = self.ring.sample_rp_sequence()
    return dict(states=[f['state'] for f in frames[:3]], pos=[f['pos'] for f in frames[:3]],
                c=rp_target(frames[3]['reward']))
