"""Replay buffer classes: drop-in for train/experience.py of the reference.

`ExperienceFrame` and `Experience(history_size, random_state)` keep the reference's exact
surface (the reference's own tests run against them).  `Experience` keeps the frame payloads
by reference on the host, exactly like the reference's deque, but every index decision --
terminal-after-terminal discard, eligibility bookkeeping, sample_sequence / sample_rp_sequence
and the RandomState draws they consume -- is made by the device ring (K5) on a one-env ring,
with the caller's RandomState state moved to the device and back around each sampling call so
that a RandomState shared with other objects (main.py:213, :270) stays in step bit for bit.

`BatchedExperience` is the batched extension used by the batched Trainer: N device rings of
compact frame records (8 B per frame) and N device RandomState streams.
"""
from collections import deque

import numpy as np
import torch

from .. import _lib
from .. import kernels as K


class ExperienceFrame(object):
  """experience.py:10-46."""

  def __init__(self, state, reward, action, terminal, pixel_change, last_action, last_reward):
    self.state = state
    self.action = action
    self.reward = reward
    self.terminal = terminal
    self.pixel_change = pixel_change
    self.last_action = last_action
    self.last_reward = last_reward

  def get_last_action_reward(self, action_size):
    return ExperienceFrame.concat_action_and_reward(self.last_action, action_size, self.last_reward, self.state)

  def get_action_reward(self, action_size):
    return ExperienceFrame.concat_action_and_reward(self.action, action_size, self.reward, self.state)

  @staticmethod
  def concat_action_and_reward(action, action_size, reward, state):
    """one-hot(action) ++ [reward] (++ state['objective'])   (experience.py:34-46)."""
    action_reward = np.zeros([action_size + 1])
    action_reward[action] = 1.0
    action_reward[-1] = float(reward)
    objective = state.get('objective') if hasattr(state, 'get') else None
    if objective is not None:
      return np.concatenate((action_reward, objective))
    return action_reward


def _sign(x):
  return 1 if x > 0 else (-1 if x < 0 else 0)


def pack_record(reward, terminal, action=0, last_action=0, last_reward=0, pos0=(0, 0), pos1=(0, 0)):
  """Host-side twin of frame_pack (csrc/maze_core.cuh).  Rewards are stored as their sign:
  the ring only ever tests `reward > 0` (experience.py:76-80); payloads stay with the caller."""
  r = (pos0[0] & 15) | ((pos0[1] & 15) << 4) | (((pos1[0] & 15) | ((pos1[1] & 15) << 4)) << 8)
  r |= (int(action) & 255) << 16
  r |= (_sign(reward) & 255) << 24
  r |= ((1 if terminal else 0) | 0x80) << 32
  r |= (int(last_action) & 255) << 40
  r |= (_sign(last_reward) & 255) << 48
  return r


class Experience(object):
  """experience.py:48-153, one env, index logic on the device."""

  def __init__(self, history_size, random_state, device='cuda:0'):
    _lib.require_device()
    self._history_size = history_size
    self._frames = deque(maxlen=history_size)
    self.random_state = random_state
    self._device = torch.device(device)
    self._ring = K.ReplayRing(1, history_size, self._device)
    self._streams = K.MtStreams([0], self._device)
    self._rec = torch.zeros(1, dtype=torch.int64, device=self._device)

  def _counts(self):
    st = self._ring.state()
    return int(st["top"][0]), int(st["n_pos"][0]), int(st["n_neg"][0])

  @property
  def _top_frame_index(self):
    return self._counts()[0]

  def get_debug_string(self):
    _, n_pos, n_neg = self._counts()
    return "{} frames, {} zero rewards, {} non zero rewards".format(len(self._frames), n_pos, n_neg)

  def add_frame(self, frame):
    before = int(self._ring.state()["count"][0]), self._top_frame_index
    self._rec[0] = pack_record(frame.reward, frame.terminal, getattr(frame, 'action', 0) or 0)
    self._ring.add(self._rec)
    st = self._ring.state()
    after = int(st["count"][0]), int(st["top"][0])
    if after == before:
      print("Terminal frames continued.")     # experience.py:64-67
      return
    self._frames.append(frame)

  def is_full(self):
    return len(self._frames) >= self._history_size

  def sample_sequence(self, sequence_size):
    self._streams.load_numpy_state(self.random_state)
    start, length, _ = self._ring.sample_sequence(self._streams, sequence_size)
    self._streams.store_numpy_state(self.random_state)
    start, length = int(start[0]), int(length[0])
    if start < 0:
      raise IndexError("sample_sequence on a replay buffer that is not full")
    return [self._frames[start + i] for i in range(length)]

  def sample_rp_sequence(self):
    self._streams.load_numpy_state(self.random_state)
    start, _ = self._ring.sample_rp(self._streams)
    self._streams.store_numpy_state(self.random_state)
    start = int(start[0])
    if start < 0:
      raise IndexError("sample_rp_sequence needs at least four frames")
    return [self._frames[start + i] for i in range(4)]


class BatchedExperience(object):
  """N independent replay rings + N RandomState streams, all device resident."""

  def __init__(self, num_envs, history_size, seeds, device='cuda:0', streams=None):
    _lib.require_device()
    self.num_envs = int(num_envs)
    self.history_size = int(history_size)
    self.device = torch.device(device)
    with torch.cuda.device(self.device):
      self.ring = K.ReplayRing(self.num_envs, self.history_size, self.device)
      self.streams = streams if streams is not None else K.MtStreams(seeds, self.device)

  def add_frames(self, frame_rec):
    with torch.cuda.device(self.device):
      self.ring.add(frame_rec)

  def is_full(self):
    """True when every env's ring is full (the batched Trainer warms all envs in lock step)."""
    return bool(self.ring.state()["full"].all())

  def sample_sequence(self, sequence_size):
    """-> start [N], len [N], dict of unpacked fields [N, L(,2)]."""
    with torch.cuda.device(self.device):
      start, length, rec = self.ring.sample_sequence(self.streams, sequence_size)
      return start, length, K.frame_unpack(rec)

  def sample_rp_sequence(self):
    with torch.cuda.device(self.device):
      start, rec = self.ring.sample_rp(self.streams)
      return start, K.frame_unpack(rec)

  def get_debug_string(self):
    st = self.ring.state()
    return "{} envs: {} frames, {} zero rewards, {} non zero rewards (env 0)".format(
        self.num_envs, int(st["count"][0]), int(st["n_pos"][0]), int(st["n_neg"][0]))


class FramedExperience(BatchedExperience):
  """BatchedExperience for generic-frame envs (lab / gym / indoor / synthetic).

  The record ring keeps the reference's index semantics (experience.py:63-153); what the reference's
  deque holds by reference in each ExperienceFrame (:10-18) lives in payload rings addressed by the
  record's slot: `frames` u8 [N,H,h,w,3] (the state BEFORE the action; 21 168 B per 84x84 frame),
  `pc` f32 [N,H,20,20] (pixel change caused by the action), `scalars` f32 [N,H,2] (reward,
  last_reward -- float rewards, indoor_environment.py:113) and `objective` f32 [N,H,G] when the
  state carries one.  H = 2000 at 1024 envs per GPU is 43 GB of frames + 3.3 GB of maps.
  """

  def __init__(self, num_envs, history_size, seeds, device='cuda:0', streams=None, frame_shape=(84, 84, 3),
               objective_size=0):
    BatchedExperience.__init__(self, num_envs, history_size, seeds, device, streams)
    n, h, d = self.num_envs, self.history_size, self.device
    fh, fw = frame_shape[:2]
    self.frame_shape = tuple(frame_shape)
    self.pc_shape = ((fh - 4) // 4, (fw - 4) // 4)
    self.objective_size = int(objective_size)
    self.frames = torch.zeros(n, h, *self.frame_shape, dtype=torch.uint8, device=d)
    self.pc = torch.zeros(n, h, *self.pc_shape, dtype=torch.float32, device=d)
    self.scalars = torch.zeros(n, h, 2, dtype=torch.float32, device=d)
    self.objective = torch.zeros(n, h, self.objective_size, dtype=torch.float32, device=d) if objective_size else None
    self._slot = torch.zeros(n, dtype=torch.int32, device=d)

  def add_frames(self, frame_rec, frame=None, pixel_change=None, reward=None, last_reward=None, objective=None):
    """add_frame (:63-93) for every env with a valid record; payloads follow the record's slot."""
    with torch.cuda.device(self.device):
      slot = self.ring.add_slots(frame_rec, out=self._slot)
      self.ring.store(self.frames, frame, slot)
      self.ring.store(self.pc, pixel_change, slot)
      self.ring.store(self.scalars, torch.stack((reward, last_reward), dim=1), slot)
      if self.objective is not None:
        self.ring.store(self.objective, objective, slot)

  def _fields(self, rec, start, length, seq_len):
    f = K.frame_unpack(rec, fields=("action", "terminal", "last_action", "valid"))
    sc = self.ring.gather(self.scalars, start, length, seq_len, time_major=False)       # [N,L,2]
    f["reward"] = sc[..., 0].contiguous()
    f["last_reward"] = sc[..., 1].contiguous()
    if self.objective is not None:
      f["objective"] = self.ring.gather(self.objective, start, length, seq_len, time_major=False)
    return f

  def sample_sequence(self, sequence_size):
    """-> start [N], len [N], dict of fields [N,L] (float rewards; frames / maps via gather_*)."""
    with torch.cuda.device(self.device):
      start, length, rec = self.ring.sample_sequence(self.streams, sequence_size)
      return start, length, self._fields(rec, start, length, sequence_size)

  def sample_rp_sequence(self):
    with torch.cuda.device(self.device):
      start, rec = self.ring.sample_rp(self.streams)
      return start, self._fields(rec, start, None, 4)

  def gather_frames(self, start, length, seq_len, time_major=True):
    """The sampled frames' states: u8 [L,N,h,w,3] (zero past `length`)."""
    with torch.cuda.device(self.device):
      return self.ring.gather(self.frames, start, length, seq_len, time_major)

  def gather_pixel_change(self, start, length, seq_len, time_major=True):
    with torch.cuda.device(self.device):
      return self.ring.gather(self.pc, start, length, seq_len, time_major)

  def _payloads(self):
    d = dict(frames=self.frames, pc=self.pc, scalars=self.scalars)
    if self.objective is not None:
      d["objective"] = self.objective
    return d

  def payload_state(self):
    """The payload rings, verbatim (checkpoints)."""
    return dict(self._payloads())

  def load_payload_state(self, st):
    for k, v in self._payloads().items():
      if tuple(st[k].shape) != tuple(v.shape):
        raise _lib.UnrealError("payload %s of the checkpoint is %s, ring has %s" % (k, tuple(st[k].shape), tuple(v.shape)))
      v.copy_(st[k].to(self.device))
