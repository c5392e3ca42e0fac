"""Adapter that gives the batched Trainer one golden FakeNet per env (test infrastructure)."""
import numpy as np
import torch

from fake_net import FakeNet


class BatchedFakeNet(object):
  def __init__(self, seeds, device, action_size=4):
    self.nets = [FakeNet(int(s), action_size) for s in seeds]
    self.n = len(self.nets)
    self.device = device
    self.a = action_size
    self.base_lstm_state_out = None

  def _mask(self, active):
    return np.ones(self.n, bool) if active is None else active.cpu().numpy().astype(bool)

  def run_base_policy_and_value(self, sess, state, lar, active=None, mode=""):
    m = self._mask(active)
    pi = np.full((self.n, self.a), 1.0 / self.a, np.float32); v = np.zeros(self.n, np.float32)
    for e, net in enumerate(self.nets):
      if m[e]:
        pi[e], v[e], _ = net.run_base_policy_and_value(None, None, None)
    return torch.from_numpy(pi).to(self.device), torch.from_numpy(v).to(self.device), None

  def run_base_value(self, sess, state, lar, need=None):
    raise NotImplementedError  # replaced per test (needs to know which envs ended)

  def run_pc_q_max(self, sess, state, lar):
    return torch.from_numpy(np.stack([net.run_pc_q_max(None, None, None) for net in self.nets])).to(self.device)

  def run_vr_value(self, sess, state, lar):
    return torch.from_numpy(np.array([net.run_vr_value(None, None, None) for net in self.nets], np.float32)).to(self.device)

  def reset_state(self, mask=None):
    pass
