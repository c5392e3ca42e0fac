"""Deterministic stand-in for the network calls made on the hot path.

Shared by ``make_golden.py`` (which drives the REFERENCE's Trainer methods with
it) and by the parity tests (which drive the oracle and the CUDA path with it),
so that both sides see identical policy / value / Q outputs.  It has the call
surface the reference's ``Trainer`` uses on ``self.local_network``
(model/model.py:630-728): ``run_base_policy_and_value``, ``run_base_value``,
``run_pc_q_max``, ``run_vr_value``, ``reset_state``, ``base_lstm_state_out``.
"""
import numpy as np


class FakeNet(object):
  def __init__(self, seed, action_size=4):
    self.rs = np.random.RandomState(seed)
    self.action_size = action_size
    self.base_lstm_state_out = (np.zeros([1, 256], np.float32), np.zeros([1, 256], np.float32))
    self.log = []   # every output in call order, for replaying on the device

  def _pi_v(self):
    logits = self.rs.randn(self.action_size).astype(np.float32) * np.float32(0.7)
    e = np.exp(logits - logits.max())
    pi = (e / e.sum()).astype(np.float32)
    v = np.float32(self.rs.randn() * 0.5)
    return pi, v

  def run_base_policy_and_value(self, sess, s_t, last_action_reward, mode=""):
    pi, v = self._pi_v()
    self.log.append(('pv', pi, v))
    return pi, v, None

  def run_base_value(self, sess, s_t, last_action_reward):
    v = np.float32(self.rs.randn() * 0.5)
    self.log.append(('v', v))
    return v

  def run_pc_q_max(self, sess, s_t, last_action_reward):
    q = self.rs.rand(20, 20).astype(np.float32)
    self.log.append(('q', q))
    return q

  def run_vr_value(self, sess, s_t, last_action_reward):
    v = np.float32(self.rs.randn() * 0.5)
    self.log.append(('vr', v))
    return v

  def reset_state(self):
    self.base_lstm_state_out = (np.zeros([1, 256], np.float32), np.zeros([1, 256], np.float32))
