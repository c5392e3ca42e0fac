"""End-to-end parity of the batched Trainer against the REFERENCE Trainer's own outputs
(tests/golden/trainer_*.npz): same maze, same Experience logic, same RandomState stream,
network replaced by the deterministic FakeNet on both sides."""
import os

import numpy as np
import pytest
import torch

from oracle import unreal_oracle as O
from fake_net import FakeNet
from batched_fake_net import BatchedFakeNet

pytestmark = pytest.mark.gpu
PC_PROBE = np.array([0, 19, 21, 63, 105, 147, 168, 189, 210, 231, 252, 294, 336, 378, 380, 399])
REL = 1e-5


def _load(golden_dir, name):
  with np.load(os.path.join(golden_dir, name)) as z:
    return {k: z[k] for k in z.files}


def _make_trainer(H, n_step, num_envs, random_state=None, seeds=None, net=None):
  from unreal_b200.train.trainer import Trainer
  from unreal_b200.environment.environment import Environment
  Environment.action_size = -1
  tr = Trainer(1, net, 7.0710678e-4, None, None, 'maze', '', True, True, True, True, 0.05, 1e-3, 20, n_step, 0.99,
               0.9, H, 13200000, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, random_state, 50.0, 0.0, 0.0,
               num_envs=num_envs, seeds=seeds)
  tr.prepare()
  return tr


class _EndAwareNet(BatchedFakeNet):
  """run_base_value is only *called* by the reference when the rollout did not end in a
  terminal (trainer.py:298-300); mirror that so each FakeNet stays in step."""

  def attach(self, trainer):
    self.trainer = trainer

  def run_base_value(self, sess, state, lar):
    tr = self.trainer
    lengths = tr._active.sum(0).cpu().numpy()
    term = tr._term.cpu().numpy()
    v = np.zeros(self.n, np.float32)
    for e, net in enumerate(self.nets):
      ended = term[lengths[e] - 1, e] != 0
      if not ended:
        v[e] = net.run_base_value(None, None, None)
    return torch.from_numpy(v).to(self.device)


def _close(a, b, rel=REL):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)) <= rel


@pytest.mark.parametrize("name", ["h100", "h2000"])
def test_single_env_matches_reference_trainer(golden_dir, name):
  g = _load(golden_dir, "trainer_%s.npz" % name)
  H, n_iter, n_step, seed, net_seed, n_fill = [int(v) for v in g["cfg"]]
  if name == "h2000":
    n_iter = 60                      # keeps the GPU suite short; h100 runs all 400 iterations
  rs = np.random.RandomState(seed)   # the caller's RandomState object itself is advanced
  net = _EndAwareNet([net_seed], "cuda:0")
  tr = _make_trainer(H, n_step, 1, random_state=rs, net=net)
  net.attach(tr)
  fills = 0
  while not tr.experience.is_full():
    assert tr.process(None, 0) == (0, None)
    fills += 1
  assert fills == n_fill
  bi = pi = vi = 0
  for it in range(n_iter):
    diff, _ = tr.process(None, 0)
    f = tr.last_feed
    b = f['base']
    L = int(b['length'][0])
    assert L == g["base_len"][it] and diff == L
    assert np.array_equal(b['pos'][:L, 0].cpu().numpy(), g["base_pos"][bi:bi + L])
    assert np.array_equal(b['last_action_rewards'][:L, 0].cpu().numpy(), g["base_lar"][bi:bi + L].astype(np.float32))
    assert np.array_equal(b['a'][:L, 0].cpu().numpy(), g["base_a"][bi:bi + L].astype(np.float32))
    _close(b['R'][:L, 0].cpu().numpy(), g["base_R"][bi:bi + L])
    _close(b['adv'][:L, 0].cpu().numpy(), g["base_adv"][bi:bi + L])
    # frames are the renders of those cells
    img = b['si'][:L, 0].cpu().numpy()
    for k in range(0, L, 7):
      x, y = g["base_pos"][bi + k]
      assert np.array_equal(img[k], O.maze_render(int(x), int(y), np.float32))
    bi += L
    p = f['pc']
    Lp = int(p['length'][0])
    assert Lp == g["pc_len"][it]
    assert np.array_equal(p['pos'][0, :Lp].cpu().numpy(), g["pc_pos"][pi:pi + Lp])
    assert np.array_equal(p['last_action_reward'][0, :Lp].cpu().numpy(), g["pc_lar"][pi:pi + Lp].astype(np.float32))
    assert np.array_equal(p['a'][0, :Lp].cpu().numpy(), g["pc_a"][pi:pi + Lp].astype(np.float32))
    R = p['R'][0, :Lp].cpu().numpy().astype(np.float64)
    _close(R.reshape(Lp, -1).sum(1) / 400, g["pc_R_sum"][pi:pi + Lp] / 400)
    _close(R.reshape(Lp, -1)[:, PC_PROBE], g["pc_R_probe"][pi:pi + Lp])
    nh = max(0, min(Lp, len(g["pc_R_head"]) - pi))
    if nh:
      _close(R[:nh], g["pc_R_head"][pi:pi + nh])
    pi += Lp
    v = f['vr']
    Lv = int(v['length'][0])
    assert Lv == g["vr_len"][it]
    assert np.array_equal(v['pos'][0, :Lv].cpu().numpy(), g["vr_pos"][vi:vi + Lv])
    assert np.array_equal(v['last_action_reward'][0, :Lv].cpu().numpy(), g["vr_lar"][vi:vi + Lv].astype(np.float32))
    _close(v['R'][0, :Lv].cpu().numpy(), g["vr_R"][vi:vi + Lv])
    vi += Lv
    r = f['rp']
    assert np.array_equal(r['pos'][0].cpu().numpy(), g["rp_pos"][it])
    assert np.array_equal(r['c'][0].cpu().numpy(), g["rp_c"][it].astype(np.float32))
  if name == "h100":
    top = int(tr.experience.ring.state()["top"][0])
    assert (top, tr.local_t) == tuple(g["final_top"])
    # the shared RandomState advanced exactly as the reference's did: next draws agree
    ref = np.random.RandomState(seed)
    w = O.RolloutOracle(H, ref, FakeNet(net_seed), n_step_TD=n_step)
    while not w.ring.is_full():
      w.fill_step()
    for _ in range(n_iter):
      w.process_base(); w.process_pc(); w.process_vr(); w.process_rp()
    assert [rs.randint(0, 1000) for _ in range(5)] == [ref.randint(0, 1000) for _ in range(5)]
  lrs = [tr._anneal_learning_rate(t) for t in (0, 1, 6600000, 13199999, 13200000, 14000000)]
  assert lrs == list(g["lr_anneal"])


def test_many_envs_each_matches_its_own_oracle_worker():
  """16 envs in lock step; every env must equal an independent oracle worker (which is pinned
  to the reference) seeded like that env, including early rollout ends and ring sampling."""
  N, H, T = 16, 60, 20
  seeds = [11 + 3 * e for e in range(N)]
  net_seeds = [500 + e for e in range(N)]
  net = _EndAwareNet(net_seeds, "cuda:0")
  tr = _make_trainer(H, T, N, random_state=np.random.RandomState(0), seeds=seeds, net=net)
  net.attach(tr)
  workers = [O.RolloutOracle(H, np.random.RandomState(s), FakeNet(ns), n_step_TD=T) for s, ns in zip(seeds, net_seeds)]
  while not tr.experience.is_full():
    tr.process(None, 0)
    for w in workers:
      w.fill_step()
  assert all(w.ring.is_full() for w in workers)
  scored = 0
  for it in range(40):
    diff, episode_score = tr.process(None, 0)
    f = tr.last_feed
    steps, scores = 0, []
    for e, w in enumerate(workers):
      b = w.process_base(); p = w.process_pc(); v = w.process_vr(); r = w.process_rp()
      L = len(b['pos'])
      steps += L
      if b['score'] is not None:
        scores.append(b['score'])
      assert int(f['base']['length'][e]) == L
      assert np.array_equal(f['base']['pos'][:L, e].cpu().numpy(), np.array(b['pos']))
      _close(f['base']['R'][:L, e].cpu().numpy(), np.array(b['R'], np.float64))
      _close(f['base']['adv'][:L, e].cpu().numpy(), np.array(b['adv'], np.float64))
      Lp = len(p['pos'])
      assert int(f['pc']['length'][e]) == Lp
      assert np.array_equal(f['pc']['pos'][e, :Lp].cpu().numpy(), np.array(p['pos']))
      _close(f['pc']['R'][e, :Lp].cpu().numpy(), np.stack(p['R']).astype(np.float64))
      Lv = len(v['pos'])
      assert int(f['vr']['length'][e]) == Lv
      _close(f['vr']['R'][e, :Lv].cpu().numpy(), np.array(v['R'], np.float64))
      assert np.array_equal(f['rp']['pos'][e].cpu().numpy(), np.array(r['pos']))
      assert list(f['rp']['c'][e].cpu().numpy()) == r['c']
    # the return contract (trainer.py:451, :481, :635-636; main.py:125 adds the diff to global_t): env steps actually
    # taken by all workers, and the score of the episodes that ended in this window (their mean when several did)
    assert diff == steps
    if scores:
      scored += 1
      assert episode_score is not None and abs(episode_score - float(np.mean(scores))) <= 1e-6
    else:
      assert episode_score is None
