"""Framed replay ring + generic-frame env adapter + FrameTrainer against the oracle:
the numpy restatement of Experience / Trainer._process_* (pinned to the reference by
tests/test_oracle_golden.py) driven with a table-hashed frame stream that the device producer
reproduces bit for bit.  Indices, frames, records: bit-exact.  Pixel change / targets: 1e-5 relative
(the north star's fp32 tolerance); the u8 K2 path is additionally expected to be bit-equal to the
reference's float32 evaluation."""
from collections import deque

import numpy as np
import pytest
import torch

from oracle import unreal_oracle as O
from fake_net import FakeNet
from batched_fake_net import BatchedFakeNet

pytestmark = pytest.mark.gpu
REL = 1e-5
DEV = "cuda:0"


def _close(a, b, rel=REL):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  assert a.shape == b.shape
  assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0), initial=0.0) <= rel


def _u8(image_f32):
  return np.rint(np.asarray(image_f32, np.float64) * 255.0).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# payload store / gather against a python deque
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("item_shape,dtype", [((84, 84, 3), torch.uint8), ((20, 20), torch.float32),
                                              ((2,), torch.float32), ((5,), torch.float32), ((36, 4), torch.uint8)])
def test_ring_store_and_gather_follow_the_deque(item_shape, dtype):
  from unreal_b200 import kernels as K
  from unreal_b200.train.experience import pack_record
  N, H, L = 5, 23, 6
  dev = torch.device(DEV)
  ring = K.ReplayRing(N, H, dev)
  payload = torch.zeros(N, H, *item_shape, dtype=dtype, device=dev)
  rs = np.random.RandomState(3)
  model = [deque(maxlen=H) for _ in range(N)]
  last_term = [False] * N
  for step in range(3 * H + 7):
    term = rs.rand(N) < 0.15
    valid = rs.rand(N) < 0.9            # inactive envs hand in an invalid record
    if dtype == torch.uint8:
      src = rs.randint(0, 256, size=(N,) + item_shape).astype(np.uint8)
    else:
      src = rs.randn(*((N,) + item_shape)).astype(np.float32)
    rec = np.array([pack_record(float(rs.randint(-1, 2)), bool(term[e])) if valid[e] else 0 for e in range(N)], np.int64)
    slot = ring.add_slots(torch.from_numpy(rec).to(dev))
    ring.store(payload, torch.from_numpy(src).to(dev), slot)
    slot = slot.cpu().numpy()
    for e in range(N):
      accepted = valid[e] and not (term[e] and last_term[e] and len(model[e]) > 0)
      assert (slot[e] >= 0) == accepted
      if accepted:
        model[e].append(src[e].copy())
        last_term[e] = bool(term[e])
  st = ring.state()
  assert st["full"].all()
  start = rs.randint(0, H - L, size=N).astype(np.int32)
  length = rs.randint(0, L + 1, size=N).astype(np.int32)
  start[1] = -1                         # "not sampled" -> zeros
  for time_major in (True, False):
    out = ring.gather(payload, torch.from_numpy(start).to(dev), torch.from_numpy(length).to(dev), L, time_major)
    out = out.cpu().numpy()
    full = ring.gather(payload, torch.from_numpy(start).to(dev), None, L, time_major).cpu().numpy()
    for e in range(N):
      for t in range(L):
        got = out[t, e] if time_major else out[e, t]
        got_full = full[t, e] if time_major else full[e, t]
        want_full = model[e][start[e] + t] if start[e] >= 0 else np.zeros(item_shape, src.dtype)
        want = want_full if t < length[e] else np.zeros(item_shape, src.dtype)
        assert np.array_equal(got, want), (e, t)
        assert np.array_equal(got_full, want_full), (e, t)
  ring.close()


def test_ring_store_rejects_mismatched_payload():
  from unreal_b200 import _lib, kernels as K
  ring = K.ReplayRing(2, 8, torch.device(DEV))
  slot = torch.zeros(2, dtype=torch.int32, device=DEV)
  with pytest.raises(_lib.UnrealError):
    ring.store(torch.zeros(2, 9, 4, device=DEV), torch.zeros(2, 4, device=DEV), slot)
  with pytest.raises(_lib.UnrealError):       # 3-byte items are not a multiple of 4
    ring.store(torch.zeros(2, 8, 3, dtype=torch.uint8, device=DEV), torch.zeros(2, 3, dtype=torch.uint8, device=DEV), slot)
  ring.close()


# ---------------------------------------------------------------------------------------------
# env adapter against the oracle env
# ---------------------------------------------------------------------------------------------
def _make_env(n, seed=5, producer='table'):
  from unreal_b200.environment.environment import Environment
  Environment.action_size = -1
  return Environment.create_environment('synthetic', '', env_args={'num_envs': n, 'device': DEV, 'producer': producer,
                                                                   'seed': seed})


def test_frame_env_matches_oracle_env_step_by_step():
  from unreal_b200 import kernels as K
  from unreal_b200.environment.frame_environment import TableFrameProducer
  N = 6
  env = _make_env(N)
  table = TableFrameProducer.make_table(5)
  assert np.array_equal(table, O.make_frame_table(5)), "device producer and oracle env hash into the same frame table"
  oracles = [O.TableFrameEnvOracle(e, table) for e in range(N)]
  rs = np.random.RandomState(0)
  exact = True
  for e, o in enumerate(oracles):
    assert np.array_equal(env.last_state['image'][e].cpu().numpy(), _u8(o.last_state['image']))
  for step in range(120):
    act = rs.randint(0, 3, size=N).astype(np.int32)
    prev_la = env.last_action.cpu().numpy().copy(); prev_lr = env.last_reward.cpu().numpy().copy()
    state, reward, terminal, pc = env.process(torch.from_numpy(act).to(DEV))
    rec = K.frame_unpack(env.frame_rec)
    reward = reward.cpu().numpy(); terminal = terminal.cpu().numpy(); pc = pc.cpu().numpy()
    img = state['image'].cpu().numpy()
    for e, o in enumerate(oracles):
      la, lr = o.last_action, o.last_reward
      s, r, t, p = o.process(int(act[e]))
      assert reward[e] == np.float32(r) and bool(terminal[e]) == t
      _close(pc[e], p)
      exact &= np.array_equal(pc[e], p.astype(np.float32))
      assert int(rec["action"][e]) == act[e] and bool(rec["terminal"][e]) == t
      assert int(rec["reward"][e]) == np.sign(r) and int(rec["last_action"][e]) == la == prev_la[e]
      assert int(rec["last_reward"][e]) == np.sign(lr) and np.float32(lr) == prev_lr[e]
      if t:
        o.reset()             # trainer.py:201-202; the adapter has already done it
      assert np.array_equal(img[e], _u8(o.last_state['image']))
      assert int(env.last_action[e]) == o.last_action and float(env.last_reward[e]) == np.float32(o.last_reward)
  assert exact, "u8 K2 is expected to be bit-equal to the reference's float32 pixel change"


def test_host_env_producer_equals_device_producer():
  """Reference-style host env objects behind HostEnvProducer give the same device stream."""
  from unreal_b200.environment import frame_environment as FE
  N = 4
  table = FE.TableFrameProducer.make_table(9)
  host_envs = [O.TableFrameEnvOracle(e, table) for e in range(N)]
  # the adapter's constructor resets once more: line the counters up with the device producer
  dev_env = FE.BatchedFrameEnvironment(FE.TableFrameProducer(N, DEV, 3, seed=9), DEV)
  for o in host_envs:
    o.counter = 0
  host_env = FE.BatchedFrameEnvironment(FE.HostEnvProducer(host_envs, DEV, 3), DEV)
  rs = np.random.RandomState(1)
  assert torch.equal(dev_env.last_state['image'], host_env.last_state['image'])
  for step in range(60):
    act = torch.from_numpy(rs.randint(0, 3, size=N).astype(np.int32)).to(DEV)
    active = torch.from_numpy((rs.rand(N) < 0.8).astype(np.uint8)).to(DEV)
    a = dev_env.process(act, active=active)
    b = host_env.process(act, active=active)
    m = active.bool()
    assert torch.equal(a[0]['image'], b[0]['image'])
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3][m], b[3][m])
    assert torch.equal(dev_env.frame_rec, host_env.frame_rec)
    assert torch.equal(dev_env.last_action, host_env.last_action) and torch.equal(dev_env.last_reward, host_env.last_reward)


# ---------------------------------------------------------------------------------------------
# FrameTrainer against one oracle worker per env
# ---------------------------------------------------------------------------------------------
class _EndAwareNet(BatchedFakeNet):
  def attach(self, trainer):
    self.trainer = trainer

  def run_base_value(self, sess, state, lar):
    tr = self.trainer
    lengths = tr._active.sum(0).cpu().numpy()
    term = tr._term.cpu().numpy()
    v = np.zeros(self.n, np.float32)
    for e, net in enumerate(self.nets):
      if not term[lengths[e] - 1, e]:
        v[e] = net.run_base_value(None, None, None)
    return torch.from_numpy(v).to(self.device)


def _make_trainer(H, T, N, seeds, net, env_seed=5):
  from unreal_b200.train.trainer import Trainer
  from unreal_b200.train.frame_trainer import FrameTrainer
  from unreal_b200.environment.environment import Environment
  Environment.action_size = -1
  tr = Trainer(1, net, 7.0710678e-4, None, None, 'synthetic', '', True, True, True, True, 0.05, 1e-3, 20, T, 0.99,
               0.9, H, 13200000, DEV, {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(0), 50.0, 0.0, 0.0,
               num_envs=N, seeds=seeds, env_args={'producer': 'table', 'seed': env_seed})
  assert isinstance(tr, FrameTrainer)
  tr.prepare()
  return tr


def test_frame_trainer_matches_one_oracle_worker_per_env():
  from unreal_b200.environment.frame_environment import TableFrameProducer
  N, H, T = 6, 60, 20
  seeds = [11 + 3 * e for e in range(N)]
  net_seeds = [500 + e for e in range(N)]
  net = _EndAwareNet(net_seeds, DEV, action_size=3)
  tr = _make_trainer(H, T, N, seeds, net)
  net.attach(tr)
  table = TableFrameProducer.make_table(5 + 1)      # env seed + thread_index, like the synthetic env derives it
  workers = [O.RolloutOracle(H, np.random.RandomState(s), FakeNet(ns, 3), n_step_TD=T, action_size=3,
                             env=O.TableFrameEnvOracle(e, table))
             for e, (s, ns) in enumerate(zip(seeds, net_seeds))]
  while not tr.experience.is_full():
    tr.process(None, 0)
    for w in workers:
      w.fill_step()
  assert all(w.ring.is_full() for w in workers)
  ends = 0
  for it in range(25):
    tr.process(None, 0)
    f = tr.last_feed
    base_si = f['base']['si'].cpu().numpy()
    pc_img = f['pc']['images'].cpu().numpy(); vr_img = f['vr']['images'].cpu().numpy()
    rp_img = f['rp']['images'].cpu().numpy()
    for e, w in enumerate(workers):
      b = w.process_base(); p = w.process_pc(); v = w.process_vr(); r = w.process_rp()
      ends += int(b['terminal_end'])
      L = len(b['states'])
      assert int(f['base']['length'][e]) == L
      assert np.array_equal(base_si[:L, e], np.stack([_u8(s['image']) for s in b['states']]))
      _close(f['base']['last_action_rewards'][:L, e].cpu().numpy(), np.stack(b['lar']))
      assert np.array_equal(f['base']['a'][:L, e].cpu().numpy(), np.stack(b['a']))
      _close(f['base']['R'][:L, e].cpu().numpy(), np.array(b['R'], np.float64))
      _close(f['base']['adv'][:L, e].cpu().numpy(), np.array(b['adv'], np.float64))
      Lp = len(p['states'])
      assert int(f['pc']['length'][e]) == Lp
      assert np.array_equal(pc_img[:Lp, e], np.stack([_u8(s['image']) for s in p['states']]))
      assert not pc_img[Lp + 1:, e].any()       # [Lp] is the bootstrap frame (masked by `length`), zeros beyond
      _close(f['pc']['last_action_reward'][e, :Lp].cpu().numpy(), np.stack(p['lar']))
      assert np.array_equal(f['pc']['a'][e, :Lp].cpu().numpy(), np.stack(p['a']))
      _close(f['pc']['R'][e, :Lp].cpu().numpy(), np.stack(p['R']).astype(np.float64))
      Lv = len(v['states'])
      assert int(f['vr']['length'][e]) == Lv
      assert np.array_equal(vr_img[:Lv, e], np.stack([_u8(s['image']) for s in v['states']]))
      _close(f['vr']['last_action_reward'][e, :Lv].cpu().numpy(), np.stack(v['lar']))
      _close(f['vr']['R'][e, :Lv].cpu().numpy(), np.array(v['R'], np.float64))
      assert np.array_equal(rp_img[e], np.stack([_u8(s['image']) for s in r['states']]))
      assert list(f['rp']['c'][e].cpu().numpy()) == r['c']
  assert ends > 0, "the test must cover rollouts that end in a terminal"


def test_frame_trainer_runs_the_real_network():
  """FrameTrainer + UnrealModel + RMSProp on synthetic indoor-shaped frames: finite losses, parameters move."""
  from unreal_b200.environment.environment import Environment
  from unreal_b200.model.model import UnrealModel
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  from unreal_b200.train.trainer import Trainer
  Environment.action_size = -1
  N = 8
  net = UnrealModel(3, 0, -1, True, True, True, True, 0.05, 0.001, DEV, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                    num_envs=N, seed=0)
  applier = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr = Trainer(0, net, 7e-4, None, applier, 'synthetic', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, 40,
               10 ** 8, DEV, {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
               num_envs=N, seeds=np.arange(N) + 3, env_args={'producer': 'table', 'seed': 2})
  tr.prepare()
  while not tr.experience.is_full():
    tr.process(None, 0)
  before = net.flat.detach().clone()
  for _ in range(2):
    steps, _ = tr.process(None, 0)
    assert N <= steps <= 20 * N          # env steps of all N envs in the window
  losses = {k: float(v) for k, v in tr.last_losses.items()}
  assert all(np.isfinite(list(losses.values()))), losses
  assert float((net.flat.detach() - before).abs().max()) > 0


def _real_agent(n, H=40, seed=0, graphs=False):
  from unreal_b200.environment.environment import Environment
  from unreal_b200.model.model import UnrealModel
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  from unreal_b200.train.trainer import Trainer
  Environment.action_size = -1
  net = UnrealModel(3, 0, -1, True, True, True, True, 0.05, 0.001, DEV, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                    num_envs=n, seed=seed)
  applier = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr = Trainer(0, net, 7e-4, None, applier, 'synthetic', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, H,
               10 ** 8, DEV, {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
               num_envs=n, seeds=np.arange(n) + 3, env_args={'producer': 'table', 'seed': 2}, use_graphs=graphs)
  tr.prepare()
  return tr


def test_frame_trainer_graphed_data_phase_equals_eager():
  """Rollout through the env adapter + framed sampling + gathers + targets replayed as ONE CUDA graph give
  the same feeds as the eager path (integers and frames bit-exact), including the capturing iteration."""
  n = 5
  tr_e, tr_g = _real_agent(n, seed=2), _real_agent(n, seed=2, graphs=True)
  for tr in (tr_e, tr_g):
    while not tr.experience.is_full():
      tr.process(None, 0)
  for it in range(4):
    de, _ = tr_e.process(None, 0)
    dg, _ = tr_g.process(None, 0)
    assert de == dg
    fe, fg = tr_e.last_feed, tr_g.last_feed
    assert torch.equal(fe['base']['a'], fg['base']['a']), it
    assert torch.equal(fe['base']['active'], fg['base']['active'])
    assert torch.equal(fe['base']['si'], fg['base']['si'])
    for k in ('pc', 'vr', 'rp'):
      assert torch.equal(fe[k]['start'], fg[k]['start']), (it, k)
      assert torch.equal(fe[k]['images'], fg[k]['images']), (it, k)
    assert torch.allclose(fe['base']['R'], fg['base']['R'], rtol=1e-3, atol=1e-4)
    assert torch.allclose(fe['pc']['R'], fg['pc']['R'], rtol=1e-3, atol=1e-4)
  assert torch.allclose(tr_e.local_network.flat, tr_g.local_network.flat, rtol=1e-3, atol=1e-5)


def test_frame_trainer_checkpoint_restores_the_exact_trajectory(tmp_path):
  from unreal_b200.train import checkpoint
  n = 4
  tr = _real_agent(n, seed=3)
  while not tr.experience.is_full():
    tr.process(None, 0)
  tr.process(None, 0)
  path = str(tmp_path / "framed_agent.pt")
  checkpoint.save(path, tr, global_t=77)

  def run(trainer):
    out = []
    for _ in range(2):
      trainer.process(None, 0)
      f = trainer.last_feed
      out.append(dict(act=f['base']['a'].argmax(-1).cpu(), si=f['base']['si'].cpu().clone(),
                      pc_start=f['pc']['start'].cpu(), pc_img=f['pc']['images'].cpu().clone(),
                      rp_start=f['rp']['start'].cpu(), total=float(trainer.last_losses['total'])))
    return out

  a = run(tr)
  tr2 = _real_agent(n, seed=99)
  assert checkpoint.load(path, tr2) == 77
  b = run(tr2)
  for x, y in zip(a, b):
    for k in ("act", "si", "pc_start", "pc_img", "rp_start"):
      assert torch.equal(x[k], y[k]), k
    assert abs(x["total"] - y["total"]) <= 1e-3 * max(1.0, abs(x["total"]))


def test_new_entry_points_reject_bad_arguments():
  """The C ABI returns UNREAL_EINVAL with a message (never crashes) for null / misaligned / out-of-range arguments of
  the framed-ring and render-fused entry points."""
  from unreal_b200 import _lib, kernels as K
  dev = torch.device(DEV)
  ring = K.ReplayRing(2, 8, dev)
  payload = torch.zeros(2, 8, 4, dtype=torch.float32, device=dev)
  start = torch.zeros(2, dtype=torch.int32, device=dev)
  with pytest.raises(_lib.UnrealError, match="seq_len"):
    ring.gather(payload, start, None, 9)                              # longer than the ring
  with pytest.raises(_lib.UnrealError):
    ring.gather(torch.zeros(3, 8, 4, device=dev), start, None, 2)     # payload of another ring shape
  sp = _lib.stream_ptr()
  assert _lib.lib.unreal_replay_gather(ring._h, None, 16, start.data_ptr(), None, 2, 1, payload.data_ptr(), sp) != 0
  assert "null" in _lib.last_error()
  assert _lib.lib.unreal_ring_store(payload.data_ptr(), payload.data_ptr(), start.data_ptr(), 2, 8, 6, sp) != 0
  assert "multiple of 4" in _lib.last_error()
  assert _lib.lib.unreal_frame_pack(None, None, None, None, None, None, None, 4, sp) != 0
  pos = torch.zeros(5, 2, dtype=torch.int32, device=dev)
  w = torch.zeros(4, 6, 16, 8, dtype=torch.bfloat16, device=dev)
  out = torch.zeros(4, 20, 20, 16, dtype=torch.bfloat16, device=dev)
  assert _lib.lib.unreal_conv1_fwd_maze(pos.data_ptr() + 4, w.data_ptr(), None, out.data_ptr(), 4, sp) != 0   # pos not 8-byte aligned
  assert "align" in _lib.last_error()
  assert _lib.lib.unreal_conv1_fwd_maze(pos.data_ptr(), w.data_ptr(), None, out.data_ptr(), 0, sp) != 0
  with pytest.raises(_lib.UnrealError, match="21-pixel"):
    K.conv1_wgrad_maze(pos, torch.zeros(2, 5 * 400, 8, dtype=torch.bfloat16, device=dev))
  h = torch.zeros(4, 256, device=dev)
  assert _lib.lib.unreal_a3c_head_loss(h.data_ptr(), None, None, None, None, None, None, None, None, 4, 9, 0.0, 0.25,
                                       None, None, None, None, None, sp) != 0
  assert "action count" in _lib.last_error()
  torch.cuda.synchronize()
  ring.close()


def test_registered_host_simulators_drive_the_trainer():
  """INTEGRATION.md B2 end to end: reference-style host env objects registered as an env type, Trainer(...) builds the
  framed trainer over them and produces the same feeds as one oracle worker per env."""
  from unreal_b200.environment import frame_environment as FE
  from unreal_b200.environment.environment import Environment
  from unreal_b200.train.trainer import Trainer
  N, H, T = 3, 40, 20
  table = O.make_frame_table(12)
  sims = [O.TableFrameEnvOracle(e, table) for e in range(N)]
  for sim in sims:
    sim.counter = 0          # the adapter's constructor resets once, like a worker's env constructor does
  Environment.register('hostsim', lambda name, args, tt, ti: FE.BatchedFrameEnvironment(
      FE.HostEnvProducer(sims, args['device'], 3), args['device']), action_size=3)
  try:
    Environment.action_size = -1
    seeds = [7 + e for e in range(N)]
    net_seeds = [90 + e for e in range(N)]
    net = _EndAwareNet(net_seeds, DEV, action_size=3)
    tr = Trainer(1, net, 7e-4, None, None, 'hostsim', '', True, True, True, True, 0.05, 1e-3, 20, T, 0.99, 0.9, H, 10 ** 7,
                 DEV, {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(0), 50.0, 0.0, 0.0, num_envs=N, seeds=seeds)
    net.attach(tr)
    tr.prepare()
    workers = [O.RolloutOracle(H, np.random.RandomState(s), FakeNet(ns, 3), n_step_TD=T, action_size=3,
                               env=O.TableFrameEnvOracle(e, table)) for e, (s, ns) in enumerate(zip(seeds, net_seeds))]
    while not tr.experience.is_full():
      tr.process(None, 0)
      for w in workers:
        w.fill_step()
    for it in range(4):
      tr.process(None, 0)
      f = tr.last_feed
      for e, w in enumerate(workers):
        b = w.process_base(); p = w.process_pc(); w.process_vr(); r = w.process_rp()
        L = len(b['states'])
        assert int(f['base']['length'][e]) == L
        assert np.array_equal(f['base']['si'][:L, e].cpu().numpy(), np.stack([_u8(s['image']) for s in b['states']]))
        _close(f['base']['R'][:L, e].cpu().numpy(), np.array(b['R'], np.float64))
        Lp = len(p['states'])
        assert int(f['pc']['length'][e]) == Lp
        _close(f['pc']['R'][e, :Lp].cpu().numpy(), np.stack(p['R']).astype(np.float64))
        assert list(f['rp']['c'][e].cpu().numpy()) == r['c']
  finally:
    Environment._registry.pop('hostsim', None)
    Environment.action_size = -1


def test_rollout_lar_kernel_matches_reference_vectors(golden_dir):
  """unreal_rollout_lar (one-hot(last_action) ++ [last_reward] ++ objective) against the REFERENCE's
  ExperienceFrame.concat_action_and_reward outputs (tests/golden/lar_golden.npz)."""
  import os
  from unreal_b200 import kernels as K
  with np.load(os.path.join(golden_dir, "lar_golden.npz")) as z:
    g = {k: z[k] for k in z.files}
  act = torch.from_numpy(g["action"].astype(np.int32)).to(DEV)
  rew = torch.from_numpy(g["reward"].astype(np.float32)).to(DEV)
  obj = torch.from_numpy(g["objective"]).to(DEV)
  assert np.array_equal(K.rollout_lar(act, rew, 3, objective=obj).cpu().numpy(), g["with_obj"].astype(np.float32))
  assert np.array_equal(K.rollout_lar(act, rew, 3).cpu().numpy(), g["without"].astype(np.float32))
