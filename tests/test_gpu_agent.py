"""Full UNREAL agent on the device: batched Trainer + UnrealModel (K7) + RMSPropApplier (K6) driving
maze envs (K1), the replay ring (K5) and the target scans (K3/K4) -- BASELINE config 3 in miniature.

Checks: (1) the feed the model trains on equals the oracle's restatement of the reference's
`_process_*` outputs re-derived from the same sampled records; (2) the loss / gradient of that feed
equals the model oracle's; (3) parameters move and stay finite over several updates."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _agent(n, H=60, seed=0):
  from unreal_b200.environment.environment import Environment
  from unreal_b200.model.model import UnrealModel
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  from unreal_b200.train.trainer import Trainer
  Environment.action_size = -1
  dev = torch.device("cuda", 0)
  net = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0,
                    0.0, num_envs=n, seed=seed)
  applier = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr = Trainer(0, net, 7e-4, None, applier, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, H,
               10 ** 7, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
               num_envs=n, seeds=np.arange(n) + 11)
  tr.prepare()
  return tr, net, applier


def test_agent_trains_end_to_end():
  from oracle import model_oracle as M
  n = 6
  tr, net, applier = _agent(n)
  steps = 0
  while not tr.experience.is_full():
    d, _ = tr.process(None, 0)
    assert d == 0
    steps += 1
    assert steps < 200
  p0 = net.flat.detach().clone()
  d, _ = tr.process(None, 0)
  assert n <= d <= 20 * n          # env steps taken by all n envs in the window (trainer.py:635-636 per worker)
  losses = tr.last_losses
  for k in ("policy", "value", "pc", "vr", "rp", "total", "grad_norm"):
    assert torch.isfinite(losses[k]).all(), k
  assert float((net.flat.detach() - p0).abs().max()) > 0

  # the feed of that update, re-evaluated by the model oracle with the PRE-update parameters
  feed = net.feed_from_trainer(tr.last_feed)
  cpu = {k: {kk: vv.detach().cpu().float() if vv.dtype != torch.bool else vv.cpu().float() for kk, vv in v.items()}
         for k, v in feed.items()}
  names = [o[0] for o in net._offsets]
  params = {name: p0[o:o + cnt].view(shape).cpu().clone() for name, shape, o, cnt in net._offsets}
  oracle = M.ModelOracle(params, 4, 0, 0.05, 0.001, emulate_bf16=True)
  total, parts = oracle.total_loss(cpu)
  for k in ("policy", "value", "pc", "vr", "rp"):
    assert abs(float(losses[k]) - float(parts[k])) <= 2e-3 * max(1.0, abs(float(parts[k]))), k
  # sampled sequences respect the reference's shape rules (trainer.py:343-372): 1..20 target steps
  assert int(tr.last_feed['pc']['length'].min()) >= 1 and int(tr.last_feed['pc']['length'].max()) <= 20
  assert tuple(feed['rp']['images'].shape) == (n, 3, 84, 84, 3)
  assert names[0] == "W_base_conv1"

  for _ in range(3):
    tr.process(None, 0)
  assert torch.isfinite(net.flat).all()
  tr.stop()


def test_checkpoint_restores_the_exact_trajectory(tmp_path):
  """Save mid-training, continue; restore into a FRESH agent and continue: same actions, same
  replay samples (bit-exact integers), same losses up to the split-K atomics' summation order."""
  from unreal_b200.train import checkpoint
  n = 4
  tr, net, ap = _agent(n, H=40, seed=3)
  while not tr.experience.is_full():
    tr.process(None, 0)
  tr.process(None, 0)
  path = str(tmp_path / "agent.pt")
  checkpoint.save(path, tr, global_t=123)

  def run(trainer):
    out = []
    for _ in range(2):
      trainer.process(None, 0)
      f = trainer.last_feed
      out.append(dict(act=f['base']['a'].argmax(-1).cpu(), pc_start=f['pc']['start'].cpu(), vr_start=f['vr']['start'].cpu(),
                      rp_start=f['rp']['start'].cpu(), R=f['base']['R'].cpu(), total=float(trainer.last_losses['total'])))
    return out, trainer.local_network.flat.detach().cpu().clone()

  a, pa = run(tr)
  tr2, net2, ap2 = _agent(n, H=40, seed=99)       # different init: everything must come from the file
  assert checkpoint.load(path, tr2) == 123
  b, pb = run(tr2)
  for x, y in zip(a, b):
    for k in ("act", "pc_start", "vr_start", "rp_start"):
      assert torch.equal(x[k], y[k]), k
    assert torch.allclose(x["R"], y["R"], rtol=1e-4, atol=1e-5)
    assert abs(x["total"] - y["total"]) <= 1e-3 * max(1.0, abs(x["total"]))
  assert torch.allclose(pa, pb, rtol=1e-3, atol=1e-5)
  tr.stop(); tr2.stop()


def test_graphed_data_phase_equals_eager():
  """The CUDA-graph replay of rollout + sampling + targets produces the same feeds as the eager
  path (integers bit-exact), iteration after iteration, including the first (capturing) one."""
  from unreal_b200.train import checkpoint
  n = 5
  tr_e, _, _ = _agent(n, H=40, seed=2)
  tr_g, _, _ = _agent(n, H=40, seed=2)
  tr_g.use_graphs = True
  for tr in (tr_e, tr_g):
    while not tr.experience.is_full():
      tr.process(None, 0)
  for it in range(5):
    # the learner update is captured too (iteration 1) and replayed with the annealed rate (global_t moves)
    de, _ = tr_e.process(None, it * 2000000)
    dg, _ = tr_g.process(None, it * 2000000)
    assert de == dg
    fe, fg = tr_e.last_feed, tr_g.last_feed
    assert torch.equal(fe['base']['a'], fg['base']['a']), it
    assert torch.equal(fe['base']['active'], fg['base']['active'])
    assert torch.equal(fe['base']['pos'], fg['base']['pos'])
    for k in ('pc', 'vr', 'rp'):
      assert torch.equal(fe[k]['start'], fg[k]['start']), (it, k)
    assert torch.allclose(fe['base']['R'], fg['base']['R'], rtol=1e-3, atol=1e-4)
    assert torch.allclose(fe['pc']['R'], fg['pc']['R'], rtol=1e-3, atol=1e-4)
  assert tr_g._ugraph is not None, "the update graph must have been captured"
  assert torch.allclose(tr_e.local_network.flat, tr_g.local_network.flat, rtol=1e-3, atol=1e-5)
  for k in ("total", "grad_norm"):
    assert torch.allclose(tr_e.last_losses[k].float(), tr_g.last_losses[k].float(), rtol=2e-3, atol=1e-4), k
  tr_e.stop(); tr_g.stop()


def test_s2d_observation_agent_equals_f32_agent():
  """obs_s2d=True (frames rendered directly as conv1's bf16 planes, no f32 frame, no s2d pass)
  trains exactly like the f32-frame agent: maze frames are 0/1, so both paths feed conv1 the same
  bf16 values."""
  n = 4
  tr_a, net_a, _ = _agent(n, H=40, seed=6)
  tr_b, net_b, _ = _agent(n, H=40, seed=6)
  tr_b.stop()
  from unreal_b200.train.trainer import Trainer
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  import numpy as np
  ap = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr_b = Trainer(0, net_b, 7e-4, None, ap, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, 40,
                 10 ** 7, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
                 num_envs=n, seeds=np.arange(n) + 11, obs_s2d=True, use_graphs=True)
  tr_b.prepare()
  for tr in (tr_a, tr_b):
    while not tr.experience.is_full():
      tr.process(None, 0)
  for it in range(3):
    tr_a.process(None, 0); tr_b.process(None, 0)
    assert torch.equal(tr_a.last_feed['base']['a'], tr_b.last_feed['base']['a']), it
    assert torch.equal(tr_a.last_feed['pc']['start'], tr_b.last_feed['pc']['start'])
    for k in ("policy", "value", "pc", "vr", "rp"):
      a, b = float(tr_a.last_losses[k]), float(tr_b.last_losses[k])
      assert abs(a - b) <= 1e-3 * max(1.0, abs(a)), (it, k, a, b)
  assert tr_b.last_feed['base']['si'].dtype == torch.bfloat16
  tr_a.stop(); tr_b.stop()


@pytest.mark.parametrize("dedup", [False, True], ids=["dense-encoder", "cell-table"])
def test_cell_observation_agent_equals_frame_agent(dedup):
  """obs_cells=True: observations are the agent cells and conv1 (forward and filter gradient) renders its input
  tiles in shared memory -- no frame in HBM anywhere.  Must train exactly like the agent that materialises
  frames: same actions, same replay samples, same losses, same parameters.  `cell-table`: the encoder + fc1 of all
  samples as a lookup in the 49-cell table with a segment-sum backward (UnrealModel.dedup_cells) -- same forward values
  bit for bit; the gradient is summed per cell in fp32 BEFORE it is rounded to bf16 for the 49-row GEMMs instead of
  after, so parameters agree to the bf16 rounding of a gradient (2^-9) rather than to fp32 summation order."""
  n = 4
  tr_a, net_a, _ = _agent(n, H=40, seed=6)
  _, net_b, _ = _agent(n, H=40, seed=6)
  net_b.dedup_cells = dedup
  net_b.refresh_shadow()
  from unreal_b200.train.trainer import Trainer
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  ap = RMSPropApplier(7e-4, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  tr_b = Trainer(0, net_b, 7e-4, None, ap, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9, 40,
                 10 ** 7, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, np.random.RandomState(1), 50.0, 0.0, 0.0,
                 num_envs=n, seeds=np.arange(n) + 11, obs_cells=True, use_graphs=True)
  tr_b.prepare()
  for tr in (tr_a, tr_b):
    while not tr.experience.is_full():
      tr.process(None, 0)
  for it in range(4):
    tr_a.process(None, 0); tr_b.process(None, 0)
    assert torch.equal(tr_a.last_feed['base']['a'], tr_b.last_feed['base']['a']), it
    assert torch.equal(tr_a.last_feed['pc']['start'], tr_b.last_feed['pc']['start'])
    assert torch.equal(tr_a.last_feed['base']['pos'], tr_b.last_feed['base']['si']), "the cell observations are the positions"
    for k in ("policy", "value", "pc", "vr", "rp"):
      a, b = float(tr_a.last_losses[k]), float(tr_b.last_losses[k])
      assert abs(a - b) <= 1e-3 * max(1.0, abs(a)), (it, k, a, b)
  assert tr_b.last_feed['base']['si'].dtype == torch.int32
  # cell-table: a parameter moves by <= lr * |g| / sqrt(rms + eps) ~ 7e-4 * 5 per update and the two gradients differ by
  # the bf16 rounding of a per-cell sum (2^-8 relative): 4 updates * 3.5e-3 * 4e-3 ~ 6e-5 absolute
  assert torch.allclose(net_a.flat, net_b.flat, rtol=1e-3, atol=1e-4 if dedup else 1e-5)
  tr_a.stop(); tr_b.stop()


def test_rejected_checkpoint_leaves_the_agent_untouched(tmp_path):
  """checkpoint.load validates keys, dtypes and shapes BEFORE it modifies anything (and reads with weights_only=True):
  a checkpoint for another ring size, one with a tampered RMSProp shard, one with an unknown key and one carrying a
  non-tensor object are all refused with the trainer exactly as it was."""
  from unreal_b200 import _lib
  from unreal_b200.train import checkpoint
  n = 4
  tr, net, ap = _agent(n, H=40, seed=3)
  while not tr.experience.is_full():
    tr.process(None, 0)
  tr.process(None, 0)
  good = checkpoint.state_dict(tr, global_t=5)
  tr2, net2, ap2 = _agent(n, H=40, seed=99)
  while not tr2.experience.is_full():
    tr2.process(None, 0)
  tr2.process(None, 0)
  before = (net2.flat.detach().clone(), tr2.streams.mt.clone(), tr2.experience.ring.export_state()["rec"].clone(),
            ap2._rms.clone(), tr2.local_t)

  def untouched():
    return (torch.equal(net2.flat.detach(), before[0]) and torch.equal(tr2.streams.mt, before[1]) and
            torch.equal(tr2.experience.ring.export_state()["rec"], before[2]) and torch.equal(ap2._rms, before[3]) and
            tr2.local_t == before[4])

  bad = []
  d = dict(good); d["rmsprop"] = dict(good["rmsprop"]); d["rmsprop"]["shard_lo"] = 128
  bad.append(d)                                                   # another rank's RMSProp slice
  d = dict(good); d["history"] = 41
  bad.append(d)
  d = dict(good); d["surprise"] = torch.zeros(1)
  bad.append(d)                                                   # unknown key
  d = dict(good); d["rng"] = dict(good["rng"]); d["rng"]["mt"] = good["rng"]["mt"][:, :2].clone()
  bad.append(d)
  d = dict(good); d["flat"] = good["flat"].to(torch.float16)
  bad.append(d)                                                   # unexpected dtype
  d = dict(good); d["rank"] = 1; d["world"] = 2
  bad.append(d)                                                   # written by rank 1 of 2
  for i, d in enumerate(bad):
    with pytest.raises(_lib.UnrealError):
      checkpoint.load_state_dict(tr2, d)
    assert untouched(), i
  path = str(tmp_path / "agent.pt")
  checkpoint.save(path, tr, global_t=5)
  assert checkpoint.rank_path(path, 1, 0) == path and checkpoint.rank_path(path, 8, 3) == path + ".rank3of8"
  assert checkpoint.load(path, tr2) == 5 and not untouched()
  import pickle

  class Evil(object):
    def __reduce__(self):
      return (print, ("code ran on load",))

  evil = str(tmp_path / "evil.pt")
  torch.save({"format": checkpoint.FORMAT, "payload": Evil()}, evil)
  with pytest.raises((pickle.UnpicklingError, RuntimeError, _lib.UnrealError)):
    checkpoint.load(evil, tr2)
  tr.stop(); tr2.stop()


def test_single_env_drop_in_agent_follows_the_references_own_trainer():
  """The drop-in classes on the device (Trainer with num_envs = 1 sharing the caller's RandomState, UnrealModel on the tcgen05
  path, RMSPropApplier) against tests/golden/agent_reference_golden.npz -- the reference's OWN Trainer.process loop (its
  model, Experience, maze and RMSPropApplier, run over the TF-1 op shim of tests/golden/make_agent_golden.py): the fill and
  four learner iterations from the same seeds and initial parameters.  Discrete things must be identical (number of fill
  calls, steps taken, scores, the agent's cell, the position of the RandomState stream: the bf16 policy has to pick the same
  action at every one of the 140 draws).  The parameter CHANGE of every variable since the start must agree with the
  reference's within the bf16 gradient noise (relative L2 over the sampled entries <= 0.25; measured per variable for one
  gradient in test_gpu_model.py: <= 3 % for 18 variables, 12-17 % for conv1)."""
  import os
  from oracle import model_oracle as M
  from unreal_b200.environment.environment import Environment
  from unreal_b200.model.model import UnrealModel
  from unreal_b200.train.rmsprop_applier import RMSPropApplier
  from unreal_b200.train.trainer import Trainer
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "agent_reference_golden.npz"))
  A, G, seed, net_seed, H, n_proc, max_t = (int(x) for x in g["meta"])
  Environment.action_size = -1
  dev = torch.device("cuda", 0)
  net = UnrealModel(A, G, -1, True, True, True, True, 0.05, 0.001, dev, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                    num_envs=1, seed=0)
  init = {k: v.numpy() for k, v in M.init_params(A, G, seed=net_seed).items()}
  net.load_vars(init)
  applier = RMSPropApplier(float(g["initial_lr"]), decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0)
  rs = np.random.RandomState(seed)
  tr = Trainer(1, net, float(g["initial_lr"]), None, applier, 'maze', '', True, True, True, True, 0.05, 0.001, 20, 20, 0.99, 0.9,
               H, max_t, "cuda:0", {'segnet_mode': 0}, (84, 84), True, 0, rs, 50.0, 0.0, 0.0, num_envs=1)
  tr.prepare()

  def check(tag, ret):
    assert tr.local_t == int(g[tag + "_local_t"]), tag
    want = g[tag + "_ret"]
    assert ret[0] == int(want[0]) and ((ret[1] is None) == bool(np.isnan(want[1]))), (tag, ret, want)
    assert tuple(int(v) for v in tr.environment.state.pos[0].cpu()) == tuple(int(v) for v in g[tag + "_pos"]), tag
    probe = np.random.RandomState(); probe.set_state(rs.get_state())
    assert int(probe.randint(0, 2 ** 31 - 1)) == int(g[tag + "_next_draw"]), tag + ": the RandomState streams have diverged"
    if tag == "fill":
      return
    for name, v in net.named_vars().items():
      idx = g["idx_" + name]
      d_ref = g[tag + "_val_" + name] - init[name].reshape(-1)[idx].astype(np.float64)
      d_got = v.detach().cpu().double().reshape(-1).numpy()[idx] - init[name].reshape(-1)[idx].astype(np.float64)
      rel = np.linalg.norm(d_got - d_ref) / max(np.linalg.norm(d_ref), 1e-12)
      assert rel <= 0.25, (tag, name, rel)

  fills = 0
  while not tr.experience.is_full():
    assert tr.process(None, 0) == (0, None)
    fills += 1
  assert fills == int(g["n_fill"])
  check("fill", (0, None))
  global_t = 0
  for it in range(n_proc):
    ret = tr.process(None, global_t)
    global_t += ret[0]
    check("it%d" % it, ret)
  tr.stop()
