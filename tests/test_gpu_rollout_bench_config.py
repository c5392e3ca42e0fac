"""Parity of the BENCHED configuration itself (bench.py, BASELINE.json configs[1]): RolloutTargets(4096 envs, 20 steps,
use_graphs=True) through `run_device` and through the host-buffer entry point `run_host`, f32 and u8 frames, against
the oracle (pinned to the reference by tests/test_oracle_golden.py) on every output of the pass:

    rewards, terminals, frame records, pixel-change maps, ALL T x N frames, R, adv, PC targets

Two passes per case, so the state carried from one CUDA-graph replay into the next (positions, last action / reward,
auto-reset at episode ends) is covered as well.  The oracle side: `O.maze_step` for every env and step (the reference's
integer logic, maze_environment.py:76-122), `O.maze_render` / `O.pixel_change` evaluated literally once per distinct
cell / (cell, next cell) pair and looked up, `O.nstep_returns_segmented`, and the PC recurrence of trainer.py:352-372.
"""
import numpy as np
import pytest
import torch

from oracle import unreal_oracle as O

pytestmark = pytest.mark.gpu
N, T = 4096, 20
GAMMA, GAMMA_PC = 0.99, 0.9


class _OracleBatch(object):
  """N independent reference mazes driven in lock step (each one `MazeOracle` + the caller-side reset)."""

  def __init__(self, n):
    self.n = n
    self.pos = [O.START] * n
    self._frame = {}
    self._pc = {}

  def frame(self, cell):
    if cell not in self._frame:
      self._frame[cell] = O.maze_render(cell[0], cell[1])
    return self._frame[cell]

  def pc(self, c0, c1):
    """Environment._calc_pixel_change (environment.py:93-99), literally, on the two rendered float64 frames."""
    if (c0, c1) not in self._pc:
      self._pc[(c0, c1)] = O.pixel_change(self.frame(c1), self.frame(c0)).astype(np.float32)
    return self._pc[(c0, c1)]

  def run(self, actions):
    t_len, n = actions.shape
    rew = np.zeros((t_len, n), np.float32); term = np.zeros((t_len, n), np.uint8)
    cell_after = np.zeros((t_len, n, 2), np.int64)       # the cell the frame written at step t shows
    pcs = np.zeros((t_len, n, 20, 20), np.float32)
    for t in range(t_len):
      for e in range(n):
        x, y = self.pos[e]
        nx, ny, r, tm = O.maze_step(x, y, int(actions[t, e]))
        rew[t, e] = r; term[t, e] = tm
        pcs[t, e] = self.pc((x, y), (nx, ny))
        self.pos[e] = O.START if tm else (nx, ny)         # trainer.py:201-202 / :292: reset at the episode end
        cell_after[t, e] = self.pos[e]
    return rew, term, pcs, cell_after


def _frame_table(dtype, dev):
  """All 49 cells' oracle frames on the device, indexed y*7+x (wall cells never occur)."""
  tab = np.zeros((49, 84, 84, 3), np.float32)
  for y in range(7):
    for x in range(7):
      if not O.WALLS[y, x]:
        tab[y * 7 + x] = O.maze_render(x, y, np.float32)
  t = torch.from_numpy(tab).to(dev)
  return (t * 255).to(torch.uint8) if dtype == torch.uint8 else t


def _pc_targets_oracle(pcs, term, boot_q):
  """trainer.py:352-372 per window, segmented at the episode ends like the kernel: fp32, the reference's order."""
  tgt = np.zeros_like(pcs)
  acc = boot_q.astype(np.float32).copy()
  g = np.float32(GAMMA_PC)
  for t in range(pcs.shape[0] - 1, -1, -1):
    acc = np.where(term[t][:, None, None] != 0, np.float32(0), acc)
    acc = pcs[t] + g * acc
    tgt[t] = acc
  return tgt


def _check_pass(eng, out, oracle, actions, values, boot, boot_q, table, host):
  dev = eng.device
  rew, term, pcs, cells = oracle.run(actions)
  g_rew = out["reward"] if host else eng.reward.cpu().numpy()
  g_term = out["terminal"] if host else eng.terminal.cpu().numpy()
  assert np.array_equal(g_rew, rew), "rewards differ"
  assert np.array_equal(g_term, term), "terminals differ"
  assert np.array_equal(eng.pc.cpu().numpy(), pcs), "pixel-change maps differ"
  # every one of the T x N frames, compared on the device against the oracle's render of that cell
  idx = torch.from_numpy(cells[..., 1] * 7 + cells[..., 0]).to(dev)
  for t in range(T):
    assert torch.equal(eng.obs[t], table[idx[t]]), "frames of step %d differ" % t
  R, adv = O.nstep_returns_segmented(rew, values, term, boot, GAMMA, np.float32)
  g_R = out["R"] if host else eng.R.cpu().numpy()
  g_adv = out["adv"] if host else eng.adv.cpu().numpy()
  assert np.array_equal(g_R, R) and np.array_equal(g_adv, adv), "n-step returns / advantages differ"
  # and within the north star's 1e-5 of the float64 evaluation (the reference era's promotion)
  R64, adv64 = O.nstep_returns_segmented(rew, values, term, boot, GAMMA, np.float64)
  assert np.max(np.abs(g_R - R64) / np.maximum(np.abs(R64), 1.0)) <= 1e-5
  tgt = eng.pc_tgt.cpu().numpy()
  want = _pc_targets_oracle(pcs, term, boot_q)
  assert np.array_equal(tgt, want), "PC targets differ"
  # frame records: the packed ExperienceFrame fields of the step
  from unreal_b200 import kernels as K
  f = K.frame_unpack(eng.frame_rec)
  assert np.array_equal(f["action"].cpu().numpy(), actions)
  assert np.array_equal(f["reward"].cpu().numpy(), rew) and np.array_equal(f["terminal"].cpu().numpy(), term)
  return rew, term


@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.uint8], ids=["f32", "u8"])
@pytest.mark.parametrize("host", [False, True], ids=["run_device", "run_host"])
@pytest.mark.parametrize("window", [False, True], ids=["k1-per-step", "k1-window"])
def test_benched_rollout_targets_match_the_oracle(obs_dtype, host, window):
  """`window`: K1 as T per-step launches, or as ONE launch over the T steps (unreal_maze_window: every work item
  re-simulates its env's integer steps from the window's start state) -- bench.py's default for u8 frames."""
  from unreal_b200.train.rollout import RolloutTargets
  dev = torch.device("cuda", 0)
  eng = RolloutTargets(N, T, GAMMA, GAMMA_PC, obs_dtype, dev, auto_reset=True, use_graphs=True, window_kernel=window)
  table = _frame_table(obs_dtype, dev)
  oracle = _OracleBatch(N)
  rs = np.random.RandomState(42 + int(host))
  ends = 0
  for p in range(2):
    # pass 1 walks towards the goal half of the time so that episodes end (and auto-reset) inside the window
    actions = rs.randint(0, 4, size=(T, N)).astype(np.int32)
    values = rs.randn(T, N).astype(np.float32)
    boot = rs.randn(N).astype(np.float32)
    boot_q = rs.rand(N, 20, 20).astype(np.float32)
    if host:
      out = eng.run_host(actions, values, boot, boot_q)
      out = {k: v.copy() for k, v in out.items()}
    else:
      eng.actions.copy_(torch.from_numpy(actions)); eng.values.copy_(torch.from_numpy(values))
      eng.boot_value.copy_(torch.from_numpy(boot)); eng.boot_q.copy_(torch.from_numpy(boot_q))
      eng.run_device()
      torch.cuda.synchronize(dev)
      out = None
    assert eng._graphs is not None, "the benched path replays CUDA graphs"
    rew, term = _check_pass(eng, out, oracle, actions, values, boot, boot_q, table, host)
    ends += int(term.sum())
  assert np.array_equal(eng.state.pos.cpu().numpy(), np.array(oracle.pos)), "carried state differs"


def test_benched_rollout_sees_episode_ends():
  """Scripted shortest path (20 moves) for a quarter of the envs: terminals at the last step of the window, reward +1,
  auto-reset, and the returns segmented there -- the edge the random-action passes above rarely reach."""
  from unreal_b200.train.rollout import RolloutTargets
  dev = torch.device("cuda", 0)
  n = 256
  eng = RolloutTargets(n, T, GAMMA, GAMMA_PC, torch.float32, dev, auto_reset=True, use_graphs=True, window_kernel=True)
  # S=(0,2) -> down to (0,6)... find a shortest path with the oracle's own move function (BFS)
  from collections import deque
  prev = {O.START: None}
  dq = deque([O.START])
  while dq:
    c = dq.popleft()
    if c == O.GOAL:
      break
    for a in range(4):
      nx, ny, hit = O.maze_move(c[0], c[1], a)
      if not hit and (nx, ny) not in prev:
        prev[(nx, ny)] = (c, a); dq.append((nx, ny))
  path = []
  c = O.GOAL
  while prev[c] is not None:
    c, a = prev[c]
    path.append(a)
  path.reverse()
  assert len(path) == 20                       # the optimal path of this map is exactly 20 moves
  rs = np.random.RandomState(1)
  actions = rs.randint(0, 4, size=(T, n)).astype(np.int32)
  actions[:, ::4] = np.array(path, np.int32)[:, None]
  values = rs.randn(T, n).astype(np.float32); boot = rs.randn(n).astype(np.float32)
  boot_q = rs.rand(n, 20, 20).astype(np.float32)
  out = {k: v.copy() for k, v in eng.run_host(actions, values, boot, boot_q).items()}
  oracle = _OracleBatch(n)
  rew, term = _check_pass(eng, out, oracle, actions, values, boot, boot_q, _frame_table(torch.float32, dev), True)
  assert term[T - 1, ::4].all() and (rew[T - 1, ::4] == 1).all()
  assert np.array_equal(eng.state.pos.cpu().numpy()[::4], np.tile(np.array(O.START), (n // 4, 1)))
