"""K3/K4 parity: n-step returns, advantages, PC Q-targets, sequence returns (C ABI)."""
import numpy as np
import pytest
import torch

from oracle import unreal_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-5   # north-star fp32 tolerance (relative), vs the float64 evaluation


@pytest.fixture(scope="module")
def K():
  from unreal_b200 import kernels, _lib
  _lib.require_device()
  return kernels


def _close(got, want, rel=REL):
  want = np.asarray(want, np.float64)
  scale = np.maximum(np.abs(want), 1.0)
  assert np.max(np.abs(got.astype(np.float64) - want) / scale) <= rel


@pytest.mark.parametrize("T,N", [(20, 4096), (20, 1), (1, 7), (50, 1000), (7, 129)])
def test_nstep_returns(K, T, N):
  rs = np.random.RandomState(T * 1000 + N)
  r = rs.randint(-1, 2, size=(T, N)).astype(np.float32)
  v = rs.randn(T, N).astype(np.float32)
  boot = rs.randn(N).astype(np.float32)
  term = (rs.rand(T, N) < 0.05).astype(np.uint8)
  dev = "cuda:0"
  R, adv = K.nstep_returns(*[torch.from_numpy(a).to(dev) for a in (r, v, term, boot)], 0.99)
  R32, adv32 = O.nstep_returns_segmented(r, v, term, boot, 0.99, np.float32)
  R64, adv64 = O.nstep_returns_segmented(r, v, term, boot, 0.99, np.float64)
  assert np.array_equal(R.cpu().numpy(), R32) and np.array_equal(adv.cpu().numpy(), adv32)   # bit-exact fp32
  _close(R.cpu().numpy(), R64)
  _close(adv.cpu().numpy(), adv64)
  # value-replay form: no v / adv
  R2, none = K.nstep_returns(torch.from_numpy(r).to(dev), None, torch.from_numpy(term).to(dev),
                             torch.from_numpy(boot).to(dev), 0.99)
  assert none is None and torch.equal(R2, R)


def test_nstep_matches_reference_rollout_form(K):
  """Windows that are exactly reference rollouts (terminal only as last step)."""
  rs = np.random.RandomState(3)
  T, N = 20, 257
  lens = rs.randint(1, T + 1, size=N)
  r = np.zeros((T, N), np.float32); v = np.zeros((T, N), np.float32); term = np.zeros((T, N), np.uint8)
  boot = rs.randn(N).astype(np.float32)
  ended = rs.rand(N) < 0.5
  for n in range(N):
    r[:lens[n], n] = rs.randint(-1, 2, size=lens[n]); v[:lens[n], n] = rs.randn(lens[n])
    if ended[n]:
      term[lens[n] - 1, n] = 1
    elif lens[n] < T:
      lens[n] = T; r[:, n] = rs.randint(-1, 2, size=T); v[:, n] = rs.randn(T)
  dev = "cuda:0"
  R, adv = K.nstep_returns(*[torch.from_numpy(a).to(dev) for a in (r, v, term, boot)], 0.99)
  R, adv = R.cpu().numpy(), adv.cpu().numpy()
  for n in range(N):
    L = lens[n]
    R1, a1 = O.nstep_returns(r[:L, n], v[:L, n], 0.0 if ended[n] else boot[n], 0.99, np.float64)
    _close(R[:L, n], R1); _close(adv[:L, n], a1)


@pytest.mark.parametrize("T,N", [(20, 512), (20, 1), (3, 37), (21, 300)])
@pytest.mark.parametrize("use_len,use_term", [(False, False), (True, False), (False, True)])
def test_pc_targets(K, T, N, use_len, use_term):
  rs = np.random.RandomState(T + N)
  pc = (rs.randint(0, 5, size=(T, N, 20, 20)) * 4 / 48.0).astype(np.float32)
  boot = rs.rand(N, 20, 20).astype(np.float32)
  lens = rs.randint(1, T + 1, size=N).astype(np.int32) if use_len else None
  term = (rs.rand(T, N) < 0.1).astype(np.uint8) if use_term else None
  dev = "cuda:0"
  got = K.pc_targets(torch.from_numpy(pc).to(dev), None if term is None else torch.from_numpy(term).to(dev),
                     None if lens is None else torch.from_numpy(lens).to(dev), torch.from_numpy(boot).to(dev),
                     0.9).cpu().numpy()
  for n in range(0, N, max(1, N // 50)):
    L = int(lens[n]) if use_len else T
    if use_term:
      R = boot[n].astype(np.float64); want = np.zeros((T, 20, 20))
      for t in range(T - 1, -1, -1):
        if term[t, n]:
          R = np.zeros((20, 20))
        R = pc[t, n].astype(np.float64) + 0.9 * R
        want[t] = R
    else:
      want = np.zeros((T, 20, 20))
      want[:L] = O.pc_targets(pc[:L, n].astype(np.float64), boot[n].astype(np.float64), 0.9)
    _close(got[:, n], want)
    assert (got[L:, n] == 0).all()


@pytest.mark.parametrize("T,N", [(20, 512), (20, 1), (3, 37), (21, 300), (1, 5)])
@pytest.mark.parametrize("use_len", [False, True])
def test_maze_pc_targets_equals_pixel_change_then_targets(K, T, N, use_len):
  """unreal_maze_pc_targets (the replayed frames' pixel-change maps evaluated in registers inside the Q-target scan,
  trainer.py:339-380) against the two kernels it replaces -- bit for bit -- and against the oracle's closed-form maps +
  float64 scan; stationary steps (no pixel change), lengths from 0 to T."""
  rs = np.random.RandomState(7 * T + N)
  dev = "cuda:0"
  p0 = rs.randint(0, 7, size=(T, N, 2)).astype(np.int32)
  step = rs.randint(-1, 2, size=(T, N, 2)).astype(np.int32)
  step[rs.rand(T, N) < 0.5, 1] = 0                               # mostly axis moves, some stationary, a few diagonal
  p1 = np.clip(p0 + step, 0, 6).astype(np.int32)
  boot = rs.rand(N, 20, 20).astype(np.float32)
  lens = rs.randint(0, T + 1, size=N).astype(np.int32) if use_len else None
  t0, t1, tb = (torch.from_numpy(a).to(dev) for a in (p0, p1, boot))
  tl = None if lens is None else torch.from_numpy(lens).to(dev)
  got = K.maze_pc_targets(t0, t1, tl, tb, 0.9)
  pc = K.maze_pixel_change(t0.view(-1, 2), t1.view(-1, 2)).view(T, N, 20, 20)
  assert torch.equal(got, K.pc_targets(pc, None, tl, tb, 0.9))
  got = got.cpu().numpy()
  for n in range(0, N, max(1, N // 25)):
    L = int(lens[n]) if use_len else T
    maps = np.stack([O.maze_pixel_change_closed_form(int(p0[t, n, 0]), int(p0[t, n, 1]), int(p1[t, n, 0]), int(p1[t, n, 1]))
                     for t in range(L)]) if L else np.zeros((0, 20, 20))
    if L:
      _close(got[:L, n], O.pc_targets(maps.astype(np.float64), boot[n].astype(np.float64), 0.9))
    assert (got[L:, n] == 0).all()


@pytest.mark.parametrize("L", [21, 20, 5, 32, 1])
def test_sequence_returns_warp_scan(K, L):
  rs = np.random.RandomState(L)
  N = 999
  r = rs.randint(-1, 2, size=(N, L)).astype(np.float32)
  lens = rs.randint(1, L + 1, size=N).astype(np.int32)
  boot = rs.randn(N).astype(np.float32)
  dev = "cuda:0"
  got = K.sequence_returns(torch.from_numpy(r).to(dev), torch.from_numpy(lens).to(dev),
                           torch.from_numpy(boot).to(dev), 0.99).cpu().numpy()
  for n in range(N):
    want = O.vr_returns(list(r[n, :lens[n]].astype(np.float64)), float(boot[n]), 0.99)
    _close(got[n, :lens[n]], want)
    assert (got[n, lens[n]:] == 0).all()


def test_empty_inputs(K):
  dev = "cuda:0"
  z = torch.empty(0, 5, device=dev)
  R, adv = K.nstep_returns(z, z, torch.empty(0, 5, dtype=torch.uint8, device=dev), torch.zeros(5, device=dev), 0.99)
  assert R.shape == (0, 5)
  out = K.pc_targets(torch.empty(4, 0, 20, 20, device=dev), None, None, torch.empty(0, 20, 20, device=dev), 0.9)
  assert out.shape == (4, 0, 20, 20)
  zi = torch.empty(4, 0, 2, dtype=torch.int32, device=dev)
  assert K.maze_pc_targets(zi, zi, None, torch.empty(0, 20, 20, device=dev), 0.9).shape == (4, 0, 20, 20)
