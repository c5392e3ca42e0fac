"""Device MT19937 streams vs numpy's RandomState: bit-exact draws and actions (C ABI)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
  from unreal_b200 import kernels, _lib
  _lib.require_device()
  return kernels


def test_randint_streams_match_numpy(K):
  seeds = [0xA3C, 0, 1, 2 ** 32 - 1, 12345] + list(range(100, 140))
  ms = K.MtStreams(seeds, "cuda:0")
  got = ms.randint(1978, k=700).cpu().numpy()          # crosses the 624-word block boundary
  got2 = ms.randint(2, k=50).cpu().numpy()
  got3 = ms.randint(1, k=5).cpu().numpy()              # single-valued: consumes nothing
  got4 = ms.randint(1997, k=30).cpu().numpy()
  for i, s in enumerate(seeds):
    rs = np.random.RandomState(s)
    assert list(got[i]) == [rs.randint(0, 1978) for _ in range(700)]
    assert list(got2[i]) == [rs.randint(2) for _ in range(50)]
    assert list(got3[i]) == [0] * 5
    assert list(got4[i]) == [rs.randint(1997) for _ in range(30)]


def test_choose_action_matches_numpy_choice(K):
  n, a = 300, 4
  seeds = np.arange(n) * 7 + 3
  ms = K.MtStreams(seeds, "cuda:0")
  rss = [np.random.RandomState(int(s)) for s in seeds]
  gen = np.random.RandomState(0)
  active = np.ones(n, np.uint8)
  for step in range(40):
    logits = gen.randn(n, a).astype(np.float32) * 2
    e = np.exp(logits - logits.max(1, keepdims=True))
    pi = (e / e.sum(1, keepdims=True)).astype(np.float32)
    if step == 20:
      active[::3] = 0                                    # finished rollouts stop drawing
    out = torch.full((n,), -1, dtype=torch.int32, device="cuda:0")
    ms.choose_action(torch.from_numpy(pi).cuda(), torch.from_numpy(active).cuda(), out)
    got = out.cpu().numpy()
    for i in range(n):
      if active[i]:
        assert got[i] == rss[i].choice(a, p=pi[i])
      else:
        assert got[i] == -1
  # streams stay aligned afterwards
  tail = ms.randint(1000, k=3).cpu().numpy()
  for i in range(n):
    assert list(tail[i]) == [rss[i].randint(0, 1000) for _ in range(3)]


def test_choice_six_actions_uniform(K):
  n, a = 64, 6
  ms = K.MtStreams(np.arange(n), "cuda:0")
  pi = torch.full((n, a), 1.0 / a, device="cuda:0")
  got = ms.choose_action(pi).cpu().numpy()
  for i in range(n):
    assert got[i] == np.random.RandomState(i).choice(a, p=np.full(a, 1.0 / a, np.float32))
