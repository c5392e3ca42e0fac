"""Pins oracle/mt19937.py against numpy's own legacy RandomState (CPU)."""
import numpy as np
import pytest

from oracle.mt19937 import LegacyRandomState


@pytest.mark.parametrize("seed", [0xA3C, 0, 1, 12345, 2 ** 32 - 1])
def test_mixed_draws_match_numpy(seed):
  a = np.random.RandomState(seed)
  b = LegacyRandomState(seed)
  drv = np.random.RandomState(7)
  for _ in range(4000):
    k = drv.randint(5)
    if k == 0:
      hi = int(drv.randint(1, 3000))
      assert a.randint(0, hi) == b.randint(0, hi)
    elif k == 1:
      assert a.random_sample() == b.random_sample()
    elif k == 2:
      p = drv.rand(4).astype(np.float32)
      p /= p.sum()
      assert a.choice(4, p=p) == b.choice(4, p)
    elif k == 3:
      assert a.randint(2) == b.randint(2)
    else:
      assert a.randint(1) == b.randint(1)   # single-valued range: no draw


def test_survey_known_values():
  # RandomState(0xA3C).randint(0,1978) x5
  r = LegacyRandomState(0xA3C)
  assert [r.randint(0, 1978) for _ in range(5)] == [1708, 1619, 21, 1597, 1836]


def test_block_boundary():
  # more than three 624-word blocks, so the lazy in-place twist wraps repeatedly
  a = np.random.RandomState(3)
  b = LegacyRandomState(3)
  xs = [a.randint(0, 1 << 31) for _ in range(2000)]
  ys = [b.randint(0, 1 << 31) for _ in range(2000)]
  assert xs == ys
