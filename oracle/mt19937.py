"""Restatement of numpy's legacy ``RandomState`` draws used on the hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference draws every replay index and every action from one
``np.random.RandomState`` (main.py:213, trainer.py:147-148, experience.py:103,
:125, :137-141).  numpy is a third-party dependency of the reference
(requirements.txt); the algorithm restated here is the published MT19937
(Matsumoto & Nishimura 1998, ``init_genrand`` / ``genrand_int32``) plus numpy's
legacy distributions layer:

* ``randint(low, high)``  -> masked rejection on 32-bit words, *no draw* when
  the range has a single value (``_rand_int64``/``_bounded_uint64`` legacy path
  with ``use_masked=True``).
* ``random_sample()``     -> ``(a >> 5, b >> 6)`` 53-bit double.
* ``choice(n, p=p)``      -> ``cdf = cumsum(float64(p)); cdf /= cdf[-1];``
  ``searchsorted(cdf, random_sample(), 'right')``.

``tests/test_oracle_rng.py`` pins this class against ``np.random.RandomState``
itself on thousands of mixed draws.
"""
import numpy as np

N = 624
M = 397
UPPER = 0x80000000
LOWER = 0x7FFFFFFF
MATRIX_A = 0x9908B0DF


class LegacyRandomState(object):
  """Word-at-a-time MT19937 with numpy-legacy ``randint`` / ``choice``."""

  def __init__(self, seed):
    self.seed(seed)

  def seed(self, seed):
    s = int(seed) & 0xFFFFFFFF
    mt = [0] * N
    mt[0] = s
    for i in range(1, N):
      s = (1812433253 * (s ^ (s >> 30)) + i) & 0xFFFFFFFF
      mt[i] = s
    self.mt = mt
    self.pos = N          # next word index; N means "regenerate first"
    self.words_drawn = 0  # bookkeeping for tests

  # The twist is done lazily one word at a time, in place.  This is equivalent
  # to the usual block regeneration because word i of the new block depends on
  # old mt[i], old mt[i+1] and mt[(i+397) % 624], which is already new exactly
  # when the block algorithm would read the new value.
  def next_u32(self):
    if self.pos >= N:
      self.pos = 0
    i = self.pos
    mt = self.mt
    y = (mt[i] & UPPER) | (mt[(i + 1) % N] & LOWER)
    v = mt[(i + M) % N] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    mt[i] = v
    self.pos = i + 1
    self.words_drawn += 1
    # tempering
    v ^= v >> 11
    v ^= (v << 7) & 0x9D2C5680
    v ^= (v << 15) & 0xEFC60000
    v ^= v >> 18
    return v & 0xFFFFFFFF

  def random_sample(self):
    a = self.next_u32() >> 5
    b = self.next_u32() >> 6
    return (a * 67108864.0 + b) / 9007199254740992.0

  def randint(self, low, high=None):
    if high is None:
      low, high = 0, low
    rng = high - low - 1
    if rng < 0:
      raise ValueError("low >= high")
    if rng == 0:
      return low
    mask = rng
    mask |= mask >> 1
    mask |= mask >> 2
    mask |= mask >> 4
    mask |= mask >> 8
    mask |= mask >> 16
    while True:
      v = self.next_u32() & mask
      if v <= rng:
        return low + v

  def choice(self, n, p):
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    u = self.random_sample()
    return int(np.searchsorted(cdf, u, side='right'))
