// numpy legacy RandomState on MT19937, one stream per env, state word-major [624, N].
// Restates what the reference draws through np.random.RandomState (main.py:213,
// trainer.py:147-148, experience.py:103,:125,:137-141); algorithm notes in oracle/mt19937.py.
#pragma once
#include "common.cuh"

namespace unreal {

struct MtStream {
  uint32_t* mt;    // this env's word 0; word i is mt[i * stride]
  int64_t stride;  // = N
  int32_t* pos;    // next word index, 624 = start of a fresh block
};

UNREAL_HD void mt_seed_core(MtStream s, uint32_t seed) {
  uint32_t v = seed;
  s.mt[0] = v;
  for (int i = 1; i < UNREAL_MT_WORDS; ++i) {
    v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
    s.mt[(int64_t)i * s.stride] = v;
  }
  *s.pos = UNREAL_MT_WORDS;
}

// One tempered word.  The state is kept in numpy's own representation -- a block of 624 words
// regenerated all at once when pos reaches 624 -- so a stream can be imported from / exported
// to np.random.RandomState.get_state()/set_state() at any point (the scalar drop-in classes
// share one RandomState with their caller, main.py:213,:270).
UNREAL_HD uint32_t mt_twist(uint32_t u, uint32_t v, uint32_t m) {
  uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

UNREAL_HD void mt_regenerate(MtStream s) {
  const int64_t st = s.stride;
  int kk = 0;
  for (; kk < UNREAL_MT_WORDS - 397; ++kk)
    s.mt[kk * st] = mt_twist(s.mt[kk * st], s.mt[(kk + 1) * st], s.mt[(kk + 397) * st]);
  for (; kk < UNREAL_MT_WORDS - 1; ++kk)
    s.mt[kk * st] = mt_twist(s.mt[kk * st], s.mt[(kk + 1) * st], s.mt[(kk + 397 - UNREAL_MT_WORDS) * st]);
  s.mt[(UNREAL_MT_WORDS - 1) * st] = mt_twist(s.mt[(UNREAL_MT_WORDS - 1) * st], s.mt[0], s.mt[396 * st]);
}

UNREAL_HD uint32_t mt_next(MtStream s) {
  int i = *s.pos;
  if (i >= UNREAL_MT_WORDS) {
    mt_regenerate(s);
    i = 0;
  }
  uint32_t v = s.mt[(int64_t)i * s.stride];
  *s.pos = i + 1;
  v ^= v >> 11;
  v ^= (v << 7) & 0x9d2c5680u;
  v ^= (v << 15) & 0xefc60000u;
  v ^= v >> 18;
  return v;
}

// RandomState.randint(0, high): masked rejection; a single-valued range consumes no word.
UNREAL_HD uint32_t mt_randint(MtStream s, uint32_t high) {
  uint32_t rng = high - 1u;
  if (rng == 0u) return 0u;
  uint32_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  do { v = mt_next(s) & mask; } while (v > rng);
  return v;
}

// RandomState.random_sample(): 53-bit double from two words.
UNREAL_HD double mt_random_sample(MtStream s) {
  uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
  return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

// RandomState.choice(A, p=pi): cdf = cumsum(float64(pi)); cdf /= cdf[-1];
// searchsorted(cdf, u, side='right')  (trainer.py:147-148).  A <= 32.
UNREAL_HD int mt_choice(MtStream s, const float* pi, int a) {
  double cdf[32];
  double acc = 0.0;
  for (int i = 0; i < a; ++i) { acc += (double)pi[i]; cdf[i] = acc; }
  double last = cdf[a - 1];
  double u = mt_random_sample(s);
  int k = 0;
  for (int i = 0; i < a; ++i) k += (cdf[i] / last <= u) ? 1 : 0;  // count of cdf <= u == 'right' insertion point (cdf sorted)
  return k;
}

}  // namespace unreal
