// TEST INFRASTRUCTURE, not a product path: compiles the host+device "core" headers of
// unreal_b200/csrc (maze_core.cuh, mt19937_core.cuh, ring_core.cuh) with plain g++ so that the
// integer logic the CUDA kernels share can be checked against the golden fixtures on a
// machine without a GPU.  Nothing in the package loads this library.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../unreal_b200/csrc/maze_core.cuh"
#include "../../unreal_b200/csrc/mt19937_core.cuh"
#include "../../unreal_b200/csrc/ring_core.cuh"

using namespace unreal;

static MazeLayout reference_layout() {
  const char* m = "--+---G--+-+++S-+---+--+++----+-+----+---------++";
  MazeLayout L;
  memset(&L, 0, sizeof(L));
  for (int i = 0; i < 49; ++i) {
    int x = i % 7, y = i / 7;
    if (m[i] == '+') L.wall_rows[y] |= 1u << x;
    if (m[i] == 'S') { L.start_x = x; L.start_y = y; }
    if (m[i] == 'G') { L.goal_x = x; L.goal_y = y; }
  }
  return L;
}

extern "C" {

// out: x1, y1, reward, terminal
void h_maze_step(int x, int y, int action, int* out) {
  MazeStep s = maze_step_core(reference_layout(), x, y, action);
  out[0] = s.x1; out[1] = s.y1; out[2] = s.reward; out[3] = s.terminal;
}

int h_pc_overlap(int i, int c) { return pc_overlap(i, c); }

uint64_t h_frame_pack(int x0, int y0, int x1, int y1, int a, int r, int t, int la, int lr) {
  return frame_pack(x0, y0, x1, y1, a, r, t, la, lr);
}

// k draws of randint(0, high) from RandomState(seed)
void h_mt_randint(uint32_t seed, uint32_t high, int k, int32_t* out) {
  std::vector<uint32_t> mt(UNREAL_MT_WORDS);
  int32_t pos = 0;
  MtStream s{mt.data(), 1, &pos};
  mt_seed_core(s, seed);
  for (int i = 0; i < k; ++i) out[i] = (int32_t)mt_randint(s, high);
}

// choice(A, p=pi[i]) for i < k from one stream
void h_mt_choice(uint32_t seed, const float* pi, int a, int k, int32_t* out) {
  std::vector<uint32_t> mt(UNREAL_MT_WORDS);
  int32_t pos = 0;
  MtStream s{mt.data(), 1, &pos};
  mt_seed_core(s, seed);
  for (int i = 0; i < k; ++i) out[i] = mt_choice(s, pi + (size_t)i * a, a);
}

// Replays the protocol of tests/golden/make_golden.py:gen_experience through ring_core.
// log rows: frame_no, kind (0 seq / 1 rp), v0, v1, top, n_pos, n_neg.  Returns rows written.
int h_ring_protocol(int H, int L, uint32_t seed, int n, int every, const int8_t* reward, const uint8_t* terminal,
                    int64_t* log, int max_rows) {
  std::vector<uint64_t> rec(H, 0);
  int64_t top = 0;
  int32_t count = 0, n_pos = 0, n_neg = 0;
  RingRef r{rec.data(), H, &top, &count, &n_pos, &n_neg};
  std::vector<uint32_t> mt(UNREAL_MT_WORDS);
  int32_t pos = 0;
  MtStream s{mt.data(), 1, &pos};
  mt_seed_core(s, seed);
  int rows = 0;
  for (int i = 0; i < n; ++i) {
    ring_add(r, frame_pack(0, 0, 0, 0, 0, reward[i], terminal[i], 0, 0));
    if (count >= H && i % every == 0) {
      if (rows + 2 > max_rows) return -1;
      int len = 0;
      int start = ring_sequence(r, (int)mt_randint(s, (uint32_t)(H - L - 1)), L, &len);
      int64_t* q = log + (size_t)rows++ * 7;
      q[0] = i; q[1] = 0; q[2] = start; q[3] = len; q[4] = top; q[5] = n_pos; q[6] = n_neg;
      bool from_neg = mt_randint(s, 2u) == 0u;
      if (n_pos == 0) from_neg = true; else if (n_neg == 0) from_neg = false;
      int k = (int)mt_randint(s, (uint32_t)(from_neg ? n_neg : n_pos));
      int64_t end = ring_select(r, from_neg, k);
      q = log + (size_t)rows++ * 7;
      q[0] = i; q[1] = 1; q[2] = end - 3 - top; q[3] = end; q[4] = top; q[5] = n_pos; q[6] = n_neg;
    }
  }
  return rows;
}

}  // extern "C"
