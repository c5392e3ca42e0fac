// K5: device-resident replay ring with reward-prediction-balanced and sequence sampling.
// Replaces Experience (train/experience.py:48-153) for N envs at once; each env keeps the
// reference's exact index semantics and draws from its own numpy-legacy MT19937 stream, so
// sampled indices are bit-identical to `Experience(H, RandomState(seed_n))` fed the same frames.
//
// HBM layout: rec [N, H] u64 env-major (a sampled sequence is 21 consecutive records = 168
// contiguous bytes; an RP rank-select streams one env's records linearly), top [N] i64,
// count / n_pos / n_neg [N] i32.  8 B per frame: 2000 frames x 8192 envs = 131 MB per GPU,
// where the reference's deque of 84x84x3 float64 frames would be 2.8 TB.
#include <new>

#include "common.cuh"
#include "maze_core.cuh"
#include "mt19937_core.cuh"
#include "ring_core.cuh"

struct unreal_replay {
  int n, h;
  uint64_t* rec;
  int64_t* top;
  int32_t* count;
  int32_t* n_pos;
  int32_t* n_neg;
};

namespace unreal {

__device__ __forceinline__ RingRef ring_of(const unreal_replay& R, int e) {
  return RingRef{R.rec + (size_t)e * R.h, R.h, R.top + e, R.count + e, R.n_pos + e, R.n_neg + e};
}

__global__ void replay_add_kernel(unreal_replay R, const uint64_t* __restrict__ frames) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R.n) return;
  ring_add(ring_of(R, e), frames[e]);
}

__global__ void replay_state_kernel(unreal_replay R, uint8_t* full, int32_t* count, int64_t* top, int32_t* n_pos,
                                    int32_t* n_neg) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R.n) return;
  if (full) full[e] = R.count[e] >= R.h;   // is_full :96-97
  if (count) count[e] = R.count[e];
  if (top) top[e] = R.top[e];
  if (n_pos) n_pos[e] = R.n_pos[e];
  if (n_neg) n_neg[e] = R.n_neg[e];
}

// one warp per env; lane 0 owns the RNG stream, lanes gather the records
__global__ void __launch_bounds__(128) replay_sample_sequence_kernel(unreal_replay R, uint32_t* mt, int32_t* mt_pos,
                                                                     int L, int32_t* out_start, int32_t* out_len,
                                                                     uint64_t* out_rec) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= R.n) return;
  const RingRef r = ring_of(R, e);
  const bool ready = *r.count >= R.h;      // the reference only samples a full buffer (trainer.py:446-448)
  int start = -1;
  if (lane == 0 && ready) {
    MtStream s{mt + e, (int64_t)R.n, mt_pos + e};
    start = (int)mt_randint(s, (uint32_t)(R.h - L - 1));   // randint(0, H - L - 1)  :103
  }
  start = __shfl_sync(0xffffffffu, start, 0);
  if (!ready) {
    if (lane == 0) { out_start[e] = -1; out_len[e] = 0; }
    if (lane < L) out_rec[(size_t)e * L + lane] = 0ull;
    return;
  }
  // lanes 0..L hold raw positions start..start+L; shift by one when the first is terminal (:105-107)
  uint64_t v = (lane <= L) ? ring_at_raw(r, start + lane) : 0ull;
  const int first_term = __shfl_sync(0xffffffffu, frame_terminal(v), 0);
  if (first_term) {
    v = __shfl_down_sync(0xffffffffu, v, 1);
    start += 1;
  }
  const unsigned in_seq = (L >= 32) ? 0xffffffffu : ((1u << L) - 1u);
  const unsigned terms = __ballot_sync(0xffffffffu, frame_terminal(v) != 0) & in_seq;
  const int len = terms ? (__ffs(terms)) : L;            // stop after the first terminal (:111-116)
  if (lane == 0) { out_start[e] = start; out_len[e] = len; }
  if (lane < L) out_rec[(size_t)e * L + lane] = (lane < len) ? v : 0ull;
}

__global__ void __launch_bounds__(128) replay_sample_rp_kernel(unreal_replay R, uint32_t* mt, int32_t* mt_pos,
                                                               int32_t* out_start, uint64_t* out_rec) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= R.n) return;
  const RingRef r = ring_of(R, e);
  const int n_pos = *r.n_pos, n_neg = *r.n_neg;
  if (n_pos + n_neg <= 0) {                 // fewer than 4 frames: the reference would raise
    if (lane == 0) out_start[e] = -1;
    if (lane < 4) out_rec[(size_t)e * 4 + lane] = 0ull;
    return;
  }
  int from_neg = 0, k = 0;
  if (lane == 0) {
    MtStream s{mt + e, (int64_t)R.n, mt_pos + e};
    from_neg = (mt_randint(s, 2u) == 0u) ? 1 : 0;        // randint(2) == 0  :125-128
    if (n_pos == 0) from_neg = 1;                        // :130-135
    else if (n_neg == 0) from_neg = 0;
    k = (int)mt_randint(s, (uint32_t)(from_neg ? n_neg : n_pos));   // :137-141
  }
  from_neg = __shfl_sync(0xffffffffu, from_neg, 0);
  k = __shfl_sync(0xffffffffu, k, 0);
  // warp rank-select over the eligible absolute range, 32 frames per step
  const int64_t top = *r.top;
  const int64_t lo = top + 3 < 3 ? 3 : top + 3;
  const int64_t hi = top + *r.count - 1;
  int64_t end = -1;
  for (int64_t base = lo; base <= hi; base += 32) {
    const int64_t a = base + lane;
    bool match = false;
    if (a <= hi) match = ((frame_reward(ring_at_abs(r, a)) > 0) ? 1 : 0) != from_neg;
    const unsigned b = __ballot_sync(0xffffffffu, match);
    const int c = __popc(b);
    if (k < c) {
      unsigned bits = b;
      for (int i = 0; i < k; ++i) bits &= bits - 1;       // drop the k lowest set bits
      end = base + (__ffs(bits) - 1);
      break;
    }
    k -= c;
  }
  const int start = (int)(end - 3 - top);                 // raw_start_frame_index :143-144
  if (lane == 0) out_start[e] = start;
  if (lane < 4) out_rec[(size_t)e * 4 + lane] = ring_at_raw(r, start + lane);
}

__global__ void frame_unpack_kernel(const uint64_t* __restrict__ rec, int m, int32_t* pos0, int32_t* pos1,
                                    int32_t* action, float* reward, uint8_t* terminal, int32_t* last_action,
                                    float* last_reward, uint8_t* valid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t r = rec[i];
  if (pos0) { pos0[2 * i] = frame_x0(r); pos0[2 * i + 1] = frame_y0(r); }
  if (pos1) { pos1[2 * i] = frame_x1(r); pos1[2 * i + 1] = frame_y1(r); }
  if (action) action[i] = frame_action(r);
  if (reward) reward[i] = (float)frame_reward(r);
  if (terminal) terminal[i] = (uint8_t)frame_terminal(r);
  if (last_action) last_action[i] = frame_last_action(r);
  if (last_reward) last_reward[i] = (float)frame_last_reward(r);
  if (valid) valid[i] = (uint8_t)frame_valid(r);
}

}  // namespace unreal

using namespace unreal;

extern "C" int unreal_replay_create(unreal_replay_t** out, int n_envs, int history_size) {
  UNREAL_REQUIRE(out != nullptr, "unreal_replay_create: out is null");
  UNREAL_REQUIRE(n_envs >= 1 && history_size >= 4, "unreal_replay_create: need n_envs >= 1 and history_size >= 4");
  unreal_replay* r = new (std::nothrow) unreal_replay();
  if (!r) { set_error("unreal_replay_create: out of host memory"); return UNREAL_ENOMEM; }
  r->n = n_envs; r->h = history_size;
  r->rec = nullptr; r->top = nullptr; r->count = r->n_pos = r->n_neg = nullptr;
  cudaError_t e = cudaMalloc(&r->rec, (size_t)n_envs * history_size * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&r->top, (size_t)n_envs * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMalloc(&r->count, (size_t)n_envs * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&r->n_pos, (size_t)n_envs * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&r->n_neg, (size_t)n_envs * sizeof(int32_t));
  if (e != cudaSuccess) {
    unreal_replay_destroy(r);
    cuda_fail(e, "unreal_replay_create: cudaMalloc");
    return e == cudaErrorMemoryAllocation ? UNREAL_ENOMEM : UNREAL_ECUDA;
  }
  *out = r;
  int rc = unreal_replay_reset(r, nullptr);
  if (rc == UNREAL_OK) { cudaError_t s = cudaStreamSynchronize(nullptr); if (s != cudaSuccess) rc = cuda_fail(s, "sync"); }
  return rc;
}

extern "C" int unreal_replay_destroy(unreal_replay_t* r) {
  if (!r) return UNREAL_OK;
  cudaFree(r->rec); cudaFree(r->top); cudaFree(r->count); cudaFree(r->n_pos); cudaFree(r->n_neg);
  delete r;
  return UNREAL_OK;
}

extern "C" int unreal_replay_reset(unreal_replay_t* r, void* stream) {
  UNREAL_REQUIRE(r != nullptr, "unreal_replay_reset: null handle");
  cudaStream_t st = as_stream(stream);
  UNREAL_CUDA(cudaMemsetAsync(r->rec, 0, (size_t)r->n * r->h * sizeof(uint64_t), st));
  UNREAL_CUDA(cudaMemsetAsync(r->top, 0, (size_t)r->n * sizeof(int64_t), st));
  UNREAL_CUDA(cudaMemsetAsync(r->count, 0, (size_t)r->n * sizeof(int32_t), st));
  UNREAL_CUDA(cudaMemsetAsync(r->n_pos, 0, (size_t)r->n * sizeof(int32_t), st));
  UNREAL_CUDA(cudaMemsetAsync(r->n_neg, 0, (size_t)r->n * sizeof(int32_t), st));
  return UNREAL_OK;
}

extern "C" int unreal_replay_add(unreal_replay_t* r, const uint64_t* frame_rec, void* stream) {
  UNREAL_REQUIRE(r && frame_rec, "unreal_replay_add: null argument");
  replay_add_kernel<<<(r->n + 127) / 128, 128, 0, as_stream(stream)>>>(*r, frame_rec);
  UNREAL_LAUNCH_CHECK("replay_add_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_replay_state(unreal_replay_t* r, uint8_t* full, int32_t* count, int64_t* top, int32_t* n_pos,
                                   int32_t* n_neg, void* stream) {
  UNREAL_REQUIRE(r != nullptr, "unreal_replay_state: null handle");
  replay_state_kernel<<<(r->n + 127) / 128, 128, 0, as_stream(stream)>>>(*r, full, count, top, n_pos, n_neg);
  UNREAL_LAUNCH_CHECK("replay_state_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_replay_sample_sequence(unreal_replay_t* r, uint32_t* mt, int32_t* mt_pos, int seq_len,
                                             int32_t* start, int32_t* len, uint64_t* rec, void* stream) {
  UNREAL_REQUIRE(r && mt && mt_pos && start && len && rec, "unreal_replay_sample_sequence: null argument");
  UNREAL_REQUIRE(seq_len >= 1 && seq_len <= 31, "unreal_replay_sample_sequence: seq_len %d not in 1..31", seq_len);
  UNREAL_REQUIRE(r->h - seq_len - 1 >= 1, "unreal_replay_sample_sequence: history %d too small for seq_len %d", r->h, seq_len);
  long long threads = (long long)r->n * 32;
  replay_sample_sequence_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(
      *r, mt, mt_pos, seq_len, start, len, rec);
  UNREAL_LAUNCH_CHECK("replay_sample_sequence_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_replay_sample_rp(unreal_replay_t* r, uint32_t* mt, int32_t* mt_pos, int32_t* start,
                                       uint64_t* rec, void* stream) {
  UNREAL_REQUIRE(r && mt && mt_pos && start && rec, "unreal_replay_sample_rp: null argument");
  long long threads = (long long)r->n * 32;
  replay_sample_rp_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(*r, mt, mt_pos, start, rec);
  UNREAL_LAUNCH_CHECK("replay_sample_rp_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_frame_unpack(const uint64_t* rec, int m, int32_t* pos0, int32_t* pos1, int32_t* action,
                                   float* reward, uint8_t* terminal, int32_t* last_action, float* last_reward,
                                   uint8_t* valid, void* stream) {
  UNREAL_REQUIRE(m >= 0, "unreal_frame_unpack: negative size");
  if (m == 0) return UNREAL_OK;
  UNREAL_REQUIRE(rec != nullptr, "unreal_frame_unpack: rec is null");
  frame_unpack_kernel<<<(m + 255) / 256, 256, 0, as_stream(stream)>>>(rec, m, pos0, pos1, action, reward, terminal,
                                                                     last_action, last_reward, valid);
  UNREAL_LAUNCH_CHECK("frame_unpack_kernel");
  return UNREAL_OK;
}

// Checkpointing (main.py:356,469-519 saves only the network; the replay memory and RNG are lost on
// restart there -- here the ring can be exported / imported verbatim).  direction 0: ring -> caller
// buffers, 1: caller buffers -> ring.  rec [N,H] u64, top [N] i64, count / n_pos / n_neg [N] i32.
extern "C" int unreal_replay_copy(unreal_replay_t* r, int direction, uint64_t* rec, int64_t* top, int32_t* count,
                                  int32_t* n_pos, int32_t* n_neg, void* stream) {
  UNREAL_REQUIRE(r && rec && top && count && n_pos && n_neg, "unreal_replay_copy: null argument");
  UNREAL_REQUIRE(direction == 0 || direction == 1, "unreal_replay_copy: direction must be 0 (export) or 1 (import)");
  cudaStream_t st = as_stream(stream);
  const size_t n = (size_t)r->n;
  struct { void* ring; void* user; size_t bytes; } parts[5] = {
      {r->rec, rec, n * r->h * sizeof(uint64_t)}, {r->top, top, n * sizeof(int64_t)},
      {r->count, count, n * sizeof(int32_t)}, {r->n_pos, n_pos, n * sizeof(int32_t)}, {r->n_neg, n_neg, n * sizeof(int32_t)}};
  for (auto& p : parts) {
    if (direction == 0) UNREAL_CUDA(cudaMemcpyAsync(p.user, p.ring, p.bytes, cudaMemcpyDeviceToDevice, st));
    else UNREAL_CUDA(cudaMemcpyAsync(p.ring, p.user, p.bytes, cudaMemcpyDeviceToDevice, st));
  }
  return UNREAL_OK;
}
