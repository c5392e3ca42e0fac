// Per-step bookkeeping of the batched rollout (Trainer._process_base, trainer.py:228-296), two small fused kernels
// instead of ~20 element-wise launches per env step: the rollout loop is otherwise large kernels (conv / GEMM / K1)
// separated by this glue, and at 8192 envs every tiny launch still costs 2-3 us of device time inside the CUDA graph.
#include "common.cuh"

namespace unreal {

// ExperienceFrame.concat_action_and_reward (experience.py:34-46) for every env: one-hot(last_action, A) ++ [last_reward]
// (++ objective [G] when present) -> lar [N, A+1+G] f32
__global__ void rollout_lar_kernel(const int32_t* __restrict__ last_action, const float* __restrict__ last_reward,
                                   const float* __restrict__ objective, int n, int A, int G, float* __restrict__ lar) {
  const int w = A + 1 + G;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * w) return;
  const int e = (int)(i / w), j = (int)(i - (int64_t)e * w);
  float v;
  if (j < A) {
    int a = last_action[e];
    a = a < 0 ? 0 : (a >= A ? A - 1 : a);
    v = (j == a) ? 1.f : 0.f;
  } else if (j == A) {
    v = last_reward[e];
  } else {
    v = objective[(int64_t)e * G + (j - A - 1)];
  }
  lar[i] = v;
}

// After the env step (trainer.py:265-296 with the terminal handling as masked arithmetic): one warp per env.
//   term_now = terminal & active;  last_rec = active ? frame_rec : last_rec;  episode_reward += reward;
//   stats += (finished episodes, sum of their scores);  ended |= term_now;  LSTM state rows of finished envs <- 0
//   (local_network.reset_state :293);  episode_reward <- 0 for them;  active &= ~term_now.
__global__ void __launch_bounds__(256) rollout_post_kernel(const float* __restrict__ reward, const uint8_t* __restrict__ terminal,
                                                           const uint64_t* __restrict__ frame_rec, int n,
                                                           uint8_t* __restrict__ active, uint8_t* __restrict__ ended,
                                                           uint64_t* __restrict__ last_rec, float* __restrict__ episode_reward,
                                                           float* __restrict__ lstm_c, float* __restrict__ lstm_h,
                                                           double* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int e = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  float fin = 0.f, score = 0.f;
  if (e < n) {
    int term_now = 0;
    if (lane == 0) {
      const int act = active[e];
      term_now = (terminal[e] != 0) && act;
      if (act) last_rec[e] = frame_rec[e];
      float er = episode_reward[e] + reward[e];
      if (term_now) { fin = 1.f; score = er; er = 0.f; ended[e] = 1; active[e] = 0; }
      episode_reward[e] = er;
    }
    term_now = __shfl_sync(0xffffffffu, term_now, 0);
    if (term_now && lstm_c != nullptr) {
      float4* c4 = reinterpret_cast<float4*>(lstm_c + (size_t)e * 256);
      float4* h4 = reinterpret_cast<float4*>(lstm_h + (size_t)e * 256);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      c4[lane] = z; c4[lane + 32] = z; h4[lane] = z; h4[lane + 32] = z;
    }
  }
  if (stats == nullptr) return;
  // block reduction of (finished, score): lane 0 of each warp holds its env's contribution
  __shared__ float s_f[8], s_s[8];
  if (lane == 0) { s_f[threadIdx.x >> 5] = fin; s_s[threadIdx.x >> 5] = score; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double f = 0.0, s = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { f += (double)s_f[q]; s += (double)s_s[q]; }
    if (f != 0.0) { atomicAdd(stats, f); atomicAdd(stats + 1, s); }
  }
}

}  // namespace unreal

using namespace unreal;

extern "C" int unreal_rollout_lar(const int32_t* last_action, const float* last_reward, const float* objective, int n, int a,
                                  int g, float* lar, void* stream) {
  UNREAL_REQUIRE(last_action && last_reward && lar && n > 0 && a >= 1 && g >= 0, "unreal_rollout_lar: null buffer or bad sizes");
  UNREAL_REQUIRE(g == 0 || objective != nullptr, "unreal_rollout_lar: objective size %d without an objective buffer", g);
  const int64_t total = (int64_t)n * (a + 1 + g);
  rollout_lar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(last_action, last_reward, objective, n, a, g, lar);
  UNREAL_LAUNCH_CHECK("rollout_lar_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_rollout_post(const float* reward, const uint8_t* terminal, const uint64_t* frame_rec, int n,
                                   uint8_t* active, uint8_t* ended, uint64_t* last_rec, float* episode_reward,
                                   float* lstm_c, float* lstm_h, double* stats, void* stream) {
  UNREAL_REQUIRE(reward && terminal && frame_rec && active && ended && last_rec && episode_reward && n > 0,
                 "unreal_rollout_post: null buffer or n <= 0");
  UNREAL_REQUIRE((lstm_c == nullptr) == (lstm_h == nullptr), "unreal_rollout_post: pass both LSTM state buffers or neither");
  UNREAL_REQUIRE(aligned16(lstm_c) && aligned16(lstm_h), "unreal_rollout_post: LSTM state must be 16-byte aligned");
  const int64_t threads = (int64_t)n * 32;
  rollout_post_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(
      reward, terminal, frame_rec, n, active, ended, last_rec, episode_reward, lstm_c, lstm_h, stats);
  UNREAL_LAUNCH_CHECK("rollout_post_kernel");
  return UNREAL_OK;
}
