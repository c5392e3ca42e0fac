// Per-env numpy-legacy RandomState streams on the device (see mt19937_core.cuh) and
// Trainer.choose_action (train/trainer.py:147-148).
#include "common.cuh"
#include "mt19937_core.cuh"

namespace unreal {

__global__ void mt_seed_kernel(uint32_t* mt, int32_t* pos, const uint32_t* seeds, int n) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  MtStream s{mt + e, (int64_t)n, pos + e};
  mt_seed_core(s, seeds[e]);
}

__global__ void choose_action_kernel(uint32_t* mt, int32_t* pos, const float* __restrict__ pi,
                                     const uint8_t* __restrict__ active, int32_t* action, int n, int a) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (active != nullptr && active[e] == 0) return;  // a finished rollout draws nothing
  MtStream s{mt + e, (int64_t)n, pos + e};
  action[e] = mt_choice(s, pi + (size_t)e * a, a);
}

__global__ void mt_randint_kernel(uint32_t* mt, int32_t* pos, uint32_t high, int32_t* out, int n, int k) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  MtStream s{mt + e, (int64_t)n, pos + e};
  for (int i = 0; i < k; ++i) out[(size_t)e * k + i] = (int32_t)mt_randint(s, high);
}

}  // namespace unreal

using namespace unreal;

extern "C" int unreal_mt_seed(uint32_t* mt, int32_t* mt_pos, const uint32_t* seeds, int n, void* stream) {
  UNREAL_REQUIRE(mt && mt_pos && seeds && n >= 0, "unreal_mt_seed: null argument or n < 0");
  if (n == 0) return UNREAL_OK;
  mt_seed_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(mt, mt_pos, seeds, n);
  UNREAL_LAUNCH_CHECK("mt_seed_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_choose_action(uint32_t* mt, int32_t* mt_pos, const float* pi, const uint8_t* active,
                                    int32_t* action, int n, int a, void* stream) {
  UNREAL_REQUIRE(mt && mt_pos && pi && action && n >= 0, "unreal_choose_action: null argument or n < 0");
  UNREAL_REQUIRE(a >= 1 && a <= 32, "unreal_choose_action: action size %d not in 1..32", a);
  if (n == 0) return UNREAL_OK;
  choose_action_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(mt, mt_pos, pi, active, action, n, a);
  UNREAL_LAUNCH_CHECK("choose_action_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_mt_randint(uint32_t* mt, int32_t* mt_pos, uint32_t high, int32_t* out, int n, int k,
                                 void* stream) {
  UNREAL_REQUIRE(mt && mt_pos && out && n >= 0 && k >= 0 && high >= 1, "unreal_mt_randint: bad argument");
  if (n == 0 || k == 0) return UNREAL_OK;
  mt_randint_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(mt, mt_pos, high, out, n, k);
  UNREAL_LAUNCH_CHECK("mt_randint_kernel");
  return UNREAL_OK;
}
