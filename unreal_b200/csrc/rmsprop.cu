// K6: fused global-norm clip + shared RMSProp over one flat fp32 parameter buffer.
// Replaces RMSPropApplier._apply_gradients / _apply_dense (train/rmsprop_applier.py:83-132):
//   tf.clip_by_global_norm(grads, clip)        (:121)  scale = clip * min(1/norm, 1/clip)
//   training_ops.apply_rms_prop per variable   (:86-93, one launch per variable, ~20 launches)
//     ms  += (g*g - ms) * (1 - decay)
//     mom  = momentum * mom + lr * g / sqrt(ms + eps)        (eps inside the sqrt)
//     var -= mom
// Here the 20 variables live back to back in one [P] buffer: one reduction launch for the
// norm and one update launch, 20 B of HBM traffic per parameter (28 B with a momentum slot).
// The sum of squares stays on the device (double) so a sharded learner can all-reduce it
// between the two launches without a host round trip.
#include "common.cuh"

namespace unreal {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, int64_t p, double* out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t p4 = p >> 2;
  float acc = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = tid; i < p4; i += stride) {
    float4 v = __ldcs(g4 + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int64_t i = (p4 << 2) + tid; i < p; i += stride) acc += g[i] * g[i];
  double d = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  __shared__ double s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(out, t);
  }
}

struct RmsArgs {
  float grad_scale, lr, one_minus_decay, momentum, eps, clip_norm;
};

__device__ __forceinline__ void rms_one(float& var, float& ms, float& mom, float g, float scale, const RmsArgs& a) {
  g = __fmul_rn(g, scale);
  ms = __fadd_rn(ms, __fmul_rn(__fsub_rn(__fmul_rn(g, g), ms), a.one_minus_decay));
  mom = __fadd_rn(__fmul_rn(mom, a.momentum), __fdiv_rn(__fmul_rn(g, a.lr), __fsqrt_rn(__fadd_rn(ms, a.eps))));
  var = __fsub_rn(var, mom);
}

template <bool kMom>
__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ var, float* __restrict__ rms,
                                                      float* __restrict__ mom, const float* __restrict__ grad,
                                                      int64_t p, const double* __restrict__ sumsq, RmsArgs a,
                                                      float* grad_norm, const float* __restrict__ lr_dev) {
  if (lr_dev != nullptr) a.lr = *lr_dev;     // learning rate from device memory (CUDA-graph replays with an annealed rate)
  // norm of the (scaled) gradient and TF's clip factor, recomputed per thread from one double
  float norm = 0.f, scale = a.grad_scale;
  if (sumsq != nullptr) {
    norm = __fmul_rn((float)sqrt(*sumsq), fabsf(a.grad_scale));
    if (a.clip_norm > 0.f) {
      const float s = __fmul_rn(a.clip_norm, fminf(__fdiv_rn(1.0f, norm), __fdiv_rn(1.0f, a.clip_norm)));
      scale = __fmul_rn(a.grad_scale, s);
    }
  }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0 && grad_norm != nullptr) *grad_norm = norm;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t p4 = p >> 2;
  float4* v4 = reinterpret_cast<float4*>(var);
  float4* r4 = reinterpret_cast<float4*>(rms);
  float4* m4 = reinterpret_cast<float4*>(mom);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  for (int64_t i = tid; i < p4; i += stride) {
    float4 v = v4[i], r = r4[i], g = __ldcs(g4 + i);
    float4 m = kMom ? m4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    rms_one(v.x, r.x, m.x, g.x, scale, a);
    rms_one(v.y, r.y, m.y, g.y, scale, a);
    rms_one(v.z, r.z, m.z, g.z, scale, a);
    rms_one(v.w, r.w, m.w, g.w, scale, a);
    v4[i] = v; r4[i] = r;
    if (kMom) m4[i] = m;
  }
  for (int64_t i = (p4 << 2) + tid; i < p; i += stride) {
    float v = var[i], r = rms[i], m = kMom ? mom[i] : 0.f;
    rms_one(v, r, m, grad[i], scale, a);
    var[i] = v; rms[i] = r;
    if (kMom) mom[i] = m;
  }
}

static int grid_for(int64_t p, int per_thread) {
  int sms = sm_count();
  if (sms <= 0) return 0;
  int64_t want = (p / per_thread + 255) / 256;
  int64_t cap = (int64_t)sms * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace unreal

using namespace unreal;

extern "C" int unreal_grad_sumsq(const float* grad, int64_t p, double* sumsq, void* stream) {
  UNREAL_REQUIRE(p >= 0 && sumsq != nullptr, "unreal_grad_sumsq: negative size or null sumsq");
  if (p == 0) return UNREAL_OK;
  UNREAL_REQUIRE(grad != nullptr && aligned16(grad), "unreal_grad_sumsq: grad must be non-null and 16-byte aligned");
  int grid = grid_for(p, 4);
  if (grid <= 0) return UNREAL_ECUDA;
  grad_sumsq_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad, p, sumsq);
  UNREAL_LAUNCH_CHECK("grad_sumsq_kernel");
  return UNREAL_OK;
}

static int rmsprop_launch(float* var, float* rms, float* mom, const float* grad, int64_t p, const double* sumsq,
                          float grad_scale, float lr, const float* lr_dev, float decay, float momentum, float eps,
                          float clip_norm, float* grad_norm, void* stream) {
  UNREAL_REQUIRE(p >= 0, "unreal_rmsprop_update: negative size");
  if (p == 0) return UNREAL_OK;
  UNREAL_REQUIRE(var && rms && grad, "unreal_rmsprop_update: var, rms and grad must be non-null");
  UNREAL_REQUIRE(aligned16(var) && aligned16(rms) && aligned16(grad) && aligned16(mom),
                 "unreal_rmsprop_update: buffers must be 16-byte aligned");
  UNREAL_REQUIRE(mom != nullptr || momentum == 0.f, "unreal_rmsprop_update: momentum != 0 needs the mom slot");
  UNREAL_REQUIRE(clip_norm <= 0.f || sumsq != nullptr, "unreal_rmsprop_update: clipping needs sumsq");
  RmsArgs a{grad_scale, lr, 1.0f - decay, momentum, eps, clip_norm};
  int grid = grid_for(p, 4);
  if (grid <= 0) return UNREAL_ECUDA;
  if (mom != nullptr) rmsprop_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(var, rms, mom, grad, p, sumsq, a, grad_norm, lr_dev);
  else rmsprop_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(var, rms, nullptr, grad, p, sumsq, a, grad_norm, lr_dev);
  UNREAL_LAUNCH_CHECK("rmsprop_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_rmsprop_update(float* var, float* rms, float* mom, const float* grad, int64_t p,
                                     const double* sumsq, float grad_scale, float lr, float decay, float momentum,
                                     float eps, float clip_norm, float* grad_norm, void* stream) {
  return rmsprop_launch(var, rms, mom, grad, p, sumsq, grad_scale, lr, nullptr, decay, momentum, eps, clip_norm, grad_norm,
                        stream);
}

extern "C" int unreal_rmsprop_update_dlr(float* var, float* rms, float* mom, const float* grad, int64_t p,
                                         const double* sumsq, float grad_scale, const float* lr_dev, float decay,
                                         float momentum, float eps, float clip_norm, float* grad_norm, void* stream) {
  UNREAL_REQUIRE(lr_dev != nullptr, "unreal_rmsprop_update_dlr: lr_dev is null");
  return rmsprop_launch(var, rms, mom, grad, p, sumsq, grad_scale, 0.f, lr_dev, decay, momentum, eps, clip_norm, grad_norm,
                        stream);
}
