// K3 / K4: discounted-return scans of the UNREAL targets (sm_100a).
//
//  K3  unreal_nstep_returns    Trainer._process_base reverse loop (train/trainer.py:298-324)
//      unreal_sequence_returns Trainer._process_vr   reverse loop (train/trainer.py:394-403)
//  K4  unreal_pc_targets       Trainer._process_pc   reverse loop (train/trainer.py:352-372)
//
// Time-major tensors ([T,N], [T,N,20,20]) put consecutive envs in consecutive lanes, so each
// thread owns one env (K3) or one float4 of one env's 20x20 map (K4), walks time backwards in
// registers and every load/store is a fully coalesced 128-byte line; the T loads of a thread
// are independent of the recurrence and are issued ahead of it in chunks.  The recurrence is
// written with explicit mul/add roundings (no FMA contraction) so fp32 results are bit-equal
// to numpy's float32 evaluation of the reference expression `r + gamma * R`.
// Env-major sequences gathered from the replay ring ([N,L]) use one warp per sequence and a
// shuffle suffix-scan instead.
#include "common.cuh"

namespace unreal {

constexpr int kChunk = 10;

__global__ void __launch_bounds__(128) nstep_returns_kernel(const float* __restrict__ r, const float* __restrict__ v,
                                                            const uint8_t* __restrict__ term,
                                                            const float* __restrict__ boot, float gamma,
                                                            float* __restrict__ out_R, float* __restrict__ out_adv,
                                                            int T, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float R = boot[n];
  for (int hi = T; hi > 0; hi -= kChunk) {
    const int lo = hi - kChunk;
    float rr[kChunk], vv[kChunk];
    uint8_t tt[kChunk];
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
      const int t = lo + k;
      if (t >= 0) {
        const size_t o = (size_t)t * N + n;
        rr[k] = __ldcs(r + o);
        tt[k] = term ? __ldcs(term + o) : (uint8_t)0;
        vv[k] = v ? __ldcs(v + o) : 0.f;
      }
    }
#pragma unroll
    for (int k = kChunk - 1; k >= 0; --k) {
      const int t = lo + k;
      if (t >= 0) {
        const size_t o = (size_t)t * N + n;
        R = tt[k] ? 0.f : R;                          // the terminal step bootstraps from 0 (:298-300)
        R = __fadd_rn(rr[k], __fmul_rn(gamma, R));    // R = ri + gamma * R   (:314)
        __stcs(out_R + o, R);
        if (out_adv) __stcs(out_adv + o, __fsub_rn(R, vv[k]));  // adv = R - Vi   (:315)
      }
    }
  }
}

// one warp per sequence, lanes = time; R_i = r_i + gamma * R_{i+1}, R_len = boot
__global__ void __launch_bounds__(128) sequence_returns_kernel(const float* __restrict__ r,
                                                               const int32_t* __restrict__ len,
                                                               const float* __restrict__ boot, float gamma,
                                                               float* __restrict__ out, int N, int L) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N) return;
  const int n_t = len ? len[warp] : L;
  float x = (lane < n_t) ? r[(size_t)warp * L + lane] : 0.f;
  if (lane == n_t - 1) x = __fadd_rn(x, __fmul_rn(gamma, boot[warp]));
  float g = gamma;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float y = __shfl_down_sync(0xffffffffu, x, d);
    if (lane + d < 32) x = __fadd_rn(x, __fmul_rn(g, y));
    g = __fmul_rn(g, g);
  }
  if (lane < L) out[(size_t)warp * L + lane] = (lane < n_t) ? x : 0.f;
}

constexpr int kPcVec = UNREAL_PC_CELLS * UNREAL_PC_CELLS / 4;  // 100 float4 per map

// One thread per float4 column of the [T, N*100] matrix, launched non-persistently in column
// order (a persistent contiguous-quota grid measured 15% slower on B200: dynamic CTA dispatch
// balances better).  Loads run one register chunk ahead of the recurrence so HBM reads stay in
// flight while the previous chunk is scanned and stored.
constexpr int kPcChunk = 5;
constexpr int kPcThreads = 128;

__device__ __forceinline__ void pc_load_chunk(float4 (&p)[kPcChunk], uint8_t (&tt)[kPcChunk],
                                              const float4* __restrict__ pc, const uint8_t* __restrict__ term,
                                              size_t per_t, size_t idx, int n, int N, int lo) {
#pragma unroll
  for (int k = 0; k < kPcChunk; ++k) {
    const int t = lo + k;
    if (t >= 0) {
      p[k] = __ldcs(pc + (size_t)t * per_t + idx);
      tt[k] = term ? term[(size_t)t * N + n] : (uint8_t)0;
    }
  }
}

__global__ void __launch_bounds__(kPcThreads) pc_targets_kernel(const float4* __restrict__ pc,
                                                                const uint8_t* __restrict__ term,
                                                                const int32_t* __restrict__ len,
                                                                const float4* __restrict__ boot, float g,
                                                                float4* __restrict__ tgt, int T, int N) {
  const size_t per_t = (size_t)N * kPcVec;   // columns = N * 100 float4
  const size_t idx = (size_t)blockIdx.x * kPcThreads + threadIdx.x;
  if (idx >= per_t) return;
  const int n = (int)(idx / kPcVec);
  const int n_t = len ? max(0, min(len[n], T)) : T;   // a not-ready ring reports len 0 -> n_batch -1: clamp, never store at t = -1
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cur[kPcChunk], nxt[kPcChunk];
  uint8_t tcur[kPcChunk], tnxt[kPcChunk];
  pc_load_chunk(cur, tcur, pc, term, per_t, idx, n, N, n_t - kPcChunk);
  float4 R = boot[idx];
  for (int t = T - 1; t >= n_t; --t) __stcs(tgt + (size_t)t * per_t + idx, zero);
  for (int hi = n_t; hi > 0; hi -= kPcChunk) {
    const int lo = hi - kPcChunk;
    if (lo > 0) pc_load_chunk(nxt, tnxt, pc, term, per_t, idx, n, N, lo - kPcChunk);
#pragma unroll
    for (int k = kPcChunk - 1; k >= 0; --k) {
      const int t = lo + k;
      if (t >= 0) {
        if (tcur[k]) R = zero;
        R.x = __fadd_rn(cur[k].x, __fmul_rn(g, R.x));   // pc_R = pixel_change + gamma_pc * pc_R   (:361)
        R.y = __fadd_rn(cur[k].y, __fmul_rn(g, R.y));
        R.z = __fadd_rn(cur[k].z, __fmul_rn(g, R.z));
        R.w = __fadd_rn(cur[k].w, __fmul_rn(g, R.w));
        __stcs(tgt + (size_t)t * per_t + idx, R);
      }
    }
#pragma unroll
    for (int k = 0; k < kPcChunk; ++k) { cur[k] = nxt[k]; tcur[k] = tnxt[k]; }
  }
}

}  // namespace unreal

using namespace unreal;

extern "C" int unreal_nstep_returns(const float* r, const float* v, const uint8_t* term, const float* boot,
                                    float gamma, float* out_R, float* out_adv, int t, int n, void* stream) {
  UNREAL_REQUIRE(t >= 0 && n >= 0, "unreal_nstep_returns: negative size");
  if (t == 0 || n == 0) return UNREAL_OK;  // empty batch: nothing to do, pointers may be null
  UNREAL_REQUIRE(r && boot && out_R, "unreal_nstep_returns: r, boot and out_R must be non-null");
  UNREAL_REQUIRE((v == nullptr) == (out_adv == nullptr), "unreal_nstep_returns: v and out_adv go together");
  nstep_returns_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(r, v, term, boot, gamma, out_R, out_adv, t, n);
  UNREAL_LAUNCH_CHECK("nstep_returns_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_sequence_returns(const float* r, const int32_t* len, const float* boot, float gamma,
                                       float* out_R, int n, int l, void* stream) {
  UNREAL_REQUIRE(n >= 0, "unreal_sequence_returns: negative size");
  UNREAL_REQUIRE(l >= 1 && l <= 32, "unreal_sequence_returns: sequence length %d not in 1..32", l);
  if (n == 0) return UNREAL_OK;
  UNREAL_REQUIRE(r && boot && out_R, "unreal_sequence_returns: r, boot and out_R must be non-null");
  long long threads = (long long)n * 32;
  sequence_returns_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(r, len, boot, gamma,
                                                                                            out_R, n, l);
  UNREAL_LAUNCH_CHECK("sequence_returns_kernel");
  return UNREAL_OK;
}

extern "C" int unreal_pc_targets(const float* pc, const uint8_t* term, const int32_t* len, const float* boot,
                                 float gamma_pc, float* tgt, int t, int n, void* stream) {
  UNREAL_REQUIRE(t >= 0 && n >= 0, "unreal_pc_targets: negative size");
  if (t == 0 || n == 0) return UNREAL_OK;
  UNREAL_REQUIRE(pc && boot && tgt, "unreal_pc_targets: pc, boot and tgt must be non-null");
  UNREAL_REQUIRE(aligned16(pc) && aligned16(boot) && aligned16(tgt), "unreal_pc_targets: buffers must be 16-byte aligned");
  const long long cols = (long long)n * kPcVec;
  pc_targets_kernel<<<(unsigned)((cols + kPcThreads - 1) / kPcThreads), kPcThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(pc), term, len, reinterpret_cast<const float4*>(boot), gamma_pc,
      reinterpret_cast<float4*>(tgt), t, n);
  UNREAL_LAUNCH_CHECK("pc_targets_kernel");
  return UNREAL_OK;
}
