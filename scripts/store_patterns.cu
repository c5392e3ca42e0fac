// Experiment: how does the WRITE pattern of a 347 MB frame batch (4096 x 84 672 B) affect
// achieved HBM bandwidth on B200?  Pure stores of a constant; no maze logic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/store_patterns scripts/store_patterns.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

constexpr int kF4 = 5292;     // float4 per frame
constexpr int kRow4 = 63;     // float4 per pixel row

// A: flat grid-stride fill
__global__ void flat_fill(float4* out, size_t n4) {
  float4 v = make_float4(1, 0, 0, 1);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(out + i, v);
}
// A2: flat, each CTA owns a contiguous slab
__global__ void slab_fill(float4* out, size_t n4) {
  size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  size_t b = (size_t)blockIdx.x * per, e = min(b + per, n4);
  float4 v = make_float4(1, 0, 0, 1);
  for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) __stcs(out + i, v);
}
// B: persistent CTA per env, stride gridDim; threads sweep the frame linearly
template <int CS>
__global__ void env_cta_linear(float4* out, int n) {
  float4 v = make_float4(1, 0, 0, 1);
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    float4* o = out + (size_t)e * kF4;
    for (int i = threadIdx.x; i < kF4; i += blockDim.x) { if (CS) __stcs(o + i, v); else o[i] = v; }
  }
}
// B2: as the product's variant 0: thread = (c4 = tid&63, rsub), rows rsub + R*j
__global__ void env_cta_rows(float4* out, int n) {
  float4 v = make_float4(1, 0, 0, 1);
  int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6, R = blockDim.x >> 6;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    float4* o = out + (size_t)e * kF4 + c4;
    if (c4 < kRow4)
      for (int r = rsub; r < 84; r += R) __stcs(o + r * kRow4, v);
  }
}
// C: CTA handles a contiguous block of envs [b*per, (b+1)*per)
__global__ void env_cta_block(float4* out, int n) {
  int per = (n + gridDim.x - 1) / gridDim.x;
  float4 v = make_float4(1, 0, 0, 1);
  for (int e = blockIdx.x * per; e < min(n, (blockIdx.x + 1) * per); ++e) {
    float4* o = out + (size_t)e * kF4;
    for (int i = threadIdx.x; i < kF4; i += blockDim.x) __stcs(o + i, v);
  }
}
// D: warp per env
__global__ void env_warp(float4* out, int n) {
  int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= n) return;
  float4 v = make_float4(1, 0, 0, 1);
  float4* o = out + (size_t)e * kF4;
  for (int i = lane; i < kF4; i += 32) __stcs(o + i, v);
}
// E: cluster-free "team": T consecutive CTAs share one env (each writes a 1/T slice), persistent
__global__ void env_team(float4* out, int n, int team) {
  int member = blockIdx.x % team, tm = blockIdx.x / team, teams = gridDim.x / team;
  int per = (kF4 + team - 1) / team;
  float4 v = make_float4(1, 0, 0, 1);
  for (int e = tm; e < n; e += teams) {
    float4* o = out + (size_t)e * kF4;
    for (int i = member * per + threadIdx.x; i < min(kF4, (member + 1) * per); i += blockDim.x) __stcs(o + i, v);
  }
}

template <typename F>
float time_it(F launch, int iters = 20) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch();
  std::vector<float> ts;
  for (int i = 0; i < iters; ++i) {
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  return ts[ts.size() / 2];
}

int main() {
  const int T = 8, N = 4096;   // T slices so successive launches touch different memory (>L2)
  size_t n4 = (size_t)N * kF4;
  float4* buf; cudaMalloc(&buf, n4 * 16 * T);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double mb = n4 * 16 / 1e6;
  int k = 0;
  auto slice = [&]() { return buf + (size_t)((k++) % T) * n4; };
  auto rep = [&](const char* name, int a, int b, float ms) {
    printf("{\"pattern\": \"%s\", \"p0\": %d, \"p1\": %d, \"us\": %.2f, \"gbs\": %.1f}\n", name, a, b, ms * 1e3, mb / ms);
    fflush(stdout);
  };
  for (int cps : {2, 4, 8, 16}) for (int th : {256, 512, 1024}) {
    if (cps * th > 2048) continue;
    rep("flat_fill", cps, th, time_it([&] { flat_fill<<<sms * cps, th>>>(slice(), n4); }));
    rep("slab_fill", cps, th, time_it([&] { slab_fill<<<sms * cps, th>>>(slice(), n4); }));
  }
  for (int th : {256, 512, 1024}) for (int cps : {1, 2, 4, 7, 8}) {
    if (cps * th > 2048) continue;
    rep("env_cta_linear_cs", cps, th, time_it([&] { env_cta_linear<1><<<sms * cps, th>>>(slice(), N); }));
    rep("env_cta_linear_wb", cps, th, time_it([&] { env_cta_linear<0><<<sms * cps, th>>>(slice(), N); }));
    rep("env_cta_rows", cps, th, time_it([&] { env_cta_rows<<<sms * cps, th>>>(slice(), N); }));
    rep("env_cta_block", cps, th, time_it([&] { env_cta_block<<<sms * cps, th>>>(slice(), N); }));
  }
  for (int th : {256, 512}) rep("env_cta_inorder", N, th, time_it([&] { env_cta_linear<1><<<N, th>>>(slice(), N); }));
  rep("env_warp", 0, 128, time_it([&] { env_warp<<<N / 4, 128>>>(slice(), N); }));
  for (int team : {2, 4, 7}) for (int th : {256, 512})
    rep("env_team", team, th, time_it([&] { env_team<<<(sms * (2048 / th) / team) * team, th>>>(slice(), N, team); }));
  return 0;
}
