// Probe: does cuTensorMapEncodeTiled accept OVERLAPPING rows (dim-1 stride 64 B < dim-0 extent 128 B) and
// how does the conv2 box {64 (4 px x 16 c), 9 X, 9 Y, 1} land under CU_TENSOR_MAP_SWIZZLE_128B?  Fills h1 [20,20,16] with unique 16-bit ids, loads the conv2 box and
// dumps shared memory; the host decodes where every element went.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t* out, int bx, int by, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t dst = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem + (dst - (uint32_t)__cvta_generic_to_shared(smem)))[i] = 0xFFFF;
  asm volatile("fence.proxy.async.shared::cta;");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes));
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(sb), "r"(0), "r"(bx), "r"(by), "r"(0) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(sb));
  __syncthreads();
  const uint16_t* s16 = reinterpret_cast<const uint16_t*>(smem + (dst - (uint32_t)__cvta_generic_to_shared(smem)));
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) out[i] = s16[i];
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  std::vector<uint16_t> h(6400);
  for (int i = 0; i < 6400; ++i) h[i] = (uint16_t)i;
  uint16_t *d, *o;
  cudaMalloc(&d, 6400 * 2); cudaMalloc(&o, 16384);
  cudaMemcpy(d, h.data(), 6400 * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  const int dy = 1;
  cuuint64_t dims[4] = {64, 9, 10, 1};
  cuuint64_t strides[3] = {64, 1280, 12800};
  cuuint32_t box[4] = {64, 9, 9, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d + dy * 320, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  if (r) return 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  for (int by = 0; by < 2; ++by) {
    probe<<<1, 128, 20480>>>(tm, o, 0, by, 64 * 9 * 9 * 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("by=%d sync: %s\n", by, cudaGetErrorString(e));
    std::vector<uint16_t> s(8192);
    cudaMemcpy(s.data(), o, 16384, cudaMemcpyDeviceToHost);
    // row pix = Y*9+X (128 B), logical 16-byte chunk j = (kx, c-half) 0..7, phys chunk = j ^ (pix & 7)
    int bad_sw = 0, bad_none = 0;
    for (int pix = 0; pix < 81; ++pix) for (int j = 0; j < 8; ++j) for (int e2 = 0; e2 < 8; ++e2) {
      int Y = pix / 9, X = pix % 9, d0 = j * 8 + e2;
      int src = ((2 * (Y + by) + dy) * 20 + 2 * X) * 16 + d0;
      if (s[pix * 64 + ((j ^ (pix & 7)) * 8) + e2] != (uint16_t)src) ++bad_sw;
      if (s[pix * 64 + j * 8 + e2] != (uint16_t)src) ++bad_none;
    }
    printf("  mismatches: 128B-swizzle %d, no-swizzle %d\n", bad_sw, bad_none);
  }
  return 0;
}
