// Times back-to-back tcgen05.mma (M=128, N=128, K=16, bf16) for operand layouts used by the engine.
//   mode 0: SS, A K-major  x B K-major  (interleaved/no-swizzle)        -> forward / dgrad
//   mode 1: SS, A MN-major x B MN-major (interleaved/no-swizzle)        -> wgrad
//   mode 2: TS, A in TMEM  x B K-major
//   mode 3: SS, A K-major x B K-major but N=64 ; mode 4: SS MN x MN, N=64
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../nefes_b200/csrc/tc05.cuh"
using namespace tc05;
__global__ void __launch_bounds__(128) k(int mode, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const int N = (mode == 3 || mode == 4) ? 64 : (mode >= 6 ? 256 : 128);
    const int alt = (mode == 5 || mode == 7) ? 1 : 0;
    const bool mn = (mode == 1 || mode == 4);
    const uint32_t idesc = idesc_bf16(128, N, mn, mn);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
    uint64_t da[8], db[8];
    for (int s = 0; s < 8; ++s) {
      if (mn) { da[s] = smem_desc(a0 + s * 256, 128, 2048); db[s] = smem_desc(b0 + s * 256, 128, 2048); }
      else { da[s] = smem_desc(a0 + s * 4096, 2048, 128); db[s] = smem_desc(b0 + (s & 3) * 2 * N * 16, N * 16, 128); }
    }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (mode == 2) mma_ts(tmem, tmem + 256 + s * 8, db[s], idesc, 1);
        else mma_ss(tmem + (alt ? (s & 1) * 256 : 0), da[s], db[s], idesc, 1);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int mode = 0; mode < 8; ++mode) {
    const int reps = 64;
    k<<<1, 128, 66 * 1024>>>(mode, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("mode %d: %s, %lld cycles for %d MMAs -> %.1f cycles/MMA (ideal %d)\n", mode, cudaGetErrorString(e), h, reps * 8,
           (double)h / (reps * 8), (mode == 3 || mode == 4) ? 32 : (mode >= 6 ? 128 : 64));
  }
  return 0;
}
