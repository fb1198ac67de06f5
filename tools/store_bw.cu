// What HBM write bandwidth can 148 persistent CTAs reach with 32 KB bulk stores (cp.async.bulk shared -> global), as a
// function of how many stores each CTA keeps in flight?  (The forward chain keeps <= 2: one per tile of its pair.)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int DEPTH>
__global__ void k(uint8_t* dst, int n_chunks) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < DEPTH * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    int j = 0;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x, ++j) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * 32768),
                   "r"(smem_u32(smem + (j % DEPTH) * 32768)), "r"(32768) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
template <int DEPTH> void run(uint8_t* d, int n) {
  cudaFuncSetAttribute(k<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * 32768);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); k<DEPTH><<<148, 128, DEPTH * 32768>>>(d, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  printf("32 KB bulk stores, %d in flight per CTA, 148 CTAs: %.3f ms for %.2f GB -> %.0f GB/s (%s)\n", DEPTH, best, n * 32768.0 / 1e9,
         n * 32768.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int n = 81920;   // 2.68 GB, the forward chain's saved bytes at the bench shape
  uint8_t* d; cudaMalloc(&d, (size_t)n * 32768);
  run<1>(d, n); run<2>(d, n); run<4>(d, n); run<6>(d, n);
  return 0;
}
