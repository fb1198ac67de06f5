// Round-2 microbenchmark: what limits a stream of tcgen05.mma -- the issuing thread, the accumulator dependency, or the
// shared-memory operand layout?  One CTA on one SM, cycles by clock64.
//   mode 0  SS, SWIZZLE_NONE interleaved images (what the engine uses), bf16
//   mode 1  SS, SWIZZLE_128B K-major tiles (64 bf16 = 128 B per row, 8-row groups 1024 B apart), bf16
//   mode 2  TS (A in TMEM), B as mode 0
//   mode 3  SS, SWIZZLE_NONE, kind::tf32 (K = 8 per MMA)
// n_acc: the issuing thread rotates over n_acc independent accumulators (D = tmem + a*N columns);
// issuers: 1 or 2 threads (different warps), each with its own accumulator set;
// fresh: every MMA overwrites (accumulate = 0) instead of accumulating.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../nefes_b200/csrc/tc05.cuh"
using namespace tc05;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {       // K-major, 128B swizzle: SBO = 1024, LBO ignored (1)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <bool UNI>
__global__ void __launch_bounds__(128) k(int mode, int N, int n_acc, int issuers, int fresh, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if (UNI ? (uwarp < issuers) : (lane == 0 && warp < issuers)) {
    const uint32_t tmem_u = UNI ? __shfl_sync(0xffffffffu, tmem, 0) : tmem;
    const bool leader = UNI ? elect_one() : true;
    const uint32_t a0 = smem_u32(smem) + (UNI ? uwarp : warp) * 32768, b0 = smem_u32(smem + 65536);
    uint64_t da[8], db[8];
    for (int s = 0; s < 8; ++s) {
      if (mode == 1) {          // 8 K-steps = two 64-wide K blocks; +32 B per K step inside the 128 B swizzle row
        da[s] = desc_sw128(a0 + (s >> 2) * 16384 + (s & 3) * 32);
        db[s] = desc_sw128(b0 + (s >> 2) * (N * 128) + (s & 3) * 32);
      } else if (mode == 3) {   // tf32 interleaved: [K/4][rows][4] fp32, K = 8 per MMA = two 16-byte chunks
        da[s] = smem_desc(a0 + s * 4096, 2048, 128);
        db[s] = smem_desc(b0 + s * 2 * N * 16, N * 16, 128);
      } else {
        da[s] = smem_desc(a0 + s * 4096, 2048, 128);
        db[s] = smem_desc(b0 + s * 2 * N * 16, N * 16, 128);
      }
    }
    const uint32_t idesc = mode == 3 ? idesc_tf32(128, N) : idesc_bf16(128, N, 0, 0);
    const uint32_t dbase = tmem_u + (UNI ? uwarp : warp) * (n_acc * N);
    uint32_t dd[8];
    for (int s = 0; s < 8; ++s) dd[s] = dbase + (s & (n_acc - 1)) * N;
    const uint32_t accum = fresh ? 0u : 1u;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const uint32_t d = dd[s];
        if (leader) {
          if (mode == 2) mma_ts(d, tmem_u + 448 + s * 8, db[s], idesc, accum);
          else if (mode == 3) mma_ss_tf32(d, da[s], db[s], idesc, accum);
          else mma_ss(d, da[s], db[s], idesc, accum);
        }
      }
    }
    long long t1 = clock64();
    if (leader) mma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    long long t2 = clock64();
    if (leader) {
      out[warp * 2] = t2 - t0;
      out[warp * 2 + 1] = t1 - t0;     // issue time only
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[4] = {"SS interleaved bf16", "SS swizzle128 bf16 ", "TS (A in TMEM) bf16", "SS interleaved tf32"};
  const int reps = 64;
  for (int uni = 0; uni < 2; ++uni)
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {64, 128, 256})
      for (int issuers : {1, 2})
        for (int n_acc : {1, 2, 4})
          for (int fresh : {0, 1}) {
            if (issuers * n_acc * N > 448) continue;
            if (fresh && n_acc > 1) continue;
            cudaMemset(d, 0, 64);
            if (uni) k<true><<<1, 128, 196608>>>(mode, N, n_acc, issuers, fresh, reps, d);
            else k<false><<<1, 128, 196608>>>(mode, N, n_acc, issuers, fresh, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s (mode %d N %d)\n", cudaGetErrorString(e), mode, N); return 1; }
            long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            const long long tot = h[0] > h[2] ? h[0] : h[2];
            const int kk = mode == 3 ? 8 : 16;
            const double floor_c = 128.0 * N * kk / 4096.0 / (mode == 3 ? 0.5 : 1.0);   // tf32 runs at half the bf16 rate
            printf("%s %s N=%3d issuers=%d acc/issuer=%d %s: %6.1f cycles per MMA slot (all issuers: %6.1f per MMA), issue-only %6.1f, floor %5.1f\n",
                   uni ? "uniform-issue" : "lane0-branch ", names[mode], N, issuers, n_acc, fresh ? "overwrite " : "accumulate", (double)tot / (reps * 8),
                   (double)tot / (reps * 8 * issuers), (double)h[1] / (reps * 8), floor_c);
          }
  return 0;
}
