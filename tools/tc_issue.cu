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

template <bool UNI>
__global__ void __launch_bounds__(640) k(int mode, int N, int n_acc, int issuers, int fresh, int reps, long long* out, int side, int nside, int lbo_rows, volatile int* stop) {
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
        db[s] = smem_desc(b0 + s * 2 * lbo_rows * 16, lbo_rows * 16, 128);
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
      atomicAdd((int*)stop, 1);
    }
  }
  if (warp >= 4 && warp < 4 + nside && side > 0) {
    // epilogue-like side traffic on TMEM columns 256..447 of this warp's lane quarter until the issuers are done
    const uint32_t taddr = tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0; int it = 0;
    while (*stop < issuers) {
      uint32_t v[32];
      tmem_ld32(taddr + (it & 3) * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i];
      if (side == 2) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = v[i] + v[i + 16];
        tmem_st16(taddr + 128 + (it & 3) * 16, w);
        tmem_st_wait();
      }
      ++it;
    }
    if (lane == 0) out[8 + (warp - 4)] = it + (acc == 12345u);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 256); int* stopf; cudaMalloc(&stopf, 4);
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
            cudaMemset(d, 0, 256);
            cudaMemset(stopf, 0, 4);
            if (uni) k<true><<<1, 640, 196608>>>(mode, N, n_acc, issuers, fresh, reps, d, 0, 0, N, stopf);
            else k<false><<<1, 640, 196608>>>(mode, N, n_acc, issuers, fresh, reps, d, 0, 0, N, stopf);
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
  // contention: uniform-issue TS MMAs against epilogue-like TMEM traffic (what the N-split chain runs concurrently)
  for (int N : {64, 128})
    for (int issuers : {1, 2})
      for (int lbo_rows : {N, 128})
        for (int side : {0, 1, 2})
          for (int nside : {8, 16}) {
            if (lbo_rows < N || (side == 0 && nside == 16)) continue;
            cudaMemset(d, 0, 256); cudaMemset(stopf, 0, 4);
            k<true><<<1, 640, 196608>>>(2, N, 2, issuers, 0, 256, d, side, nside, lbo_rows, stopf);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[32]; cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
            const long long tot = h[0] > h[2] ? h[0] : h[2];
            long long its = 0; for (int w = 0; w < nside; ++w) its += h[8 + w];
            printf("contention: TS N=%3d (B image %3d rows) issuers=%d + %2d warps of %s: %6.1f cycles per MMA (all issuers), side traffic %.1f B/cycle TMEM read\n",
                   N, lbo_rows, issuers, side ? nside : 0, side == 0 ? "nothing      " : side == 1 ? "tcgen05.ld   " : "tcgen05.ld+st", (double)tot / (256 * 8 * issuers),
                   (double)its * 4096 / tot);
          }
  return 0;
}
