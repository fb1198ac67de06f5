// Round-2 microbenchmark: what does ONE iteration of an MMA-issuer loop cost, segment by segment, and who makes it slow?
// One CTA.  Warp 0 is the issuer (converged, elected lane issues):
//     loop { [a passing mbarrier wait] ; tcgen05.fence ; 4 or 8 TS MMAs (N = 64) ; tcgen05.commit ; __syncwarp }
// stamped with clock64 after every segment.  Side warps (the "other roles" of a real kernel) do one of:
//     side 0  nothing (exit)
//     side 1  spin in mbarrier.try_wait on a barrier that never completes            (idle epilogue / producer warps)
//     side 2  spin in mbarrier.test_wait (no hardware suspend)
//     side 3  tcgen05.ld + pack + tcgen05.st + fence loop                            (busy epilogue warps)
//     side 4  nanosleep(200) polling of test_wait
//     side 5  a 48 KB straight-line FFMA body in a loop                                (instruction-cache pressure)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../nefes_b200/csrc/tc05.cuh"
using namespace tc05;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(608) k(int side, int nside, int mmas, int with_wait, int reps, long long* out, volatile int* stop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_never, bar_pass, bar_final;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_never, 1); mbar_init(&bar_pass, 1); mbar_init(&bar_final, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) mbar_arrive(&bar_pass);      // phase 0 of bar_pass is complete: wait(parity 0) passes at once
  __syncthreads();
  if (warp == 0) {
    const uint32_t tm = uni(tmem);
    const bool leader = elect_one();
    const uint32_t b0 = smem_u32(smem);
    const uint32_t idesc = idesc_bf16(128, 64, 0, 0);
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    uint32_t done = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (with_wait) mbar_wait(&bar_pass, 0);
      const long long t1 = clock64();
      tc_fence_after();
      const long long t2 = clock64();
      if (leader) {
        const uint64_t db0 = smem_desc(b0, 2048, 128);
        for (int kk = 0; kk < mmas; ++kk) mma_ts(tm, tm + 448 + (kk & 7) * 8, db0 + (uint64_t)((kk & 7) * 256), idesc, kk > 0 ? 1u : 0u);
      }
      const long long t3 = clock64();
      if (leader) mma_commit(&bar_done);
      const long long t4 = clock64();
      __syncwarp();
      const long long t5 = clock64();
      if (with_wait == 2) { mbar_wait(&bar_done, done & 1); ++done; }     // wait for the MMAs to retire (commit latency)
      const long long t6 = clock64();
      acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; acc[5] += t6 - t5;
    }
    if (leader) mma_commit(&bar_final);
    mbar_wait(&bar_final, 0);
    if (lane == 0) { for (int i = 0; i < 6; ++i) out[i] = acc[i]; atomicExch((int*)stop, 1); }
  } else if (warp <= nside && side > 0) {
    const uint32_t taddr = tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    while (*stop == 0) {
      if (side == 1) { for (int i = 0; i < 64 && !mbar_try_wait(&bar_never, 0); ++i) {} }
      else if (side == 2) { for (int i = 0; i < 64 && !mbar_test_wait(&bar_never, 0); ++i) {} }
      else if (side == 5) {
        float x = __int_as_float(sink | 0x3f800000u);
#define F1(c) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x) : "f"((float)(c)));
#define F8(c) F1(c) F1(c + 1) F1(c + 2) F1(c + 3) F1(c + 4) F1(c + 5) F1(c + 6) F1(c + 7)
#define F64(c) F8(c) F8(c + 8) F8(c + 16) F8(c + 24) F8(c + 32) F8(c + 40) F8(c + 48) F8(c + 56)
#define F512(c) F64(c) F64(c + 64) F64(c + 128) F64(c + 192) F64(c + 256) F64(c + 320) F64(c + 384) F64(c + 448)
        F512(0) F512(512) F512(1024) F512(1536) F512(2048) F512(2560)
        sink += __float_as_int(x) & 1;
      }
      else if (side == 4) { for (int i = 0; i < 8 && !mbar_test_wait(&bar_never, 0); ++i) __nanosleep(200); }
      else {
        uint32_t v[32], w[16];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack_bf16(fmaxf(__uint_as_float(v[2 * i]), 0.f), fmaxf(__uint_as_float(v[2 * i + 1]), 0.f));
        tmem_st16(taddr + 64, w);
        tmem_st_wait();
        tc_fence_before();
        sink += w[0];
      }
    }
    if (sink == 0x12345u) out[7] = sink;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 256); int* stopf; cudaMalloc(&stopf, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const char* sides[6] = {"none", "try_wait spin", "test_wait spin", "tcgen05.ld/st loop", "nanosleep poll", "48 KB code loop"};
  const int reps = 256;
  for (int with_wait : {2})
    for (int mmas : {4, 8})
      for (int side : {0, 3, 5})
        for (int nside : {4, 8, 18}) {
          if (side == 0 && nside != 8) continue;
          cudaMemset(d, 0, 256); cudaMemset(stopf, 0, 4);
          k<<<1, 608, 65536>>>(side, nside, mmas, with_wait, reps, d, stopf);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
          printf("%s, %d MMAs/iter, side = %2d warps of %-18s: wait %5.0f | fence %5.0f | issue %5.0f | commit %5.0f | syncwarp %5.0f | retire-wait %5.0f   (cycles per iteration; pipe floor %d)\n",
                 with_wait == 0 ? "no wait      " : with_wait == 1 ? "passing wait " : "wait + retire", mmas, side ? nside : 0, sides[side],
                 (double)h[0] / reps, (double)h[1] / reps, (double)h[2] / reps, (double)h[3] / reps, (double)h[4] / reps, (double)h[5] / reps, mmas * 32);
        }
  return 0;
}
