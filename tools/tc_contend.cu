// Which epilogue activity slows a concurrent stream of SS-mode tcgen05.mma (M=128 N=128 K=16, bf16)?
// One CTA: thread 0 issues MMAs back to back; W other warps run one kind of memory traffic.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../nefes_b200/csrc/tc05.cuh"
using namespace tc05;
__device__ __forceinline__ void sts128(void* p, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}

__global__ void __launch_bounds__(576) k(int kind, int nwarps, int ts, int reps, uint4* gbuf, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const int N = 128;
    const uint32_t idesc = idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    uint64_t da[8], db[8];
    for (int s = 0; s < 8; ++s) { da[s] = smem_desc(a0 + s * 4096, 2048, 128); db[s] = smem_desc(b0 + s * 2 * N * 16, N * 16, 128); }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (ts) mma_ts(tmem, tmem + 384 + s * 8, db[s], idesc, 1);
        else mma_ss(tmem, da[s], db[s], idesc, 1);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
    stop = 1;
  } else if (warp >= 2 && warp < 2 + nwarps) {
    uint8_t* base = smem + 131072 + (warp - 2) * 2048;       // private 2 KB per warp (not an MMA operand)
    uint4 v = make_uint4(lane, warp, 1, 2);
    uint4 acc = make_uint4(0, 0, 0, 0);
    long long n = 0;
    uint4* gp = gbuf + ((size_t)blockIdx.x * 32 + warp) * 65536 + lane;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (kind == 1) sts128(base + i * 512 + lane * 16, v);                 // STS.128 conflict-free
        else if (kind == 2) { uint4 t = lds128(base + i * 16); acc.x += t.x; }   // LDS.128 broadcast
        else if (kind == 3) gp[((n * 4 + i) & 2047) * 32] = v;                                               // STG.128 512 B per warp
        else if (kind == 4) { sts128(base + i * 512 + lane * 16, v); fence_async_smem(); }
        else if (kind == 5) { uint32_t r32[32]; tmem_ld32(tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16), r32); tmem_ld_wait(); acc.y += r32[0]; }
      }
      ++n;
    }
    if (lane == 0) { out[1 + warp] = n * 4; out[40] = acc.x + acc.y; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  uint4* g; cudaMalloc(&g, (size_t)32 * 65536 * 16 * 2);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[6] = {"nothing", "STS.128", "LDS.128 bcast", "STG.128", "STS.128+fence.proxy.async", "tcgen05.ld x32"};
  for (int ts = 0; ts < 2; ++ts)
    for (int kind = 0; kind < 6; ++kind)
      for (int nw : {8, 16}) {
        if (kind == 0 && nw == 16) continue;
        cudaMemset(d, 0, 64 * 8);
        const int reps = 256;
        k<<<1, 576, 196608>>>(kind, kind == 0 ? 0 : nw, ts, reps, g, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[64]; cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
        long long ops = 0; for (int w = 2; w < 2 + nw; ++w) ops += h[1 + w];
        printf("%s MMA N=128 + %2d warps of %-26s: %.1f cycles/MMA ; side traffic %.2f warp-instr/cycle\n", ts ? "TS" : "SS", kind == 0 ? 0 : nw,
               names[kind], (double)h[0] / (reps * 8), (double)ops / h[0]);
      }
  return 0;
}
