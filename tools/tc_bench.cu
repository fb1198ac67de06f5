// Microbenchmarks that size the fused-chain design (run on one SM, cycles by clock64):
//   A. tcgen05.mma issue rate for the operand forms the engine uses (SS K-major, SS MN-major, TS = A in TMEM)
//   B. tcgen05.ld (TMEM -> registers) bandwidth with 4 and 8 warps, alone and while MMAs run
//   C. tcgen05.st bandwidth
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../nefes_b200/csrc/tc05.cuh"
using namespace tc05;

__global__ void __launch_bounds__(320) k(int mode, int N, int reps, int ld_warps, int with_mma, int ld_reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0 && (mode >= 0) && (with_mma || ld_warps == 0)) {
    const bool mn = (mode == 1);
    const uint32_t idesc = idesc_bf16(128, N, mn, mn);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    uint64_t da[8], db[8];
    for (int s = 0; s < 8; ++s) {
      if (mn) { da[s] = smem_desc(a0 + s * 256, 128, 2048); db[s] = smem_desc(b0 + s * 256, 128, 2048); }
      else { da[s] = smem_desc(a0 + s * 4096, 2048, 128); db[s] = smem_desc(b0 + (s & 3) * 2 * N * 16, N * 16, 128); }
    }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (mode == 2) mma_ts(tmem, tmem + 384 + s * 8, db[s], idesc, 1);
        else mma_ss(tmem, da[s], db[s], idesc, 1);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  if (warp >= 2 && warp < 2 + ld_warps) {
    const uint32_t taddr = tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < ld_reps; ++r) {
      uint32_t v[32];
      if (mode == -2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = acc + i;
        // st: 4 x8
        uint32_t w8[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int i = 0; i < 8; ++i) w8[i] = v[j * 8 + i];
          tmem_st8(taddr + (r & 3) * 32 + j * 8, w8);
        }
        tmem_st_wait();
      } else {
        tmem_ld32(taddr + (r & 3) * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
      }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[1 + warp] = t1 - t0; out[16] = acc; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 32 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  auto run = [&](int mode, int N, int reps, int ldw, int with_mma, long long* h, int ld_reps = 256) {
    cudaMemset(d, 0, 32 * 8);
    k<<<1, 320, 132 * 1024>>>(mode, N, reps, ldw, with_mma, ld_reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, d, 32 * 8, cudaMemcpyDeviceToHost);
  };
  long long h[32];
  const char* names[3] = {"SS K-major", "SS MN-major", "TS (A in TMEM)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {64, 128, 144, 192, 256}) {
      if (mode == 2 && N > 256) continue;
      run(mode, N, 64, 0, 1, h);
      printf("MMA %-15s M=128 N=%3d K=16: %.1f cycles/MMA (floor %d)\n", names[mode], N, (double)h[0] / 512, N / 2);
    }
  for (int ldw : {1, 4, 8}) {
    run(-1, 128, 256, ldw, 0, h);
    long long mx = 0; for (int w = 2; w < 2 + ldw; ++w) mx = h[1 + w] > mx ? h[1 + w] : mx;
    printf("tcgen05.ld x32, %d warps, alone: %.1f cycles per 4 KB warp-load -> %.1f B/cycle/SM\n", ldw, (double)mx / 256, 256.0 * ldw * 4096 / mx);
    run(-2, 128, 256, ldw, 0, h);
    mx = 0; for (int w = 2; w < 2 + ldw; ++w) mx = h[1 + w] > mx ? h[1 + w] : mx;
    printf("tcgen05.st 4x8, %d warps, alone: %.1f cycles per 4 KB warp-store -> %.1f B/cycle/SM\n", ldw, (double)mx / 256, 256.0 * ldw * 4096 / mx);
  }
  for (int mode : {0, 2}) {
    run(mode, 128, 512, 8, 1, h, 1024);
    long long mx = 0; for (int w = 2; w < 10; ++w) mx = h[1 + w] > mx ? h[1 + w] : mx;
    printf("concurrent: MMA %s N=128 %.1f cycles/MMA (%lld total) while 8 warps tcgen05.ld at %.1f B/cycle/SM (%lld total)\n", names[mode], (double)h[0] / 4096, h[0], 1024.0 * 8 * 4096 / mx, mx);
  }
  return 0;
}
