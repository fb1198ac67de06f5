// Issue rates of the epilogue's arithmetic, 8 independent chains per thread, W warps per SM:
//   0 cvt.rn.relu.bf16x2.f32 (F2FP)   1 add.f32x2 (FADD2)   2 fma.f32   3 set.gt bf16x2 + and (HSET2 + LOP3)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int KIND>
__global__ void k(int iters, uint32_t* out, long long* cyc) {
  float f[8]; uint32_t r[8]; uint64_t p[8];
  for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3f + i; r[i] = threadIdx.x + i; p[i] = ((uint64_t)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 1.f); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(__uint_as_float(r[i])));
      if (KIND == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]));
      if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
      if (KIND == 3) asm volatile("{.reg .b32 m; set.gt.u32.bf16x2 m, %0, %1; and.b32 %0, %0, m;}" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) acc ^= r[i] ^ __float_as_uint(f[i]) ^ (uint32_t)p[i] ^ (uint32_t)(p[i] >> 32);
  out[threadIdx.x] = acc;
}
template <int KIND> void run(const char* name, uint32_t* o, long long* c) {
  for (int warps : {4, 8, 16}) {
    k<KIND><<<1, warps * 32>>>(2000, o, c);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double n_per_sched = 2000.0 * 8 * (KIND == 3 ? 2 : 1) * (warps / 4.0);
    printf("%-34s %2d warps: %.2f cycles per warp-instruction per scheduler\n", name, warps, (double)h / n_per_sched);
  }
}
int main() {
  uint32_t* o; long long* c; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 64);
  run<0>("cvt.rn.relu.bf16x2.f32 (F2FP)", o, c);
  run<1>("add.f32x2 (FADD2)", o, c);
  run<2>("fma.f32 (FFMA)", o, c);
  run<3>("set.gt.bf16x2 + and (HSET2+LOP3)", o, c);
  return 0;
}
