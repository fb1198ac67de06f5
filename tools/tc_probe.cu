// Probe for the tcgen05 building blocks used by the bf16 field kernels (run on a B200):
//   ./tc_probe <variant>
// 0 SS K-major (generic smem stores)   1 SS K-major via cp.async.bulk   2 TS (A in TMEM)
// 3 MN-major x MN-major (wgrad shape)   4 SS with LBO/SBO swapped (diagnostic)   5 SS, N = 16/64/144
// Prints max |err| against a double-precision host reference on bf16-rounded inputs.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../nefes_b200/csrc/tc05.cuh"

using namespace tc05;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ int g_timeout = 0;
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int i = 0; i < 20000000; ++i) if (mbar_try_wait(bar, parity)) return true;
  g_timeout = 1;
  return false;
}

// images are [C/8][ROWS][8] bf16.  A: ROWS=128 (M) ; B: ROWS=N.
// D[m][n] = sum_k A[m][k] B[n][k]      (variants 0,1,2,4,5)
// D[i][j] = sum_p X[p][i] Y[p][j]      (variant 3; X,Y images [C/8][128 pts][8])
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ imgA, const __nv_bfloat16* __restrict__ imgB,
                                                    float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_rows = 128, b_rows = (variant == 3) ? 128 : N;
  const int kdim = (variant == 3) ? 128 : K;                 // reduction length
  const int a_cols = (variant == 3) ? 128 : K, b_cols = (variant == 3) ? N : K;   // "c" extent of each image
  const uint32_t a_bytes = a_rows * a_cols * 2, b_bytes = b_rows * b_cols * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 1023) / 1024) * 1024;

  if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (variant == 1) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_load, a_bytes + b_bytes);
      bulk_g2s(sA, imgA, a_bytes, &bar_load);
      bulk_g2s(sB, imgB, b_bytes, &bar_load);
    }
    if (!wait_bounded(&bar_load, 0)) return;
  } else {
    if (variant != 2)
      for (uint32_t i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(imgA)[i];
    for (uint32_t i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(imgB)[i];
    fence_async_smem();
  }
  if (variant == 2) {
    // A -> TMEM columns [256, 256 + K/2): thread = row, packed bf16 pairs along k
    const uint32_t* rowsrc = reinterpret_cast<const uint32_t*>(imgA);
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
      for (int q = 0; q < 8; ++q) {
        const int k = 2 * (c0 + q);                          // element pair (k, k+1) of row tid
        v[q] = rowsrc[((k / 8) * a_rows * 8 + tid * 8 + (k % 8)) / 2];
      }
      tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    if (variant == 3) {
      // MN-major both: MN = c (channels), K = row (points).  SBO = stride between 8-channel groups = ROWS*16,
      // LBO = stride between 8-point groups = 128.  One MMA covers 16 points = 2 k-groups -> +256 B per step.
      const uint32_t idesc = idesc_bf16(128, N, 1, 1);
      for (int s = 0; s < kdim / 16; ++s) {
        const uint64_t da = smem_desc(smem_u32(sA) + s * 256, 128, 128 * 16);
        const uint64_t db = smem_desc(smem_u32(sB) + s * 256, 128, 128 * 16);
        mma_ss(tmem, da, db, idesc, s > 0);
      }
    } else {
      const uint32_t idesc = idesc_bf16(128, N, 0, 0);
      for (int s = 0; s < kdim / 16; ++s) {
        uint32_t lbo_a = a_rows * 16, sbo_a = 128, lbo_b = b_rows * 16, sbo_b = 128;
        if (variant == 4) { uint32_t t = lbo_a; lbo_a = sbo_a; sbo_a = t; t = lbo_b; lbo_b = sbo_b; sbo_b = t; }
        const uint64_t db = smem_desc(smem_u32(sB) + s * 2 * b_rows * 16, lbo_b, sbo_b);
        if (variant == 2) {
          mma_ts(tmem, tmem + 256 + s * 8, db, idesc, s > 0);
        } else {
          const uint64_t da = smem_desc(smem_u32(sA) + s * 2 * a_rows * 16, lbo_a, sbo_a);
          mma_ss(tmem, da, db, idesc, s > 0);
        }
      }
    }
    mma_commit(&bar_mma);
  }
  if (!wait_bounded(&bar_mma, 0)) return;
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int q = 0; q < 16; ++q) D[(size_t)tid * N + c0 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static int run(int variant, int N, int K) {
  const int M = 128;
  const int a_rows = 128, b_rows = (variant == 3) ? 128 : N;
  const int a_cols = (variant == 3) ? 128 : K, b_cols = (variant == 3) ? N : K;
  std::vector<float> A((size_t)a_rows * a_cols), B((size_t)b_rows * b_cols);
  srand(1234 + variant);
  for (auto& x : A) x = bf((rand() / (float)RAND_MAX) * 2 - 1);
  for (auto& x : B) x = bf((rand() / (float)RAND_MAX) * 2 - 1);
  std::vector<__nv_bfloat16> iA(A.size()), iB(B.size());
  for (int r = 0; r < a_rows; ++r) for (int c = 0; c < a_cols; ++c) iA[(size_t)(c / 8) * a_rows * 8 + r * 8 + c % 8] = __float2bfloat16(A[(size_t)r * a_cols + c]);
  for (int r = 0; r < b_rows; ++r) for (int c = 0; c < b_cols; ++c) iB[(size_t)(c / 8) * b_rows * 8 + r * 8 + c % 8] = __float2bfloat16(B[(size_t)r * b_cols + c]);
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, iA.size() * 2)); CK(cudaMalloc(&dB, iB.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, iA.data(), iA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, iB.data(), iB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
  const int smem = 2 * 1024 * ((int)(((size_t)a_rows * a_cols * 2 + 1023) / 1024) + (int)(((size_t)b_rows * b_cols * 2 + 1023) / 1024)) ;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_kernel<<<1, 128, smem < 200 * 1024 ? smem : 200 * 1024>>>(dA, dB, dD, N, K, variant);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int to = 0;
  CK(cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int)));
  std::vector<float> D((size_t)M * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    double ref = 0;
    if (variant == 3) for (int p = 0; p < 128; ++p) ref += (double)A[(size_t)p * a_cols + i] * B[(size_t)p * b_cols + j];
    else for (int k = 0; k < K; ++k) ref += (double)A[(size_t)i * K + k] * B[(size_t)j * K + k];
    double e = fabs(ref - D[(size_t)i * N + j]);
    if (!(e <= maxerr)) maxerr = e;
    if (fabs(ref) > maxref) maxref = fabs(ref);
  }
  printf("variant %d N=%d K=%d: timeout=%d max|err|=%.3e (max|ref|=%.2f) -> %s\n", variant, N, K, to, maxerr, maxref,
         (!to && maxerr < 1e-3 * maxref + 1e-4) ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return 0;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  if (variant == 5) { run(0, 16, 64); run(0, 64, 64); run(0, 144, 128); run(0, 128, 192); run(2, 64, 160); return 0; }
  run(variant, 128, 64);
  return 0;
}
