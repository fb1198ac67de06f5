// Write-only HBM bandwidth on this GPU, by the store forms the forward chain can use for its saved activation images:
// st.global.v4 from registers, cp.async.bulk shared -> global (16 KB per store, as the chain's per-block store), the same
// with an L2 evict_first hint, cudaMemsetAsync, and a plain copy for reference.  The forward chain with saves writes
// 2.65 GB per launch at ~4.7 TB/s whichever store form it uses; this says what the ceiling for a pure write stream is.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/hbm_write.cu -o tools/hbm_write && tools/hbm_write
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_stg(uint4* dst, size_t n16) {
  const uint4 v = make_uint4(threadIdx.x, 1u, 2u, 3u);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__global__ void k_copy(uint4* dst, const uint4* src, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
template <int HINT>
__global__ void k_bulk(uint8_t* dst, size_t n_chunks, uint32_t chunk) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (uint32_t i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t pol = 0;
    if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      if (HINT)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + c * chunk), "r"(smem_u32(sm)), "r"(chunk), "l"(pol) : "memory");
      else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk), "r"(smem_u32(sm)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <class F> float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)2654 << 20;   // what one fine forward launch saves at 6144 rays
  uint8_t *d, *s;
  if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMalloc(&s, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(s, 1, bytes);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("# %s, %d SMs, %.2f GB per pass\n", p.name, sms, bytes / 1e9);
  auto report = [&](const char* name, float ms, double factor) { printf("%-46s %8.3f ms  %8.1f GB/s\n", name, ms, factor * bytes / ms / 1e6); };
  for (int bps : {1, 2, 4, 8})
    for (int th : {256, 1024}) {
      char nm[96];
      snprintf(nm, sizeof nm, "st.global.v4  %d CTA/SM x %d threads", bps, th);
      report(nm, time_ms([&] { k_stg<<<sms * bps, th>>>((uint4*)d, bytes / 16); }, 10), 1.0);
    }
  for (uint32_t chunk : {16384u, 32768u})
    for (int bps : {1, 2, 4}) {
      char nm[96];
      cudaFuncSetAttribute(k_bulk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      snprintf(nm, sizeof nm, "bulk s2g %u B  %d CTA/SM", chunk, bps);
      report(nm, time_ms([&] { k_bulk<0><<<sms * bps, 128, chunk>>>(d, bytes / chunk, chunk); }, 10), 1.0);
      snprintf(nm, sizeof nm, "bulk s2g %u B  %d CTA/SM  evict_first", chunk, bps);
      report(nm, time_ms([&] { k_bulk<1><<<sms * bps, 128, chunk>>>(d, bytes / chunk, chunk); }, 10), 1.0);
    }
  report("cudaMemsetAsync", time_ms([&] { cudaMemsetAsync(d, 0, bytes); }, 10), 1.0);
  report("copy kernel (read + write counted)", time_ms([&] { k_copy<<<sms * 8, 1024>>>((uint4*)d, (const uint4*)s, bytes / 16); }, 10), 2.0);
  report("cudaMemcpyAsync d2d (read + write counted)", time_ms([&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); }, 10), 2.0);
  return 0;
}
