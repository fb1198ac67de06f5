// Caller-side NeRF-W colour loss of the stage-1 training step, forward and backward, as two launches
// (script/models/losses.py:96-132, NerfWLoss.forward with the transient head present):
//   loss = coef * ( 0.5 mean((rgb_coarse - t)^2) + mean((rgb_fine - t)^2 / (2 beta^2)) + 3 + mean(log beta)
//                   + lambda_u mean(transient_sigmas) )
// The torch expression of the same thing is ~45 tiny elementwise / reduction launches (forward + autograd), which is a
// launch-bound bubble between the fine compositing forward and backward; SURVEY.md 8f row 4 (training-side glue).
#include "common.cuh"

namespace nefes {

// acc[0..3] = sum (rgb0-t)^2, sum (rgb-t)^2/(2 beta^2), sum log beta, sum tsig; acc[4] = blocks done (as float bits)
__global__ void nerfw_loss_fwd_kernel(const float* __restrict__ rgb0, const float* __restrict__ rgb, const float* __restrict__ beta,
                                      const float* __restrict__ tsig, const float* __restrict__ target, int64_t N, int S,
                                      float coef, float lambda_u, float* __restrict__ acc, unsigned int* __restrict__ done,
                                      float* __restrict__ loss) {
  float c = 0.f, f = 0.f, b = 0.f, s = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    const float be = beta[n];
    const float inv = 1.f / (2.f * be * be);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float t = target[n * 3 + k];
      const float d0 = rgb0[n * 3 + k] - t, d1 = rgb[n * 3 + k] - t;
      c += d0 * d0;
      f += d1 * d1 * inv;
    }
    b += logf(be);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * S; i += stride) s += tsig[i];
  __shared__ float red[4][8];
  c = warp_sum(c); f = warp_sum(f); b = warp_sum(b); s = warp_sum(s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = c; red[1][warp] = f; red[2][warp] = b; red[3][warp] = s; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {                     // last block: all partial sums are in
      __threadfence();
      const float a0 = atomicAdd(acc + 0, 0.f), a1 = atomicAdd(acc + 1, 0.f), a2 = atomicAdd(acc + 2, 0.f), a3 = atomicAdd(acc + 3, 0.f);
      const float n3 = 3.f * (float)N;
      *loss = coef * (0.5f * a0 / n3 + a1 / n3 + 3.f + a2 / (float)N + lambda_u * a3 / ((float)N * (float)S));
    }
  }
}

// g = upstream cotangent of the scalar loss (device pointer)
__global__ void nerfw_loss_bwd_kernel(const float* __restrict__ rgb0, const float* __restrict__ rgb, const float* __restrict__ beta,
                                      const float* __restrict__ target, const float* __restrict__ g, int64_t N, int S, float coef,
                                      float lambda_u, float* __restrict__ d_rgb0, float* __restrict__ d_rgb,
                                      float* __restrict__ d_beta, float* __restrict__ d_tsig) {
  const float gs = g[0] * coef;
  const float n3 = 3.f * (float)N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    const float be = beta[n];
    const float ib2 = 1.f / (be * be);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float t = target[n * 3 + k];
      const float d0 = rgb0[n * 3 + k] - t, d1 = rgb[n * 3 + k] - t;
      d_rgb0[n * 3 + k] = gs * d0 / n3;
      d_rgb[n * 3 + k] = gs * d1 * ib2 / n3;
      q += d1 * d1;
    }
    d_beta[n] = gs * (-q * ib2 / be / n3 + 1.f / (be * (float)N));
  }
  const float ds = gs * lambda_u / ((float)N * (float)S);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * S; i += stride) d_tsig[i] = ds;
}

}  // namespace nefes

extern "C" {

int nefes_nerfw_loss_fwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* transient_sigmas,
                         const float* target, int64_t N, int S, float coef, float lambda_u, float* scratch8, float* loss,
                         void* stream) {
  NEFES_REQUIRE(rgb_coarse && rgb_fine && beta && transient_sigmas && target && scratch8 && loss, NEFES_EINVAL,
                "nefes_nerfw_loss_fwd: null pointer");
  NEFES_REQUIRE(N >= 1 && S >= 1, NEFES_EINVAL, "nefes_nerfw_loss_fwd: bad shape N=%lld S=%d", (long long)N, S);
  cudaStream_t st = (cudaStream_t)stream;
  NEFES_CUDA(cudaMemsetAsync(scratch8, 0, 8 * sizeof(float), st));
  const int blocks = (int)nefes::ceil_div(N * S, 256 * 8) < 592 ? (int)nefes::ceil_div(N * S, 256 * 8) : 592;
  nefes::nerfw_loss_fwd_kernel<<<blocks, 256, 0, st>>>(rgb_coarse, rgb_fine, beta, transient_sigmas, target, N, S, coef, lambda_u,
                                                       scratch8, reinterpret_cast<unsigned int*>(scratch8 + 4), loss);
  NEFES_CHECK_LAUNCH("nerfw_loss_fwd");
  return NEFES_OK;
}

int nefes_nerfw_loss_bwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* target,
                         const float* d_loss, int64_t N, int S, float coef, float lambda_u, float* d_rgb_coarse,
                         float* d_rgb_fine, float* d_beta, float* d_transient_sigmas, void* stream) {
  NEFES_REQUIRE(rgb_coarse && rgb_fine && beta && target && d_loss && d_rgb_coarse && d_rgb_fine && d_beta && d_transient_sigmas,
                NEFES_EINVAL, "nefes_nerfw_loss_bwd: null pointer");
  NEFES_REQUIRE(N >= 1 && S >= 1, NEFES_EINVAL, "nefes_nerfw_loss_bwd: bad shape N=%lld S=%d", (long long)N, S);
  const int blocks = (int)nefes::ceil_div(N * S, 256 * 8) < 592 ? (int)nefes::ceil_div(N * S, 256 * 8) : 592;
  nefes::nerfw_loss_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rgb_coarse, rgb_fine, beta, target, d_loss, N, S, coef,
                                                                        lambda_u, d_rgb_coarse, d_rgb_fine, d_beta,
                                                                        d_transient_sigmas);
  NEFES_CHECK_LAUNCH("nerfw_loss_bwd");
  return NEFES_OK;
}

}  // extern "C"
