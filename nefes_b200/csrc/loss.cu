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

// Feature loss of the stage-2/3 step (script/models/losses.py:134-173, ColorFeatureFusionNerfWLoss: f_loss = nn.L1Loss /
// nn.MSELoss, reduction 'mean', of feat_fine [+ feat_coarse] [N,128] against the target feature map), one input per launch
// side: value = mean |a - t|   (mode 0)   or   mean (a - t)^2   (mode 1); with b != null the second term is added.
__global__ void feat_loss_fwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ t,
                                     int64_t n4, int mode, float inv_n, float* __restrict__ acc, unsigned int* __restrict__ done,
                                     float* __restrict__ loss) {
  float s = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 tv = t[i], av = a[i];
    const float d0 = av.x - tv.x, d1 = av.y - tv.y, d2 = av.z - tv.z, d3 = av.w - tv.w;
    s += mode == 0 ? fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3) : d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    if (b != nullptr) {
      const float4 bv = b[i];
      const float e0 = bv.x - tv.x, e1 = bv.y - tv.y, e2 = bv.z - tv.z, e3 = bv.w - tv.w;
      s += mode == 0 ? fabsf(e0) + fabsf(e1) + fabsf(e2) + fabsf(e3) : e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    }
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
    atomicAdd(acc, v);
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      __threadfence();
      *loss = atomicAdd(acc, 0.f) * inv_n;
    }
  }
}

__global__ void feat_loss_bwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ t,
                                     const float* __restrict__ g, int64_t n4, int mode, float inv_n, float4* __restrict__ d_a,
                                     float4* __restrict__ d_b) {
  const float gs = g[0] * inv_n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // torch: d|x|/dx = sign(x) with sign(0) = 0; d x^2/dx = 2x
  auto dv = [&](float x) { return mode == 0 ? (x > 0.f ? gs : (x < 0.f ? -gs : 0.f)) : 2.f * gs * x; };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 tv = t[i], av = a[i];
    d_a[i] = make_float4(dv(av.x - tv.x), dv(av.y - tv.y), dv(av.z - tv.z), dv(av.w - tv.w));
    if (b != nullptr) {
      const float4 bv = b[i];
      d_b[i] = make_float4(dv(bv.x - tv.x), dv(bv.y - tv.y), dv(bv.z - tv.z), dv(bv.w - tv.w));
    }
  }
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

int nefes_feat_loss_fwd(const float* feat_a, const float* feat_b, const float* target, int64_t n_elems, int mode, float* scratch2,
                        float* loss, void* stream) {
  NEFES_REQUIRE(feat_a && target && scratch2 && loss, NEFES_EINVAL, "nefes_feat_loss_fwd: null pointer");
  NEFES_REQUIRE(n_elems >= 4 && n_elems % 4 == 0 && (mode == 0 || mode == 1), NEFES_EINVAL,
                "nefes_feat_loss_fwd: n_elems=%lld must be a positive multiple of 4, mode 0 (L1) or 1 (MSE)", (long long)n_elems);
  cudaStream_t st = (cudaStream_t)stream;
  NEFES_CUDA(cudaMemsetAsync(scratch2, 0, 2 * sizeof(float), st));
  const int64_t n4 = n_elems / 4;
  const int blocks = (int)nefes::ceil_div(n4, 256 * 4) < 592 ? (int)nefes::ceil_div(n4, 256 * 4) : 592;
  nefes::feat_loss_fwd_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(feat_a), reinterpret_cast<const float4*>(feat_b),
                                                      reinterpret_cast<const float4*>(target), n4, mode, 1.f / (float)n_elems, scratch2,
                                                      reinterpret_cast<unsigned int*>(scratch2 + 1), loss);
  NEFES_CHECK_LAUNCH("feat_loss_fwd");
  return NEFES_OK;
}

int nefes_feat_loss_bwd(const float* feat_a, const float* feat_b, const float* target, const float* d_loss, int64_t n_elems,
                        int mode, float* d_feat_a, float* d_feat_b, void* stream) {
  NEFES_REQUIRE(feat_a && target && d_loss && d_feat_a && (feat_b == nullptr || d_feat_b), NEFES_EINVAL,
                "nefes_feat_loss_bwd: null pointer");
  NEFES_REQUIRE(n_elems >= 4 && n_elems % 4 == 0 && (mode == 0 || mode == 1), NEFES_EINVAL,
                "nefes_feat_loss_bwd: n_elems=%lld must be a positive multiple of 4, mode 0 (L1) or 1 (MSE)", (long long)n_elems);
  const int64_t n4 = n_elems / 4;
  const int blocks = (int)nefes::ceil_div(n4, 256 * 4) < 592 ? (int)nefes::ceil_div(n4, 256 * 4) : 592;
  nefes::feat_loss_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(feat_a), reinterpret_cast<const float4*>(feat_b), reinterpret_cast<const float4*>(target), d_loss,
      n4, mode, 1.f / (float)n_elems, reinterpret_cast<float4*>(d_feat_a), reinterpret_cast<float4*>(d_feat_b));
  NEFES_CHECK_LAUNCH("feat_loss_bwd");
  return NEFES_OK;
}

}  // extern "C"
