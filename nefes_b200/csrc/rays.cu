// K1 get_rays (+ backward to the camera pose) and K2 stratified coarse depths.
// Arithmetic follows the reference op by op with explicit round-to-nearest intrinsics so the
// compiler cannot contract mul+add into FMA: results are bit-equal to the CPU path.
#include "common.cuh"

namespace nefes {

// script/models/ray_utils.py:5-16.  One thread per pixel; rays_d[c] = sum_k cam[k]*c2w[c][k].
__global__ void get_rays_fwd_kernel(const float* __restrict__ c2w, int B, int H, int W, float focal,
                                    float* __restrict__ rays_o, float* __restrict__ rays_d) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t npix = (int64_t)H * W;
  if (idx >= B * npix) return;
  const int b = (int)(idx / npix);
  const int p = (int)(idx % npix);
  const int j = p / W, i = p % W;
  const float* M = c2w + b * 12;
  const float cx = __fdiv_rn(__fsub_rn((float)i, (float)W * .5f), focal);
  const float cy = -__fdiv_rn(__fsub_rn((float)j, (float)H * .5f), focal);
  const float cz = -1.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float p0 = __fmul_rn(cx, M[c * 4 + 0]);
    const float p1 = __fmul_rn(cy, M[c * 4 + 1]);
    const float p2 = __fmul_rn(cz, M[c * 4 + 2]);
    rays_d[idx * 3 + c] = __fadd_rn(__fadd_rn(p0, p1), p2);
    rays_o[idx * 3 + c] = M[c * 4 + 3];
  }
}

// d_c2w[b][c][k<3] = sum_p d_rays_d[p][c] * cam[p][k];  d_c2w[b][c][3] = sum_p d_rays_o[p][c].
// grid (blocks_per_image, B); block-level tree reduction then 12 atomics per block.
__global__ void get_rays_bwd_kernel(const float* __restrict__ d_o, const float* __restrict__ d_d,
                                    int H, int W, float focal, float* __restrict__ d_c2w) {
  const int b = blockIdx.y;
  const int npix = H * W;
  float acc[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) acc[q] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const int j = p / W, i = p % W;
    const float cam[3] = {((float)i - (float)W * .5f) / focal, -((float)j - (float)H * .5f) / focal, -1.f};
    const int64_t base = ((int64_t)b * npix + p) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gd = d_d ? d_d[base + c] : 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[c * 4 + k] += gd * cam[k];
      acc[c * 4 + 3] += d_o ? d_o[base + c] : 0.f;
    }
  }
  __shared__ float red[12][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const float v = warp_sum(acc[q]);
    if (lane == 0) red[q][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    atomicAdd(&d_c2w[b * 12 + threadIdx.x], s);
  }
}

// script/models/rendering.py:96-112 (lindisp=False).
__global__ void sample_coarse_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                     int ld_nf, const float* __restrict__ t_vals,
                                     const float* __restrict__ t_rand, int N, int S,
                                     float* __restrict__ z_vals) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * S) return;
  const int r = (int)(idx / S), s = (int)(idx % S);
  const float nr = near[(int64_t)r * ld_nf], fr = far[(int64_t)r * ld_nf];
  auto zat = [&](int k) {
    const float t = t_vals[k];
    return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
  };
  float z = zat(s);
  if (t_rand != nullptr) {
    const float up = (s + 1 < S) ? __fmul_rn(.5f, __fadd_rn(zat(s + 1), z)) : z;
    const float lo = (s > 0) ? __fmul_rn(.5f, __fadd_rn(z, zat(s - 1))) : z;
    z = __fadd_rn(lo, __fmul_rn(__fsub_rn(up, lo), t_rand[idx]));
  }
  z_vals[idx] = z;
}

}  // namespace nefes

extern "C" {

int nefes_get_rays_fwd(const float* c2w, int B, int H, int W, float focal, float* rays_o,
                       float* rays_d, void* stream) {
  NEFES_REQUIRE(c2w && rays_o && rays_d, NEFES_EINVAL, "nefes_get_rays_fwd: null pointer");
  NEFES_REQUIRE(B > 0 && H > 0 && W > 0 && focal > 0.f, NEFES_EINVAL,
                "nefes_get_rays_fwd: bad shape B=%d H=%d W=%d focal=%g", B, H, W, focal);
  const int64_t n = (int64_t)B * H * W;
  nefes::get_rays_fwd_kernel<<<(unsigned)nefes::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      c2w, B, H, W, focal, rays_o, rays_d);
  NEFES_CHECK_LAUNCH("get_rays_fwd");
  return NEFES_OK;
}

int nefes_get_rays_bwd(const float* d_rays_o, const float* d_rays_d, int B, int H, int W,
                       float focal, float* d_c2w, void* stream) {
  NEFES_REQUIRE(d_c2w, NEFES_EINVAL, "nefes_get_rays_bwd: null d_c2w");
  NEFES_REQUIRE(B > 0 && H > 0 && W > 0 && focal > 0.f, NEFES_EINVAL, "nefes_get_rays_bwd: bad shape");
  NEFES_CUDA(cudaMemsetAsync(d_c2w, 0, sizeof(float) * 12 * B, (cudaStream_t)stream));
  const int blocks = (int)nefes::ceil_div((int64_t)H * W, 256 * 4);
  nefes::get_rays_bwd_kernel<<<dim3(blocks, B), 256, 0, (cudaStream_t)stream>>>(
      d_rays_o, d_rays_d, H, W, focal, d_c2w);
  NEFES_CHECK_LAUNCH("get_rays_bwd");
  return NEFES_OK;
}

int nefes_sample_coarse(const float* near, const float* far, int ld_nf, const float* t_vals,
                        const float* t_rand, int N, int S, float* z_vals, void* stream) {
  NEFES_REQUIRE(near && far && t_vals && z_vals, NEFES_EINVAL, "nefes_sample_coarse: null pointer");
  NEFES_REQUIRE(N >= 0 && S > 0 && ld_nf >= 0, NEFES_EINVAL, "nefes_sample_coarse: bad shape");
  if (N == 0) return NEFES_OK;
  const int64_t n = (int64_t)N * S;
  nefes::sample_coarse_kernel<<<(unsigned)nefes::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      near, far, ld_nf, t_vals, t_rand, N, S, z_vals);
  NEFES_CHECK_LAUNCH("sample_coarse");
  return NEFES_OK;
}

}  // extern "C"
