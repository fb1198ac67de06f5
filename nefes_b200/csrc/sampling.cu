// K3 hierarchical sampling: pdf -> cdf (fp64 running sum, as torch's CPU cumsum accumulates),
// inverse-CDF lookup (searchsorted right=True), and the merge sort of coarse + fine depths.
// One CTA per ray; every scan is a warp-shuffle scan.  script/models/rendering.py:23-66,132-141.
#include "common.cuh"

namespace nefes {

constexpr int kMaxBins = 256;

// inclusive scan of doubles over a CTA (blockDim.x <= 256) -- warp shuffles + one smem hop
__device__ __forceinline__ double block_inclusive_scan(double v, double* warp_tot /*[8]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  double base = 0.0;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  __syncthreads();
  return v + base;
}

__device__ __forceinline__ float block_sum(float v, float* red /*[8]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) s += red[w];
  __syncthreads();
  return s;
}

// rendering.py:49-64 for one u: returns the sample, writes inds.
__device__ __forceinline__ float invert_one(const float* s_cdf, const float* s_bins, int nb, float u,
                                            int* ind_out) {
  int lo = 0, hi = nb;                       // first index with cdf[idx] > u  (right=True)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  const int below = max(lo - 1, 0), above = min(lo, nb - 1);
  const float cb = s_cdf[below], ca = s_cdf[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.f;
  const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  const float bb = s_bins[below], ba = s_bins[above];
  *ind_out = lo;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

// weights [nb-1] (already in smem as raw weights) -> s_cdf [nb].  rendering.py:26-29.
__device__ __forceinline__ void build_cdf(const float* s_w, int nb, float* s_cdf, float* red, double* dred) {
  const int t = threadIdx.x;
  const float w = (t < nb - 1) ? __fadd_rn(s_w[t], 1e-5f) : 0.f;
  const float tot = block_sum(w, red);
  const float pdf = (t < nb - 1) ? __fdiv_rn(w, tot) : 0.f;
  const double run = block_inclusive_scan((double)pdf, dred);
  if (t < nb - 1) s_cdf[t + 1] = (float)run;
  if (t == 0) s_cdf[0] = 0.f;
  __syncthreads();
}

__global__ void sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                  const float* __restrict__ cdf_in, const float* __restrict__ u,
                                  int u_per_ray, int nb, int ns, float* __restrict__ samples,
                                  int32_t* __restrict__ inds, float* __restrict__ cdf_out) {
  __shared__ float s_bins[kMaxBins], s_cdf[kMaxBins], s_w[kMaxBins], red[8];
  __shared__ double dred[8];
  const int r = blockIdx.x, t = threadIdx.x;
  if (t < nb) s_bins[t] = bins[(int64_t)r * nb + t];
  if (cdf_in != nullptr) {
    if (t < nb) s_cdf[t] = cdf_in[(int64_t)r * nb + t];
    __syncthreads();
  } else {
    if (t < nb - 1) s_w[t] = weights[(int64_t)r * (nb - 1) + t];
    __syncthreads();
    build_cdf(s_w, nb, s_cdf, red, dred);
  }
  if (cdf_out != nullptr && t < nb) cdf_out[(int64_t)r * nb + t] = s_cdf[t];
  if (t < ns) {
    const float uu = u_per_ray ? u[(int64_t)r * ns + t] : u[t];
    int ind;
    samples[(int64_t)r * ns + t] = invert_one(s_cdf, s_bins, nb, uu, &ind);
    if (inds != nullptr) inds[(int64_t)r * ns + t] = ind;
  }
}

// rendering.py:132-141: bins = mids(z_coarse) [S-1], weights = w[1:-1] [S-2], ns samples, then the
// sorted union [S+ns].  Sort = rank sort (each thread counts the elements ordered before its own).
__global__ void sample_fine_kernel(const float* __restrict__ z_coarse, const float* __restrict__ w_coarse,
                                   const float* __restrict__ u, int u_per_ray, int S, int ns,
                                   float* __restrict__ z_fine, float* __restrict__ z_samples,
                                   int32_t* __restrict__ inds) {
  __shared__ float s_bins[kMaxBins], s_cdf[kMaxBins], s_w[kMaxBins], s_all[2 * kMaxBins], red[8];
  __shared__ double dred[8];
  const int r = blockIdx.x, t = threadIdx.x;
  const int nb = S - 1;
  if (t < S) s_all[t] = z_coarse[(int64_t)r * S + t];
  if (t < S - 2) s_w[t] = w_coarse[(int64_t)r * S + t + 1];
  __syncthreads();
  if (t < nb) s_bins[t] = __fmul_rn(.5f, __fadd_rn(s_all[t + 1], s_all[t]));
  build_cdf(s_w, nb, s_cdf, red, dred);      // contains the barriers that publish s_bins
  if (t < ns) {
    const float uu = u_per_ray ? u[(int64_t)r * ns + t] : u[t];
    int ind;
    const float z = invert_one(s_cdf, s_bins, nb, uu, &ind);
    s_all[S + t] = z;
    if (z_samples != nullptr) z_samples[(int64_t)r * ns + t] = z;
    if (inds != nullptr) inds[(int64_t)r * ns + t] = ind;
  }
  __syncthreads();
  const int tot = S + ns;
  for (int i = t; i < tot; i += blockDim.x) {
    const float v = s_all[i];
    int rank = 0;
    for (int j = 0; j < tot; ++j) {
      const float o = s_all[j];
      rank += (o < v) || (o == v && j < i);
    }
    z_fine[(int64_t)r * tot + rank] = v;
  }
}

}  // namespace nefes

extern "C" {

static int launch_pdf(const float* bins, const float* weights, const float* cdf, const float* u,
                      int u_per_ray, int N, int nb, int ns, float* samples, int32_t* inds,
                      float* cdf_out, void* stream, const char* who) {
  NEFES_REQUIRE(bins && u && samples && (weights || cdf), NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(N >= 0 && nb >= 2 && nb <= nefes::kMaxBins && ns >= 1 && ns <= nefes::kMaxBins,
                NEFES_EINVAL, "%s: need 2 <= nb <= 256, 1 <= ns <= 256 (nb=%d ns=%d)", who, nb, ns);
  if (N == 0) return NEFES_OK;
  const int threads = (int)nefes::round_up(nb > ns ? nb : ns, 32);
  nefes::sample_pdf_kernel<<<N, threads, 0, (cudaStream_t)stream>>>(bins, weights, cdf, u, u_per_ray,
                                                                   nb, ns, samples, inds, cdf_out);
  NEFES_CHECK_LAUNCH(who);
  return NEFES_OK;
}

int nefes_sample_pdf(const float* bins, const float* weights, const float* u, int u_per_ray, int N,
                     int nb, int ns, float* samples, int32_t* inds, float* cdf_out, void* stream) {
  return launch_pdf(bins, weights, nullptr, u, u_per_ray, N, nb, ns, samples, inds, cdf_out, stream,
                    "nefes_sample_pdf");
}

int nefes_sample_pdf_from_cdf(const float* bins, const float* cdf, const float* u, int u_per_ray,
                              int N, int nb, int ns, float* samples, int32_t* inds, void* stream) {
  return launch_pdf(bins, nullptr, cdf, u, u_per_ray, N, nb, ns, samples, inds, nullptr, stream,
                    "nefes_sample_pdf_from_cdf");
}

int nefes_sample_fine(const float* z_coarse, const float* weights_coarse, const float* u,
                      int u_per_ray, int N, int S, int ns, float* z_fine, float* z_samples,
                      int32_t* inds, void* stream) {
  NEFES_REQUIRE(z_coarse && weights_coarse && u && z_fine, NEFES_EINVAL, "nefes_sample_fine: null pointer");
  NEFES_REQUIRE(N >= 0 && S >= 3 && S <= nefes::kMaxBins && ns >= 1 && ns <= nefes::kMaxBins,
                NEFES_EINVAL, "nefes_sample_fine: need 3 <= S <= 256, 1 <= ns <= 256 (S=%d ns=%d)", S, ns);
  if (N == 0) return NEFES_OK;
  const int threads = (int)nefes::round_up(S > ns ? S : ns, 32);
  nefes::sample_fine_kernel<<<N, threads, 0, (cudaStream_t)stream>>>(z_coarse, weights_coarse, u,
                                                                    u_per_ray, S, ns, z_fine,
                                                                    z_samples, inds);
  NEFES_CHECK_LAUNCH("nefes_sample_fine");
  return NEFES_OK;
}

}  // extern "C"
