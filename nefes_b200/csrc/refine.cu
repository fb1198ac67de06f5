// The refinement iteration's glue around the render, as five small kernels so that one
// iteration is ~30 engine launches and nothing else:
//   pose_rays_fwd   (r, t, init_c2w) -> c2w = [Exp(r) R0 | t + t0] -> packed ray_batch rows     poses.py:25-50 (lietorch=False),
//                                                                       lie_group_helper.py:60-81, ray_utils.py:5-16, rendering.py:197-243
//   cosine_loss_*   1 - mean_c cos(feat[:, c], target[c, :]) and its gradient                   DFM_pose_refine.py:236-255 (per_pixel=False)
//   pose_rays_bwd   d ray_batch -> d c2w (view-direction normalisation and camera rays chained)
//   pose_adam_step  d c2w -> d r, d t through the so(3) exponential; torch.optim.Adam update     DFM_pose_refine.py:380-440
#include "common.cuh"

namespace nefes {

struct Pose34 { float m[12]; };
// What lies between the six learned parameters and the pose the renderer sees:
//   se3 = 0  c2w = [Exp(r) R0 | t + t0]                 poses.py:25-50 with lietorch=False (lie_group_helper.py:60-81)
//   se3 = 1  c2w = [Exp(r) R0 | V(r) t + t0]            poses.py:31-32, 44: SE3.exp([t, r]).matrix(), the translation goes through
//                                                       V(r) = I + (1 - cos n)/n^2 K + (n - sin n)/n^3 K^2   (lietorch=True)
//   then the translation column becomes ((x * sc) + move) * sc2                 dm/direct_pose_model.py:210-232 (fix_coord_supp)
struct PoseChain { int se3; float sc, mv[3], sc2; };

// sin(n)/n, (1 - cos n)/n^2, (n - sin n)/n^3 and their derivatives in n, fp64, series below 1e-4 (the closed forms cancel)
struct ExpCoef { double a, b, c, da, db, dc; };
__device__ __forceinline__ ExpCoef exp_coef(double n) {
  ExpCoef e;
  if (n < 1e-4) {
    const double n2 = n * n;
    e.a = 1.0 - n2 / 6.0;          e.da = -n / 3.0;
    e.b = 0.5 - n2 / 24.0;         e.db = -n / 12.0;
    e.c = 1.0 / 6.0 - n2 / 120.0;  e.dc = -n / 60.0;
  } else {
    const double sn = sin(n), cs = cos(n), n2 = n * n;
    e.a = sn / n;                  e.da = (n * cs - sn) / n2;
    e.b = (1.0 - cs) / n2;         e.db = (n * sn - 2.0 * (1.0 - cs)) / (n2 * n);
    e.c = (n - sn) / (n2 * n);     e.dc = ((1.0 - cs) * n - 3.0 * (n - sn)) / (n2 * n2);
  }
  return e;
}

// c2w in fp32, op for op as LearnPose.forward / lie_group_helper.Exp compute it (se3 = 0); the SE(3) exponential (se3 = 1)
// is evaluated in fp64 and rounded -- lietorch is not vendored by the reference, so that branch follows the closed form
__device__ __forceinline__ Pose34 pose_c2w(const float* __restrict__ pose6, const float* __restrict__ init, const PoseChain ch) {
  const float r0 = pose6[0], r1 = pose6[1], r2 = pose6[2];
  const float K[9] = {0.f, -r2, r1, r2, 0.f, -r0, -r1, r0, 0.f};
  float R[9], tr[3];
  if (ch.se3) {
    const double n = sqrt((double)r0 * r0 + (double)r1 * r1 + (double)r2 * r2);
    const ExpCoef e = exp_coef(n);
    const double t[3] = {pose6[3], pose6[4], pose6[5]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double vt = t[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double kk = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) kk += (double)K[i * 3 + k] * (double)K[k * 3 + j];
        R[i * 3 + j] = (float)((i == j ? 1.0 : 0.0) + e.a * K[i * 3 + j] + e.b * kk);
        vt += (e.b * K[i * 3 + j] + e.c * kk) * t[j];
      }
      tr[i] = (float)vt;
    }
  } else {
    const float n = sqrtf(r0 * r0 + r1 * r1 + r2 * r2) + 1e-15f;
    const float a = sinf(n) / n, b = (1.f - cosf(n)) / (n * n);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float kk = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) kk += K[i * 3 + k] * K[k * 3 + j];
        R[i * 3 + j] = (i == j ? 1.f : 0.f) + a * K[i * 3 + j] + b * kk;
      }
      tr[i] = pose6[3 + i];
    }
  }
  Pose34 P;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * init[k * 4 + j];
      P.m[i * 4 + j] = s;
    }
    P.m[i * 4 + 3] = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(tr[i], init[i * 4 + 3]), ch.sc), ch.mv[i]), ch.sc2);
  }
  return P;
}

__global__ void pose_rays_fwd_kernel(const float* __restrict__ pose6, const float* __restrict__ init, int H, int W, float focal,
                                     float near, float far, float* __restrict__ c2w_out, float* __restrict__ rays, int ld,
                                     const PoseChain ch) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const Pose34 P = pose_c2w(pose6, init, ch);
  if (p == 0 && c2w_out != nullptr)
#pragma unroll
    for (int q = 0; q < 12; ++q) c2w_out[q] = P.m[q];
  const int j = p / W, i = p % W;
  // ray_utils.py:5-16, same operation order as get_rays_fwd_kernel
  const float cx = __fdiv_rn(__fsub_rn((float)i, (float)W * .5f), focal);
  const float cy = -__fdiv_rn(__fsub_rn((float)j, (float)H * .5f), focal);
  float d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    d[c] = __fadd_rn(__fadd_rn(__fmul_rn(cx, P.m[c * 4 + 0]), __fmul_rn(cy, P.m[c * 4 + 1])), __fmul_rn(-1.f, P.m[c * 4 + 2]));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  float* r = rays + (int64_t)p * ld;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r[c] = P.m[c * 4 + 3];
    r[3 + c] = d[c];
    r[8 + c] = __fdiv_rn(d[c], nrm);                 // rendering.py:222: viewdirs / norm
  }
  r[6] = near;
  r[7] = far;
  for (int c = 11; c < ld; ++c) r[c] = 0.f;          // img_idx histogram: accepted and ignored by the field (nerfh_nff.py:168)
}

// d_c2w[c][k<3] += sum_p d_d[p][c] cam[p][k], d_c2w[c][3] += sum_p d_o[p][c], where d_d includes the cotangent of the
// normalised view direction: v = d/|d|  ->  d_d += (g_v - v (v . g_v)) / |d|
__global__ void pose_rays_bwd_kernel(const float* __restrict__ d_rays, const float* __restrict__ rays, int ld, int H, int W,
                                     float focal, float* __restrict__ d_c2w) {
  float acc[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) acc[q] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    const float* g = d_rays + (int64_t)p * ld;
    const float* r = rays + (int64_t)p * ld;
    const int j = p / W, i = p % W;
    const float cam[3] = {((float)i - (float)W * .5f) / focal, -((float)j - (float)H * .5f) / focal, -1.f};
    const float nrm = sqrtf(r[3] * r[3] + r[4] * r[4] + r[5] * r[5]);
    const float vg = r[8] * g[8] + r[9] * g[9] + r[10] * g[10];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gd = g[3 + c] + (g[8 + c] - r[8 + c] * vg) / nrm;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[c * 4 + k] += gd * cam[k];
      acc[c * 4 + 3] += g[c];
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
    atomicAdd(&d_c2w[threadIdx.x], s);
  }
}

// per-channel sums over the pixels: stats[0][c] = sum_n a b, stats[1][c] = sum_n a^2, stats[2][c] = sum_n b^2 with
// a = feat[n][c] (row-major [N,C]) and b = target[c][n] ([C,N]).  Block = C threads (thread = channel) x 32 pixels; the
// target tile goes through shared memory so that both tensors are read along their contiguous axis.
constexpr int kLossPix = 32;
__global__ void cosine_stats_kernel(const float* __restrict__ feat, const float* __restrict__ target, const float* __restrict__ mask,
                                    int N, int C, float* __restrict__ stats) {
  extern __shared__ float tile[];                    // [C][kLossPix + 1]
  const int n0 = blockIdx.x * kLossPix, c = threadIdx.x;
  for (int e = threadIdx.x; e < C * kLossPix; e += blockDim.x) {
    const int cc = e / kLossPix, nn = e % kLossPix;
    tile[cc * (kLossPix + 1) + nn] = n0 + nn < N ? target[(int64_t)cc * N + n0 + nn] : 0.f;
  }
  __syncthreads();
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (int nn = 0; nn < kLossPix && n0 + nn < N; ++nn) {
    if (mask != nullptr && !(mask[n0 + nn] > 0.f)) continue;        // masked_feature_loss: only the valid pixels enter the sums
    const float a = feat[(int64_t)(n0 + nn) * C + c], b = tile[c * (kLossPix + 1) + nn];
    ab += a * b; aa += a * a; bb += b * b;
  }
  atomicAdd(&stats[c], ab);
  atomicAdd(&stats[C + c], aa);
  atomicAdd(&stats[2 * C + c], bb);
}

// loss = 1 - mean_c cos_c, cos_c = ab / (max(|a|, eps) max(|b|, eps))  (F.cosine_similarity, eps = 1e-6);
// d_feat[n][c] = -(1/C) (b / (|a| |b|) - cos_c a / |a|^2).  Block 0 also writes the loss, to loss[0] and, when a
// history is kept, to loss_hist[(int)*step] (the device-side iteration counter of pose_adam_step).
__global__ void cosine_grad_kernel(const float* __restrict__ feat, const float* __restrict__ target, const float* __restrict__ mask,
                                   const float* __restrict__ stats, int N, int C, float* __restrict__ loss, float* __restrict__ loss_hist, const float* __restrict__ step,
                                   int hist_cap, float* __restrict__ d_feat) {
  extern __shared__ float tile[];                    // [C][kLossPix + 1] + [C] cos
  float* s_cos = tile + C * (kLossPix + 1);
  const int n0 = blockIdx.x * kLossPix, c = threadIdx.x;
  const float eps = 1e-6f;
  const float na = fmaxf(sqrtf(stats[C + c]), eps), nb = fmaxf(sqrtf(stats[2 * C + c]), eps);
  const float cosc = stats[c] / (na * nb);
  s_cos[c] = cosc;
  for (int e = threadIdx.x; e < C * kLossPix; e += blockDim.x) {
    const int cc = e / kLossPix, nn = e % kLossPix;
    tile[cc * (kLossPix + 1) + nn] = n0 + nn < N ? target[(int64_t)cc * N + n0 + nn] : 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < C; ++k) s += s_cos[k];
    const float l = 1.f - s / (float)C;
    if (loss != nullptr) loss[0] = l;
    if (loss_hist != nullptr && step != nullptr) {
      const int it = (int)step[0];
      if (it >= 0 && it < hist_cap) loss_hist[it] = l;
    }
  }
  if (d_feat == nullptr) return;
  const float k1 = -1.f / ((float)C * na * nb), k2 = cosc / ((float)C * na * na);
  for (int nn = 0; nn < kLossPix && n0 + nn < N; ++nn) {
    const int64_t o = (int64_t)(n0 + nn) * C + c;
    const bool on = mask == nullptr || mask[n0 + nn] > 0.f;
    d_feat[o] = on ? k1 * tile[c * (kLossPix + 1) + nn] + k2 * feat[o] : 0.f;
  }
}

// ---- bicubic up-sampling + crop of a pixel-major feature map (dm/DFM_APR_refine.py:114-124: torch.nn.Upsample(size=(H, W),
// mode='bicubic'), align_corners=False, then [:, :, crop:-crop, crop:-crop]).  torch's kernel: cubic convolution with
// A = -0.75, source coordinate (X + 0.5) * w / W - 0.5, the four taps clamped to the image.  Only the pixels inside the crop
// window are produced: out [(H - 2 crop) * (W - 2 crop), C].
__device__ __forceinline__ void cubic_taps(int X, int in_size, int out_size, int idx[4], float wt[4]) {
  const float A = -0.75f;
  const float src = ((float)X + 0.5f) * ((float)in_size / (float)out_size) - 0.5f;
  const float fl = floorf(src);
  const float t = src - fl;
  const int i0 = (int)fl;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  wt[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  wt[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  wt[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  wt[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
#pragma unroll
  for (int k = 0; k < 4; ++k) idx[k] = min(max(i0 - 1 + k, 0), in_size - 1);
}
__global__ void upsample_crop_fwd_kernel(const float* __restrict__ x, int h, int w, int C, int H, int W, int crop, float* __restrict__ out) {
  const int Wc = W - 2 * crop, Hc = H - 2 * crop;
  const int64_t n = (int64_t)Hc * Wc * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t p = e / C;
    const int X = (int)(p % Wc) + crop, Y = (int)(p / Wc) + crop;
    int ix[4], iy[4]; float wx[4], wy[4];
    cubic_taps(X, w, W, ix, wx);
    cubic_taps(Y, h, H, iy, wy);
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float r = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) r += wx[b] * x[((int64_t)iy[a] * w + ix[b]) * C + c];
      s += wy[a] * r;
    }
    out[e] = s;
  }
}
// transpose of the above, separable and deterministic: rows first (tmp [h, Wc, C]), then columns (d_x [h, w, C])
__global__ void upsample_crop_bwd_rows_kernel(const float* __restrict__ g, int h, int C, int H, int W, int crop, float* __restrict__ tmp) {
  const int Wc = W - 2 * crop;
  const int64_t n = (int64_t)h * Wc * C;
  const int reach = 2 * ((H + h - 1) / h) + 2;          // output rows whose taps can touch input row y
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t q = e / C;
    const int Xc = (int)(q % Wc), y = (int)(q / Wc);
    const int Yc = (int)(((float)y + 0.5f) * ((float)H / (float)h));
    float s = 0.f;
    for (int Y = max(crop, Yc - reach); Y <= min(H - crop - 1, Yc + reach); ++Y) {
      int iy[4]; float wy[4];
      cubic_taps(Y, h, H, iy, wy);
      float wsum = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) wsum += iy[a] == y ? wy[a] : 0.f;
      if (wsum != 0.f) s += wsum * g[((int64_t)(Y - crop) * Wc + Xc) * C + c];
    }
    tmp[e] = s;
  }
}
__global__ void upsample_crop_bwd_cols_kernel(const float* __restrict__ tmp, int h, int w, int C, int W, int crop, float* __restrict__ dx) {
  const int Wc = W - 2 * crop;
  const int64_t n = (int64_t)h * w * C;
  const int reach = 2 * ((W + w - 1) / w) + 2;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t q = e / C;
    const int x = (int)(q % w), y = (int)(q / w);
    const int Xm = (int)(((float)x + 0.5f) * ((float)W / (float)w));
    float s = 0.f;
    for (int X = max(crop, Xm - reach); X <= min(W - crop - 1, Xm + reach); ++X) {
      int ix[4]; float wx[4];
      cubic_taps(X, w, W, ix, wx);
      float wsum = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) wsum += ix[b] == x ? wx[b] : 0.f;
      if (wsum != 0.f) s += wsum * tmp[((int64_t)y * Wc + (X - crop)) * C + c];
    }
    dx[e] = s;
  }
}

// One thread: d_c2w [3,4] -> (d_r, d_t) through c2w = [Exp(r) R0 | t + t0], then torch.optim.Adam on the two parameter
// groups.  The chain is evaluated in fp64 (a handful of flops; the fp32 autograd chain it replaces is noisier, not
// different).  state: m[6], v[6], step.  Afterwards d_c2w and `zero` (the loss statistics) are cleared for the next
// iteration, so an iteration needs no memset.
__global__ void pose_adam_kernel(float* __restrict__ pose6, const float* __restrict__ init, float* __restrict__ d_c2w,
                                 float* __restrict__ zero, int n_zero, float* __restrict__ state, float lr_r, float lr_t,
                                 float beta1, float beta2, float eps, const PoseChain ch) {
  if (threadIdx.x == 0) {
    const double r[3] = {pose6[0], pose6[1], pose6[2]};
    const double nn = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    ExpCoef e;
    if (ch.se3) {
      e = exp_coef(nn);
    } else {                                          // the reference's own formula: n = |r| + 1e-15 (lie_group_helper.py:66)
      const double n = nn + 1e-15;
      e.a = sin(n) / n; e.b = (1.0 - cos(n)) / (n * n);
      e.da = (n * cos(n) - sin(n)) / (n * n); e.db = (n * sin(n) - 2.0 * (1.0 - cos(n))) / (n * n * n);
      e.c = 0.0; e.dc = 0.0;
    }
    const double K[9] = {0.0, -r[2], r[1], r[2], 0.0, -r[0], -r[1], r[0], 0.0};
    double KK[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j];
        KK[i * 3 + j] = s;
      }
    // dL/dExp = dL/dR_total @ R0^T
    double GE[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += (double)d_c2w[i * 4 + k] * (double)init[j * 4 + k];
        GE[i * 3 + j] = s;
      }
    // cotangent of the translation BEFORE fix_coord_supp's scale / shift / scale
    const double gt[3] = {(double)d_c2w[3] * ch.sc * ch.sc2, (double)d_c2w[7] * ch.sc * ch.sc2, (double)d_c2w[11] * ch.sc * ch.sc2};
    const double t[3] = {pose6[3], pose6[4], pose6[5]};
    double g[6];
    for (int k = 0; k < 3; ++k) {
      double E[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // skew(e_k)
      if (k == 0) { E[5] = -1.0; E[7] = 1.0; }
      if (k == 1) { E[2] = 1.0; E[6] = -1.0; }
      if (k == 2) { E[1] = -1.0; E[3] = 1.0; }
      const double dn = nn > 0.0 ? r[k] / nn : 0.0;  // torch: the norm's subgradient at 0 is 0
      double s = 0.0, gv = gt[k];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double ek = 0.0;                            // (E K + K E)_ij
          for (int q = 0; q < 3; ++q) ek += E[i * 3 + q] * K[q * 3 + j] + K[i * 3 + q] * E[q * 3 + j];
          const double dR = e.da * dn * K[i * 3 + j] + e.a * E[i * 3 + j] + e.db * dn * KK[i * 3 + j] + e.b * ek;
          s += GE[i * 3 + j] * dR;
          if (ch.se3) {
            // translation = V(r) t: d/dr_k through V = I + b K + c K^2, and d/dt_k = column k of V
            const double dV = e.db * dn * K[i * 3 + j] + e.b * E[i * 3 + j] + e.dc * dn * KK[i * 3 + j] + e.c * ek;
            s += gt[i] * dV * t[j];
            if (j == k) gv += gt[i] * (e.b * K[i * 3 + j] + e.c * KK[i * 3 + j]);
          }
        }
      g[k] = s;
      g[3 + k] = gv;
    }
    const float step = state[12] + 1.f;
    state[12] = step;
    const float bc1 = 1.f - powf(beta1, step), bc2s = sqrtf(1.f - powf(beta2, step));
    for (int k = 0; k < 6; ++k) {
      const float gk = (float)g[k], lr = k < 3 ? lr_r : lr_t;
      const float m = state[k] + (gk - state[k]) * (1.f - beta1);           // exp_avg.lerp_(grad, 1 - beta1)
      const float v = state[6 + k] * beta2 + (1.f - beta2) * gk * gk;
      state[k] = m;
      state[6 + k] = v;
      pose6[k] -= (lr / bc1) * (m / (sqrtf(v) / bc2s + eps));
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 12; e += blockDim.x) d_c2w[e] = 0.f;
  for (int e = threadIdx.x; e < n_zero; e += blockDim.x) zero[e] = 0.f;
}

}  // namespace nefes

extern "C" {

// chain6 (HOST pointer, may be NULL = {0, 1, 0, 0, 0, 1}): {se3, pose_scale, move_x, move_y, move_z, pose_scale2}
static nefes::PoseChain pose_chain(const float* chain6) {
  nefes::PoseChain ch = {0, 1.f, {0.f, 0.f, 0.f}, 1.f};
  if (chain6 != nullptr) {
    ch.se3 = chain6[0] != 0.f ? 1 : 0; ch.sc = chain6[1]; ch.mv[0] = chain6[2]; ch.mv[1] = chain6[3]; ch.mv[2] = chain6[4]; ch.sc2 = chain6[5];
  }
  return ch;
}

int nefes_pose_rays_fwd(const float* pose6, const float* init_c2w, int H, int W, float focal, float near, float far,
                        float* c2w_out, float* ray_batch, int ld, const float* chain6, void* stream) {
  NEFES_REQUIRE(pose6 && init_c2w && ray_batch, NEFES_EINVAL, "nefes_pose_rays_fwd: null pointer");
  NEFES_REQUIRE(H > 0 && W > 0 && focal > 0.f && ld >= 11, NEFES_EINVAL, "nefes_pose_rays_fwd: bad shape H=%d W=%d ld=%d", H, W, ld);
  nefes::pose_rays_fwd_kernel<<<(unsigned)nefes::ceil_div((int64_t)H * W, 128), 128, 0, (cudaStream_t)stream>>>(
      pose6, init_c2w, H, W, focal, near, far, c2w_out, ray_batch, ld, pose_chain(chain6));
  NEFES_CHECK_LAUNCH("pose_rays_fwd");
  return NEFES_OK;
}

int nefes_pose_rays_bwd(const float* d_ray_batch, const float* ray_batch, int ld, int H, int W, float focal, float* d_c2w,
                        void* stream) {
  NEFES_REQUIRE(d_ray_batch && ray_batch && d_c2w, NEFES_EINVAL, "nefes_pose_rays_bwd: null pointer");
  NEFES_REQUIRE(H > 0 && W > 0 && focal > 0.f && ld >= 11, NEFES_EINVAL, "nefes_pose_rays_bwd: bad shape");
  const int blocks = (int)nefes::ceil_div((int64_t)H * W, 256 * 2);
  nefes::pose_rays_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_ray_batch, ray_batch, ld, H, W, focal, d_c2w);
  NEFES_CHECK_LAUNCH("pose_rays_bwd");
  return NEFES_OK;
}

int nefes_cosine_loss_fwd(const float* feat, const float* target, const float* mask, int N, int C, float* stats, void* stream) {
  NEFES_REQUIRE(feat && target && stats, NEFES_EINVAL, "nefes_cosine_loss_fwd: null pointer");
  NEFES_REQUIRE(N > 0 && C > 0 && C <= 1024 && C % 32 == 0, NEFES_EINVAL, "nefes_cosine_loss_fwd: bad shape N=%d C=%d", N, C);
  const size_t smem = sizeof(float) * C * (nefes::kLossPix + 1);
  nefes::cosine_stats_kernel<<<(unsigned)nefes::ceil_div(N, nefes::kLossPix), C, smem, (cudaStream_t)stream>>>(feat, target, mask, N, C, stats);
  NEFES_CHECK_LAUNCH("cosine_stats");
  return NEFES_OK;
}

int nefes_cosine_loss_bwd(const float* feat, const float* target, const float* mask, const float* stats, int N, int C, float* loss,
                          float* loss_hist, const float* step, int hist_cap, float* d_feat, void* stream) {
  NEFES_REQUIRE(feat && target && stats, NEFES_EINVAL, "nefes_cosine_loss_bwd: null pointer");
  NEFES_REQUIRE(N > 0 && C > 0 && C <= 1024 && C % 32 == 0, NEFES_EINVAL, "nefes_cosine_loss_bwd: bad shape N=%d C=%d", N, C);
  const size_t smem = sizeof(float) * (C * (nefes::kLossPix + 1) + C);
  nefes::cosine_grad_kernel<<<(unsigned)nefes::ceil_div(N, nefes::kLossPix), C, smem, (cudaStream_t)stream>>>(
      feat, target, mask, stats, N, C, loss, loss_hist, step, hist_cap, d_feat);
  NEFES_CHECK_LAUNCH("cosine_grad");
  return NEFES_OK;
}

int nefes_upsample_crop_fwd(const float* x, int h, int w, int C, int H, int W, int crop, float* out, void* stream) {
  NEFES_REQUIRE(x && out, NEFES_EINVAL, "nefes_upsample_crop_fwd: null pointer");
  NEFES_REQUIRE(h > 0 && w > 0 && C > 0 && H >= h && W >= w && crop >= 0 && H > 2 * crop && W > 2 * crop, NEFES_EINVAL,
                "nefes_upsample_crop_fwd: bad shape %dx%d -> %dx%d crop %d", h, w, H, W, crop);
  const int64_t n = (int64_t)(H - 2 * crop) * (W - 2 * crop) * C;
  const int64_t b = nefes::ceil_div(n, 256);
  nefes::upsample_crop_fwd_kernel<<<(unsigned)(b < 148 * 16 ? b : 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, h, w, C, H, W, crop, out);
  NEFES_CHECK_LAUNCH("upsample_crop_fwd");
  return NEFES_OK;
}

int nefes_upsample_crop_bwd(const float* d_out, int h, int w, int C, int H, int W, int crop, float* tmp, float* d_x, void* stream) {
  NEFES_REQUIRE(d_out && tmp && d_x, NEFES_EINVAL, "nefes_upsample_crop_bwd: null pointer");
  NEFES_REQUIRE(h > 0 && w > 0 && C > 0 && H >= h && W >= w && crop >= 0 && H > 2 * crop && W > 2 * crop, NEFES_EINVAL,
                "nefes_upsample_crop_bwd: bad shape %dx%d -> %dx%d crop %d", h, w, H, W, crop);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t b = nefes::ceil_div((int64_t)h * (W - 2 * crop) * C, 256);
  nefes::upsample_crop_bwd_rows_kernel<<<(unsigned)(b < 148 * 16 ? b : 148 * 16), 256, 0, st>>>(d_out, h, C, H, W, crop, tmp);
  NEFES_CHECK_LAUNCH("upsample_crop_bwd_rows");
  b = nefes::ceil_div((int64_t)h * w * C, 256);
  nefes::upsample_crop_bwd_cols_kernel<<<(unsigned)(b < 148 * 16 ? b : 148 * 16), 256, 0, st>>>(tmp, h, w, C, W, crop, d_x);
  NEFES_CHECK_LAUNCH("upsample_crop_bwd_cols");
  return NEFES_OK;
}

int nefes_pose_adam_step(float* pose6, const float* init_c2w, float* d_c2w, float* zero, int n_zero, float* state13,
                         float lr_r, float lr_t, float beta1, float beta2, float eps, const float* chain6, void* stream) {
  NEFES_REQUIRE(pose6 && init_c2w && d_c2w && state13, NEFES_EINVAL, "nefes_pose_adam_step: null pointer");
  NEFES_REQUIRE(n_zero >= 0 && (zero != nullptr || n_zero == 0), NEFES_EINVAL, "nefes_pose_adam_step: bad zero range");
  nefes::pose_adam_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(pose6, init_c2w, d_c2w, zero, n_zero, state13, lr_r, lr_t, beta1,
                                                               beta2, eps, pose_chain(chain6));
  NEFES_CHECK_LAUNCH("pose_adam");
  return NEFES_OK;
}

}  // extern "C"
