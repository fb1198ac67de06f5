// Post-render appearance + fusion stage, forward and backward, as engine launches:
//   FusionNet          script/models/nerfh_nff.py:356-418 (4 convolutions 131 -> 64 -> 64 -> 64 -> 128, kernels 3,3,3,5, ReLU
//                      between, BatchNorm2d on the output; the rgb channels are normalised with the ImageNet mean / std first)
//                      and :578-603 (run_fusion_net: [B*N,3] + [B*N,128] -> NCHW blob -> net)
//   affine colour      nerfh_nff.py:511-522, :605-626: the exposure network (tiny-cuda-nn FullyFusedMLP 10 -> 32 -> 32 -> 32 -> 12,
//   transform          ReLU, no biases, inputs padded to 16 with ONES) turns an image's 10-bin histogram into a 3x3 colour
//                      matrix and an offset; rgb' = sigmoid(K rgb + b) per ray of that image
//
// Everything is PIXEL-MAJOR ([P, C], P = B*H*W image-major, which is what the render hands over and what the losses take):
// a convolution is im2col ([P, Cin*k*k], column order (c, kh, kw) = the flattened torch weight [Cout, Cin, kh, kw]) followed
// by the fp32 GEMM of the field's parity path (sgemm.cuh): fp32 arithmetic, so the parity bar is the reference's fp32
// one.  At the stage-3 shape (28 patches of 16x16 = 7168 pixels) the four GEMMs are 2.5 GMAC forward; the stage is ~1 % of
// the render's work and launch-bound in the reference (cuDNN, ~20 launches + NCHW shuffles).
// The backward recomputes the im2col matrices (they are 12x the activations) and keeps only X0, A1..A3 and the pre-BN output.
#include "common.cuh"
#include "tf32_gemm.cuh"
#include <algorithm>

namespace nefes {

constexpr int kFusIn = 131, kFusHid = 64;   // kFeat = 128 (common.cuh)
__constant__ float kImgMean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float kImgStd[3] = {0.229f, 0.224f, 0.225f};

// X0[p] = [(rgb - mean) / std | feat]
__global__ void fusion_pack_kernel(const float* __restrict__ rgb, const float* __restrict__ feat, int64_t P, float* __restrict__ X0) {
  const int64_t n = P * kFusIn;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / kFusIn;
    const int c = (int)(e % kFusIn);
    X0[e] = c < 3 ? (rgb[p * 3 + c] - kImgMean[c]) / kImgStd[c] : feat[p * kFeat + (c - 3)];
  }
}

// col[p][c*k*k + kh*k + kw] = x[image(p)][y + kh - pad][x + kw - pad][c], zero outside the image.
// One block per pixel: its k x k x C neighbourhood is read tap by tap (C contiguous floats each) into shared memory and leaves
// as the pixel's K contiguous columns -- both sides coalesced.  (One thread per element gathered 4-byte words at 0.56 TB/s of
// writes: 60 us per call, eight calls per stage-3 step.)
__global__ void im2col_kernel(const float* __restrict__ x, int B, int H, int W, int C, int k, float* __restrict__ col) {
  extern __shared__ float patch[];                            // [k*k][C]
  const int kk = k * k, pad = k / 2, K = C * kk;
  const int64_t P = (int64_t)B * H * W;
  for (int64_t p = blockIdx.x; p < P; p += gridDim.x) {
    const int px = (int)(p % W), py = (int)((p / W) % H);
    const int64_t img = p / ((int64_t)H * W);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
      const int tap = i / C, c = i - tap * C;
      const int yy = py + tap / k - pad, xx = px + tap % k - pad;
      patch[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[((img * H + yy) * W + xx) * C + c] : 0.f;
    }
    __syncthreads();
    float* row = col + p * K;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
      const int c = j / kk, tap = j - c * kk;
      row[j] = patch[tap * C + c];
    }
    __syncthreads();
  }
}

// dx[p][c] = sum_{kh,kw} dcol[q][c*k*k + kh*k + kw] with q = p - (kh - pad, kw - pad) inside the image; gated by relu_src > 0
__global__ void col2im_kernel(const float* __restrict__ dcol, int B, int H, int W, int C, int k, const float* __restrict__ relu_src,
                              float* __restrict__ dx) {
  const int kk = k * k, pad = k / 2, K = C * kk;
  const int64_t n = (int64_t)B * H * W * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / C;
    const int c = (int)(e % C);
    const int px = (int)(p % W), py = (int)((p / W) % H);
    const int64_t img = p / ((int64_t)H * W);
    float s = 0.f;
    for (int kh = 0; kh < k; ++kh) {
      const int yy = py - (kh - pad);
      if (yy < 0 || yy >= H) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int xx = px - (kw - pad);
        if (xx < 0 || xx >= W) continue;
        s += dcol[((img * H + yy) * W + xx) * K + c * kk + kh * k + kw];
      }
    }
    if (relu_src != nullptr && !(relu_src[e] > 0.f)) s = 0.f;
    dx[e] = s;
  }
}

// per-channel sums over the pixels: out[0][c] = sum_p a[p][c], out[1][c] = sum_p a[p][c] * b[p][c] (b may be null: a^2)
__global__ void chan_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t P, int C, int64_t rows_per_block,
                                 float* __restrict__ out) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const int64_t p0 = (int64_t)blockIdx.x * rows_per_block, p1 = min(P, p0 + rows_per_block);
  float s = 0.f, q = 0.f;
  for (int64_t p = p0; p < p1; ++p) {
    const float v = a[p * C + c];
    s += v;
    q += v * (b != nullptr ? b[p * C + c] : v);
  }
  atomicAdd(out + c, s);
  atomicAdd(out + C + c, q);
}

// BatchNorm2d(128) forward.  training: batch statistics (biased variance for the normalisation, unbiased for the running
// estimate, momentum as torch); eval: the running statistics.  stat[0..C) = mean, stat[C..2C) = 1/sqrt(var + eps) are kept
// for the backward.  residual (fusion_residule): out += feat.
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int64_t P, int C, int training, float eps, float momentum,
                                   float* __restrict__ run_mean, float* __restrict__ run_var, float* __restrict__ stat) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (training) {
    const float mean = sums[c] / (float)P;
    const float var = fmaxf(sums[C + c] / (float)P - mean * mean, 0.f);
    stat[c] = mean;
    stat[C + c] = rsqrtf(var + eps);
    if (run_mean != nullptr) {
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * var * ((float)P / (float)max((int64_t)1, P - 1));
    }
  } else {
    stat[c] = run_mean[c];
    stat[C + c] = rsqrtf(run_var[c] + eps);
  }
}
__global__ void bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ stat, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ residual, int64_t P, int C,
                                float* __restrict__ out) {
  const int64_t n = P * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    float v = y[e];
    if (stat != nullptr) v = (v - stat[c]) * stat[C + c] * gamma[c] + beta[c];
    if (residual != nullptr) v += residual[e];
    out[e] = v;
  }
}
// dY from d_out.  training: dY = gamma * invstd * (g - mean_p g - xhat * mean_p (g xhat)); eval: dY = gamma * invstd * g.
// sums[0][c] = sum_p g, sums[1][c] = sum_p g * y (raw): sum_p g xhat = invstd * (sums[1] - mean * sums[0]).
__global__ void bn_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ stat,
                              const float* __restrict__ gamma, const float* __restrict__ sums, int64_t P, int C, int training,
                              float* __restrict__ dy, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  const int64_t n = P * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const float mean = stat[c], inv = stat[C + c];
    const float sg = sums[c], sgx = inv * (sums[C + c] - mean * sums[c]);
    const float xhat = (y[e] - mean) * inv;
    dy[e] = training ? gamma[c] * inv * (g[e] - sg / (float)P - xhat * sgx / (float)P) : gamma[c] * inv * g[e];
    if (e < C) {                                       // one thread per channel also writes the affine gradients
      if (d_gamma != nullptr) d_gamma[c] += sgx;
      if (d_beta != nullptr) d_beta[c] += sg;
    }
  }
}

__global__ void fusion_unpack_grad_kernel(const float* __restrict__ dX0, int64_t P, float* __restrict__ d_rgb, float* __restrict__ d_feat,
                                          const float* __restrict__ residual_grad) {
  const int64_t n = P * kFusIn;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / kFusIn;
    const int c = (int)(e % kFusIn);
    if (c < 3) { if (d_rgb != nullptr) d_rgb[p * 3 + c] = dX0[e] / kImgStd[c]; }
    else if (d_feat != nullptr) d_feat[p * kFeat + (c - 3)] = dX0[e] + (residual_grad != nullptr ? residual_grad[p * kFeat + (c - 3)] : 0.f);
  }
}

__global__ void colsum_acc_kernel(const float* __restrict__ G, int64_t P, int C, int64_t rows_per_block, float* __restrict__ db) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const int64_t p0 = (int64_t)blockIdx.x * rows_per_block, p1 = min(P, p0 + rows_per_block);
  float s = 0.f;
  for (int64_t p = p0; p < p1; ++p) s += G[p * C + c];
  atomicAdd(db + c, s);
}

struct FusionWs { float *X0, *A1, *A2, *A3, *Y, *stat, *sums, *col, *G, *G2, *colL[4]; };
constexpr int kFusColK[4] = {kFusIn * 9, kFusHid * 9, kFusHid * 9, kFusHid * 25};   // im2col widths of the four convolutions
inline int64_t fusion_ws_floats(int64_t P) {
  return P * (kFusIn + 3 * kFusHid + kFeat) + 1024 + P * (int64_t)(kFusHid * 25) + 2 * P * (int64_t)kFusIn +
         P * (int64_t)(kFusColK[0] + kFusColK[1] + kFusColK[2] + kFusColK[3]);
}
inline FusionWs fusion_carve(void* ws, int64_t P) {
  FusionWs w;
  float* f = reinterpret_cast<float*>(ws);
  w.X0 = f; f += P * kFusIn;
  w.A1 = f; f += P * kFusHid;
  w.A2 = f; f += P * kFusHid;
  w.A3 = f; f += P * kFusHid;
  w.Y = f; f += P * kFeat;
  w.stat = f; f += 512;
  w.sums = f; f += 512;
  w.col = f; f += P * (int64_t)(kFusHid * 25);       // the largest im2col matrix (conv4: 64 * 25; conv1: 131 * 9 = 1179 < 1600)
  w.G = f; f += P * (int64_t)kFusIn;                  // gradient of a layer input (<= 131 channels)
  w.G2 = f; f += P * (int64_t)kFusIn;
  // the im2col matrix of every layer is kept from the forward: the weight gradient reads it again (113 MB at 7168 pixels)
  for (int l = 0; l < 4; ++l) { w.colL[l] = f; f += P * (int64_t)kFusColK[l]; }
  return w;
}
inline unsigned grid_for(int64_t n) { const int64_t b = ceil_div(n, 256); return (unsigned)(b < 148 * 16 ? b : 148 * 16); }

inline int conv_fwd(cudaStream_t st, const float* x, int B, int H, int W, int Cin, int k, const float* wgt, const float* bias, int Cout,
                    int act, float* col, float* y) {
  const int64_t P = (int64_t)B * H * W;
  const int K = Cin * k * k;
  im2col_kernel<<<(unsigned)(P < 148 * 64 ? P : 148 * 64), 256, (size_t)K * 4, st>>>(x, B, H, W, Cin, k, col);
  NEFES_CHECK_LAUNCH("im2col");
  return linear_fwd(st, col, K, wgt, K, bias, y, Cout, P, Cout, K, act, 0);
}
// dy [P,Cout] (already gated by this layer's ReLU) -> dW, db accumulated; dx [P,Cin] gated by relu_src (the input's ReLU) or null
inline int conv_bwd(cudaStream_t st, const float* col_x, const float* dy, int B, int H, int W, int Cin, int k, const float* wgt, int Cout,
                    float* col, float* dW, float* db, const float* relu_src, float* dx) {
  const int64_t P = (int64_t)B * H * W;
  const int K = Cin * k * k;
  if (dW != nullptr) {                                  // col_x: im2col of the layer's input, kept by the forward
    if (int e = linear_wgrad(st, dy, Cout, col_x, K, dW, K, P, Cout, K)) return e;
  }
  if (db != nullptr) {
    const int64_t rpb = 64;
    colsum_acc_kernel<<<(unsigned)ceil_div(P, rpb), 128, 0, st>>>(dy, P, Cout, rpb, db);
    NEFES_CHECK_LAUNCH("colsum_acc");
  }
  if (dx != nullptr) {
    if (int e = linear_dgrad(st, dy, Cout, wgt, K, col, K, P, Cout, K, nullptr, 0, 0)) return e;    // d_col over the same buffer
    col2im_kernel<<<grid_for(P * Cin), 256, 0, st>>>(col, B, H, W, Cin, k, relu_src, dx);
    NEFES_CHECK_LAUNCH("col2im");
  }
  return NEFES_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// affine colour transform.  params: the exposure network's flat buffer, [32x16 | 32x32 | 32x32 | 16x32] row-major [out, in].
// ab[b][0..9) = K (row-major 3x3), ab[b][9..12) = offset; hidden activations h1..h3 kept for the backward.
__global__ void exposure_mlp_kernel(const float* __restrict__ params, const float* __restrict__ hist, int B, float* __restrict__ ab,
                                    float* __restrict__ hid) {
  const int b = blockIdx.x, t = threadIdx.x;           // 32 threads
  __shared__ float x[32], h[32];
  x[t] = t < 10 ? truncf(hist[b * 10 + t]) : (t < 16 ? 1.f : 0.f);      // hist.long(): truncation; padded inputs are ones
  __syncwarp();
  const float* w = params;
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += w[t * 16 + i] * x[i];
  h[t] = fmaxf(s, 0.f); hid[(b * 3 + 0) * 32 + t] = h[t];
  __syncwarp();
  w += 32 * 16;
  for (int l = 1; l < 3; ++l) {
    s = 0.f;
    for (int i = 0; i < 32; ++i) s += w[t * 32 + i] * h[i];
    __syncwarp();
    h[t] = fmaxf(s, 0.f); hid[(b * 3 + l) * 32 + t] = h[t];
    __syncwarp();
    w += 32 * 32;
  }
  if (t < 12) {
    s = 0.f;
    for (int i = 0; i < 32; ++i) s += w[t * 32 + i] * h[i];
    ab[b * 12 + t] = s;
  }
}
__global__ void affine_color_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ ab, int64_t n_per_img, int64_t N,
                                        float* __restrict__ out) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    const float* a = ab + (p / n_per_img) * 12;
    const float r = rgb[p * 3], g = rgb[p * 3 + 1], b = rgb[p * 3 + 2];
#pragma unroll
    for (int c = 0; c < 3; ++c) out[p * 3 + c] = sigmoid_f(a[c * 3] * r + a[c * 3 + 1] * g + a[c * 3 + 2] * b + a[9 + c]);
  }
}
// d_rgb = K^T (g * s (1 - s)); d_ab[img] += [outer(g s (1 - s), rgb) | g s (1 - s)]
__global__ void affine_color_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ out, const float* __restrict__ g,
                                        const float* __restrict__ ab, int64_t n_per_img, int64_t N, float* __restrict__ d_rgb,
                                        float* __restrict__ d_ab) {
  const int64_t img = blockIdx.y;
  float acc[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) acc[q] = 0.f;
  const float* a = ab + img * 12;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = img * n_per_img + i;
    if (p >= N) break;
    const float x[3] = {rgb[p * 3], rgb[p * 3 + 1], rgb[p * 3 + 2]};
    float gz[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float s = out[p * 3 + c]; gz[c] = g[p * 3 + c] * s * (1.f - s); }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (d_rgb != nullptr) d_rgb[p * 3 + c] = a[c] * gz[0] + a[3 + c] * gz[1] + a[6 + c] * gz[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[c * 3 + k] += gz[c] * x[k];
      acc[9 + c] += gz[c];
    }
  }
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const float v = warp_sum(acc[q]);
    if ((threadIdx.x & 31) == 0 && d_ab != nullptr) atomicAdd(d_ab + img * 12 + q, v);
  }
}
// back through the four matrices; one block (32 threads) per image, weight gradients accumulated with atomics
__global__ void exposure_mlp_bwd_kernel(const float* __restrict__ params, const float* __restrict__ hist, const float* __restrict__ hid,
                                        const float* __restrict__ d_ab, int B, float* __restrict__ d_params) {
  const int b = blockIdx.x, t = threadIdx.x;
  __shared__ float gcur[32], gnext[32], x[32];
  const float* w3 = params + 32 * 16 + 2 * 32 * 32;
  float* dw3 = d_params + 32 * 16 + 2 * 32 * 32;
  gcur[t] = t < 12 ? d_ab[b * 12 + t] : 0.f;
  __syncwarp();
  // output layer [16 x 32]: rows 12..15 are padding (no gradient)
  const float* h3 = hid + (b * 3 + 2) * 32;
  for (int o = 0; o < 12; ++o) atomicAdd(dw3 + o * 32 + t, gcur[o] * h3[t]);
  float s = 0.f;
  for (int o = 0; o < 12; ++o) s += w3[o * 32 + t] * gcur[o];
  gnext[t] = h3[t] > 0.f ? s : 0.f;
  __syncwarp();
  for (int l = 2; l >= 1; --l) {                      // hidden layers 2, 1 ([32 x 32]), inputs h_{l}
    const float* w = params + 32 * 16 + (l - 1) * 32 * 32;
    float* dw = d_params + 32 * 16 + (l - 1) * 32 * 32;
    const float* hin = hid + (b * 3 + (l - 1)) * 32;
    gcur[t] = gnext[t];
    __syncwarp();
    for (int o = 0; o < 32; ++o) atomicAdd(dw + o * 32 + t, gcur[o] * hin[t]);
    s = 0.f;
    for (int o = 0; o < 32; ++o) s += w[o * 32 + t] * gcur[o];
    __syncwarp();
    gnext[t] = hin[t] > 0.f ? s : 0.f;
    __syncwarp();
  }
  x[t] = t < 10 ? truncf(hist[b * 10 + t]) : (t < 16 ? 1.f : 0.f);
  gcur[t] = gnext[t];
  __syncwarp();
  if (t < 16)
    for (int o = 0; o < 32; ++o) atomicAdd(d_params + o * 16 + t, gcur[o] * x[t]);
}

}  // namespace nefes

extern "C" {

int64_t nefes_fusion_workspace(int64_t n_pixels) { return n_pixels > 0 ? nefes::fusion_ws_floats(n_pixels) * (int64_t)sizeof(float) : 0; }

int nefes_fusion_fwd(const nefes_fusion_params_t* p, const float* rgb, const float* feat, int B, int H, int W, int training, int no_bn,
                     int residual, float momentum, float eps, void* workspace, float* out, void* stream) {
  using namespace nefes;
  NEFES_REQUIRE(p && rgb && feat && workspace && out, NEFES_EINVAL, "nefes_fusion_fwd: null pointer");
  NEFES_REQUIRE(B > 0 && H > 0 && W > 0, NEFES_EINVAL, "nefes_fusion_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  for (int l = 0; l < 4; ++l) NEFES_REQUIRE(p->weight[l] && p->bias[l], NEFES_EINVAL, "nefes_fusion_fwd: missing parameters of conv %d", l);
  NEFES_REQUIRE(no_bn || (p->bn_weight && p->bn_bias && p->bn_running_mean && p->bn_running_var), NEFES_EINVAL, "nefes_fusion_fwd: missing BatchNorm tensors");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = (int64_t)B * H * W;
  FusionWs w = fusion_carve(workspace, P);
  fusion_pack_kernel<<<grid_for(P * kFusIn), 256, 0, st>>>(rgb, feat, P, w.X0);
  NEFES_CHECK_LAUNCH("fusion_pack");
  if (int e = conv_fwd(st, w.X0, B, H, W, kFusIn, 3, p->weight[0], p->bias[0], kFusHid, ACT_RELU, w.colL[0], w.A1)) return e;
  if (int e = conv_fwd(st, w.A1, B, H, W, kFusHid, 3, p->weight[1], p->bias[1], kFusHid, ACT_RELU, w.colL[1], w.A2)) return e;
  if (int e = conv_fwd(st, w.A2, B, H, W, kFusHid, 3, p->weight[2], p->bias[2], kFusHid, ACT_RELU, w.colL[2], w.A3)) return e;
  if (int e = conv_fwd(st, w.A3, B, H, W, kFusHid, 5, p->weight[3], p->bias[3], kFeat, ACT_NONE, w.colL[3], w.Y)) return e;
  if (!no_bn) {
    if (training) {
      NEFES_CUDA(cudaMemsetAsync(w.sums, 0, 2 * kFeat * sizeof(float), st));
      const int64_t rpb = 64;
      chan_sums_kernel<<<(unsigned)ceil_div(P, rpb), kFeat, 0, st>>>(w.Y, nullptr, P, kFeat, rpb, w.sums);
      NEFES_CHECK_LAUNCH("bn_sums");
    }
    bn_finalize_kernel<<<1, kFeat, 0, st>>>(w.sums, P, kFeat, training, eps, momentum, const_cast<float*>(p->bn_running_mean),
                                            const_cast<float*>(p->bn_running_var), w.stat);
    NEFES_CHECK_LAUNCH("bn_finalize");
  }
  bn_apply_kernel<<<grid_for(P * kFeat), 256, 0, st>>>(w.Y, no_bn ? nullptr : w.stat, p->bn_weight, p->bn_bias, residual ? feat : nullptr, P,
                                                       kFeat, out);
  NEFES_CHECK_LAUNCH("bn_apply");
  return NEFES_OK;
}

int nefes_fusion_bwd(const nefes_fusion_params_t* p, const nefes_fusion_grads_t* g, const float* d_out, int B, int H, int W, int training,
                     int no_bn, int residual, void* workspace, float* d_rgb, float* d_feat, void* stream) {
  using namespace nefes;
  NEFES_REQUIRE(p && g && d_out && workspace, NEFES_EINVAL, "nefes_fusion_bwd: null pointer");
  NEFES_REQUIRE(B > 0 && H > 0 && W > 0, NEFES_EINVAL, "nefes_fusion_bwd: bad shape B=%d H=%d W=%d", B, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = (int64_t)B * H * W;
  FusionWs w = fusion_carve(workspace, P);
  const float* dY = d_out;
  if (!no_bn) {
    NEFES_CUDA(cudaMemsetAsync(w.sums, 0, 2 * kFeat * sizeof(float), st));
    const int64_t rpb = 64;
    chan_sums_kernel<<<(unsigned)ceil_div(P, rpb), kFeat, 0, st>>>(d_out, w.Y, P, kFeat, rpb, w.sums);
    NEFES_CHECK_LAUNCH("bn_bwd_sums");
    bn_bwd_kernel<<<grid_for(P * kFeat), 256, 0, st>>>(d_out, w.Y, w.stat, p->bn_weight, w.sums, P, kFeat, training, w.G2, g->bn_weight, g->bn_bias);
    NEFES_CHECK_LAUNCH("bn_bwd");
    dY = w.G2;
  }
  // conv4 <- A3, conv3 <- A2, conv2 <- A1, conv1 <- X0; G holds the gradient of the layer being processed's INPUT
  const bool need_in = d_rgb != nullptr || d_feat != nullptr;
  if (int e = conv_bwd(st, w.colL[3], dY, B, H, W, kFusHid, 5, p->weight[3], kFeat, w.col, g->weight[3], g->bias[3], w.A3, w.G)) return e;
  // the three 64-channel gradients alternate between G and G2 (G2 is free once conv4 has consumed dY)
  if (int e = conv_bwd(st, w.colL[2], w.G, B, H, W, kFusHid, 3, p->weight[2], kFusHid, w.col, g->weight[2], g->bias[2], w.A2, w.G2)) return e;
  if (int e = conv_bwd(st, w.colL[1], w.G2, B, H, W, kFusHid, 3, p->weight[1], kFusHid, w.col, g->weight[1], g->bias[1], w.A1, w.G)) return e;
  if (int e = conv_bwd(st, w.colL[0], w.G, B, H, W, kFusIn, 3, p->weight[0], kFusHid, w.col, g->weight[0], g->bias[0], nullptr, need_in ? w.G2 : nullptr)) return e;
  if (need_in) {
    fusion_unpack_grad_kernel<<<grid_for(P * kFusIn), 256, 0, st>>>(w.G2, P, d_rgb, d_feat, residual ? d_out : nullptr);
    NEFES_CHECK_LAUNCH("fusion_unpack_grad");
  }
  return NEFES_OK;
}

int nefes_affine_color_fwd(const float* exposure_params, const float* hist, const float* rgb, int B, int64_t n_per_image, float* ab12,
                           float* hidden96, float* out, void* stream) {
  using namespace nefes;
  NEFES_REQUIRE(exposure_params && hist && rgb && ab12 && hidden96 && out, NEFES_EINVAL, "nefes_affine_color_fwd: null pointer");
  NEFES_REQUIRE(B > 0 && n_per_image > 0, NEFES_EINVAL, "nefes_affine_color_fwd: bad shape B=%d n=%lld", B, (long long)n_per_image);
  cudaStream_t st = (cudaStream_t)stream;
  exposure_mlp_kernel<<<B, 32, 0, st>>>(exposure_params, hist, B, ab12, hidden96);
  NEFES_CHECK_LAUNCH("exposure_mlp");
  const int64_t N = (int64_t)B * n_per_image;
  affine_color_fwd_kernel<<<grid_for(N), 256, 0, st>>>(rgb, ab12, n_per_image, N, out);
  NEFES_CHECK_LAUNCH("affine_color_fwd");
  return NEFES_OK;
}

int nefes_affine_color_bwd(const float* exposure_params, const float* hist, const float* rgb, const float* out, const float* d_out,
                           const float* ab12, const float* hidden96, int B, int64_t n_per_image, float* d_ab12, float* d_rgb,
                           float* d_exposure_params, void* stream) {
  using namespace nefes;
  NEFES_REQUIRE(exposure_params && hist && rgb && out && d_out && ab12 && hidden96 && d_ab12, NEFES_EINVAL, "nefes_affine_color_bwd: null pointer");
  NEFES_REQUIRE(B > 0 && n_per_image > 0, NEFES_EINVAL, "nefes_affine_color_bwd: bad shape B=%d n=%lld", B, (long long)n_per_image);
  cudaStream_t st = (cudaStream_t)stream;
  NEFES_CUDA(cudaMemsetAsync(d_ab12, 0, (size_t)B * 12 * sizeof(float), st));
  const int64_t N = (int64_t)B * n_per_image;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(n_per_image, 256), 64), (unsigned)B);
  affine_color_bwd_kernel<<<grid, 256, 0, st>>>(rgb, out, d_out, ab12, n_per_image, N, d_rgb, d_ab12);
  NEFES_CHECK_LAUNCH("affine_color_bwd");
  if (d_exposure_params != nullptr) {
    exposure_mlp_bwd_kernel<<<B, 32, 0, st>>>(exposure_params, hist, hidden96, d_ab12, B, d_exposure_params);
    NEFES_CHECK_LAUNCH("exposure_mlp_bwd");
  }
  return NEFES_OK;
}

}  // extern "C"
