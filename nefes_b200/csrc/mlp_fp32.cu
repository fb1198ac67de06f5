// K5, NEFES_PREC_FP32 path: PE + the NeFeS MLP layer by layer on the SIMT fp32 GEMM.
// This is the parity anchor (<= 1e-3 rel of the reference, in practice ~1e-5) and the numerical
// reference the tcgen05 bf16 path is debugged against.  script/models/nerfh_nff.py:168-231,
// :525-576.  Forward keeps every post-activation tensor backward needs in `saved`.
#include "tf32_gemm.cuh"

extern "C" int nefes_encode_pe_fwd(const float*, int64_t, int, float*, int, void*);
extern "C" int nefes_encode_pe_bwd(const float*, const float*, int, int64_t, int, float*, void*);

namespace nefes {

// ---- workspace carving --------------------------------------------------------------------
struct Saved {            // fp32, row-major, M rows unless noted
  float* E;               // [M,64]   xyz PE (63 used)
  float* H[8];            // [M,128]  trunk activations h1..h8
  float* FD;              // [M,160]  [xyz_encoding_final | dir PE 27 | pad]
  float* DT;              // [M,ldDT] [dir hidden 64 | transient hidden-0 64]
  float* T2; float* T3;   // [M,64]
  int ldDT;
  int64_t bytes;
};
static Saved carve_saved(void* base, int64_t M, int mode) {
  Saved s = {};
  float* p = (float*)base;
  auto take = [&](int64_t n) { float* q = p; p += n; return q; };
  s.E = take(M * 64);
  for (int l = 0; l < 8; ++l) s.H[l] = take(M * 128);
  if (mode != NEFES_MODE_SIGMA) {
    s.FD = take(M * 160);
    s.ldDT = (mode == NEFES_MODE_FULL) ? 128 : 64;
    s.DT = take(M * s.ldDT);
    if (mode == NEFES_MODE_FULL) { s.T2 = take(M * 64); s.T3 = take(M * 64); }
  }
  s.bytes = (int64_t)((char*)p - (char*)base);
  return s;
}
struct ScratchBwd {
  float *GA, *GB;         // [M,128] ping-pong
  float *GFD;             // [M,160]
  float *GDT;             // [M,128]
  float *GT2, *GT3;       // [M,64]
  float *G8;              // [M,8]  pre-activation grads: 0..4 transient heads, 5 static sigma
  float *GE;              // [M,64]
  float *GDr;             // [N,32] per-ray dir-PE grads
  int64_t bytes;
};
static ScratchBwd carve_bwd(void* base, int64_t M, int64_t N, int mode) {
  ScratchBwd s = {};
  float* p = (float*)base;
  auto take = [&](int64_t n) { float* q = p; p += n; return q; };
  s.GA = take(M * 128); s.GB = take(M * 128);
  s.G8 = take(M * 8); s.GE = take(M * 64);
  if (mode != NEFES_MODE_SIGMA) {
    s.GFD = take(M * 160); s.GDT = take(M * 128); s.GDr = take(N * 32);
    if (mode == NEFES_MODE_FULL) { s.GT2 = take(M * 64); s.GT3 = take(M * 64); }
  }
  s.bytes = (int64_t)((char*)p - (char*)base);
  return s;
}

// ---- small helper kernels -------------------------------------------------------------------
// FD[m, 128 + c] = dirPE[m / S, c]  (c < 27), zero pad up to 160
__global__ void dirpe_broadcast_kernel(const float* __restrict__ EDr, int S, int64_t M, float* __restrict__ FD) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 32) return;
  const int64_t m = idx >> 5;
  const int c = (int)(idx & 31);
  FD[m * 160 + 128 + c] = (c < kDirCh) ? EDr[(m / S) * 32 + c] : 0.f;
}
// GDr[n, c] = sum_s GFD[n*S + s, 128 + c]
__global__ void dirpe_reduce_kernel(const float* __restrict__ GFD, int S, int64_t N, float* __restrict__ GDr) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 32) return;
  const int64_t n = idx >> 5;
  const int c = (int)(idx & 31);
  float a = 0.f;
  if (c < kDirCh)
    for (int s = 0; s < S; ++s) a += GFD[(n * S + s) * 160 + 128 + c];
  GDr[idx] = a;
}
// pre-activation grads of the activated heads from their OUTPUTS:
//   softplus: dy/dx = 1 - exp(-y);  sigmoid: y (1 - y).
__global__ void head_grad_kernel(const float* __restrict__ raw, const float* __restrict__ d_raw, int C,
                                 int sig_col, int64_t M, float* __restrict__ G8) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* y = raw + m * C;
  const float* g = d_raw + m * C;
  float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  o[5] = g[sig_col] * (1.f - expf(-y[sig_col]));
  if (C == 137) {
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = g[132 + c] * y[132 + c] * (1.f - y[132 + c]);
    o[3] = g[135] * (1.f - expf(-y[135]));
    o[4] = g[136] * (1.f - expf(-y[136]));
  }
  float4* dst = reinterpret_cast<float4*>(G8 + m * 8);
  dst[0] = make_float4(o[0], o[1], o[2], o[3]);
  dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}
// db[j] += sum_i G[i*ld + j], j < J <= 160
__global__ void colsum_kernel(const float* __restrict__ G, int64_t ld, int64_t M, int J, int64_t rows_per_block,
                              float* __restrict__ db) {
  const int j = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t i1 = min(M, i0 + rows_per_block);
  if (j >= J) return;
  float a0 = 0.f, a1 = 0.f;
  int64_t i = i0;
  for (; i + 2 <= i1; i += 2) { a0 += G[i * ld + j]; a1 += G[(i + 1) * ld + j]; }
  if (i < i1) a0 += G[i * ld + j];
  atomicAdd(&db[j], a0 + a1);
}
static int colsum(cudaStream_t st, const float* G, int64_t ld, int64_t M, int J, float* db) {
  const int64_t rpb = 1024;
  colsum_kernel<<<(unsigned)ceil_div(M, rpb), (unsigned)round_up(J, 32), 0, st>>>(G, ld, M, J, rpb, db);
  NEFES_CHECK_LAUNCH("colsum");
  return NEFES_OK;
}

#define TRY(x) do { if (int e__ = (x)) return e__; } while (0)

// ---- forward --------------------------------------------------------------------------------
int mlp_fwd_fp32(const float* P, int net, int mode, const float* pts, const float* dirs, int64_t N,
                 int S, float* raw, void* saved, void* scratch, cudaStream_t st) {
  const Layout& L = layout_for(net);
  const int64_t M = N * S;
  Saved w = carve_saved(saved, M, mode);
  const int C = (mode == NEFES_MODE_SIGMA) ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137);
  auto Wp = [&](int l) { return P + L.w[l]; };
  auto Bp = [&](int l) { return P + L.b[l]; };

  TRY(nefes_encode_pe_fwd(pts, M, kXyzFreqs, w.E, 64, st));
  const float* h = w.E;
  int64_t ldh = 64;
  int K = kXyzCh;
  for (int l = 0; l < 8; ++l) {
    if (l == 4) {  // skip: input is [xyz PE | h4]  (nerfh_nff.py:551-553)
      TRY(linear_fwd(st, w.E, 64, Wp(L_T4), 191, nullptr, w.H[4], 128, M, 128, kXyzCh, ACT_NONE, 0));
      TRY(linear_fwd(st, w.H[3], 128, Wp(L_T4) + kXyzCh, 191, Bp(L_T4), w.H[4], 128, M, 128, 128, ACT_RELU, 1));
    } else {
      TRY(linear_fwd(st, h, ldh, Wp(L_T0 + l), K, Bp(L_T0 + l), w.H[l], 128, M, 128, K, ACT_RELU, 0));
    }
    h = w.H[l]; ldh = 128; K = 128;
  }
  const int sig_col = (mode == NEFES_MODE_SIGMA) ? 0 : 131;
  TRY(linear_fwd(st, w.H[7], 128, Wp(L_SIGMA), 128, Bp(L_SIGMA), raw + sig_col, C, M, 1, 128, ACT_SOFTPLUS, 0));
  if (mode == NEFES_MODE_SIGMA) return NEFES_OK;

  float* EDr = (float*)scratch;                                   // [N,32]
  TRY(nefes_encode_pe_fwd(dirs, N, kDirFreqs, EDr, 32, st));
  dirpe_broadcast_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, st>>>(EDr, S, M, w.FD);
  NEFES_CHECK_LAUNCH("dirpe_broadcast");
  TRY(linear_fwd(st, w.H[7], 128, Wp(L_FINAL), 128, Bp(L_FINAL), w.FD, 160, M, 128, 128, ACT_NONE, 0));
  // dir_encoding (and transient_encoding.0: adjacent rows, same input) -> DT
  TRY(linear_fwd(st, w.FD, 160, Wp(L_DIR), 155, Bp(L_DIR), w.DT, w.ldDT, M, w.ldDT, 155, ACT_RELU, 0));
  TRY(linear_fwd(st, w.DT, w.ldDT, Wp(L_RGB), 64, Bp(L_RGB), raw, C, M, kHeadCh, 64, ACT_NONE, 0));
  if (mode == NEFES_MODE_FULL) {
    TRY(linear_fwd(st, w.DT + 64, 128, Wp(L_TENC1), 64, Bp(L_TENC1), w.T2, 64, M, 64, 64, ACT_RELU, 0));
    TRY(linear_fwd(st, w.T2, 64, Wp(L_TENC2), 64, Bp(L_TENC2), w.T3, 64, M, 64, 64, ACT_RELU, 0));
    // transient rgb(3, sigmoid) | sigma(1, softplus) | beta(1, softplus) -> raw[:,132:137]
    TRY(linear_fwd(st, w.T3, 64, Wp(L_TRGB), 64, Bp(L_TRGB), raw + 132, C, M, 5, 64, ACT_THEADS, 0));
  }
  return NEFES_OK;
}

// ---- backward -------------------------------------------------------------------------------
int mlp_bwd_fp32(const float* P, int net, int mode, const float* pts, const float* dirs, int64_t N,
                 int S, const float* raw, const float* d_raw, const void* saved, void* scratch,
                 float* dP, float* d_pts, float* d_dirs, cudaStream_t st) {
  const Layout& L = layout_for(net);
  const int64_t M = N * S;
  const Saved w = carve_saved(const_cast<void*>(saved), M, mode);
  ScratchBwd s = carve_bwd(scratch, M, N, mode);
  const int C = (mode == NEFES_MODE_SIGMA) ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137);
  auto Wp = [&](int l) { return P + L.w[l]; };
  auto dW = [&](int l) { return dP + L.w[l]; };
  auto dB = [&](int l) { return dP + L.b[l]; };
  const bool wg = dP != nullptr;
  const bool need_in = d_pts != nullptr;       // gradient to the sample positions (pose refinement)

  const int sig_col = (mode == NEFES_MODE_SIGMA) ? 0 : 131;
  head_grad_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(raw, d_raw, C, sig_col, M, s.G8);
  NEFES_CHECK_LAUNCH("head_grad");

  float* g = s.GA;          // gradient w.r.t. h8 (post-ReLU masked) is built here
  if (mode != NEFES_MODE_SIGMA) {
    const int nDT = w.ldDT;                    // 64 (static) or 128 (full)
    if (mode == NEFES_MODE_FULL) {
      if (wg) {
        TRY(linear_wgrad(st, s.G8, 8, w.T3, 64, dW(L_TRGB), 64, M, 5, 64));
        TRY(colsum(st, s.G8, 8, M, 5, dB(L_TRGB)));
      }
      TRY(linear_dgrad(st, s.G8, 8, Wp(L_TRGB), 64, s.GT3, 64, M, 5, 64, w.T3, 64, 0));
      if (wg) {
        TRY(linear_wgrad(st, s.GT3, 64, w.T2, 64, dW(L_TENC2), 64, M, 64, 64));
        TRY(colsum(st, s.GT3, 64, M, 64, dB(L_TENC2)));
      }
      TRY(linear_dgrad(st, s.GT3, 64, Wp(L_TENC2), 64, s.GT2, 64, M, 64, 64, w.T2, 64, 0));
      if (wg) {
        TRY(linear_wgrad(st, s.GT2, 64, w.DT + 64, 128, dW(L_TENC1), 64, M, 64, 64));
        TRY(colsum(st, s.GT2, 64, M, 64, dB(L_TENC1)));
      }
      TRY(linear_dgrad(st, s.GT2, 64, Wp(L_TENC1), 64, s.GDT + 64, 128, M, 64, 64, w.DT + 64, 128, 0));
    }
    // static rgb/feature head: d_raw[:, :131] (no activation)
    if (wg) {
      TRY(linear_wgrad(st, d_raw, C, w.DT, nDT, dW(L_RGB), 64, M, kHeadCh, 64));
      TRY(colsum(st, d_raw, C, M, kHeadCh, dB(L_RGB)));
    }
    TRY(linear_dgrad(st, d_raw, C, Wp(L_RGB), 64, s.GDT, nDT, M, kHeadCh, 64, w.DT, nDT, 0));
    // dir_encoding (+ transient_encoding.0)
    if (wg) {
      TRY(linear_wgrad(st, s.GDT, nDT, w.FD, 160, dW(L_DIR), 155, M, nDT, 155));
      TRY(colsum(st, s.GDT, nDT, M, nDT, dB(L_DIR)));
    }
    const bool need_dir = d_dirs != nullptr;
    TRY(linear_dgrad(st, s.GDT, nDT, Wp(L_DIR), 155, s.GFD, 160, M, nDT, need_dir ? 155 : 128, nullptr, 0, 0));
    if (need_dir) {
      dirpe_reduce_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(s.GFD, S, N, s.GDr);
      NEFES_CHECK_LAUNCH("dirpe_reduce");
      TRY(nefes_encode_pe_bwd(dirs, s.GDr, 32, N, kDirFreqs, d_dirs, st));
    }
    // xyz_encoding_final
    if (wg) {
      TRY(linear_wgrad(st, s.GFD, 160, w.H[7], 128, dW(L_FINAL), 128, M, 128, 128));
      TRY(colsum(st, s.GFD, 160, M, 128, dB(L_FINAL)));
    }
    TRY(linear_dgrad(st, s.GFD, 160, Wp(L_FINAL), 128, g, 128, M, 128, 128, nullptr, 0, 0));
  }
  // static sigma head joins at h8; the ReLU mask of h8 is applied once, on the sum
  if (wg) {
    TRY(linear_wgrad(st, s.G8 + 5, 8, w.H[7], 128, dW(L_SIGMA), 128, M, 1, 128));
    TRY(colsum(st, s.G8 + 5, 8, M, 1, dB(L_SIGMA)));
  }
  TRY(linear_dgrad(st, s.G8 + 5, 8, Wp(L_SIGMA), 128, g, 128, M, 1, 128, w.H[7], 128,
                   mode != NEFES_MODE_SIGMA ? 1 : 0));

  // trunk, layer 8 down to layer 1
  float* gin = g;
  float* gout = s.GB;
  for (int l = 7; l >= 0; --l) {
    if (l == 4) {
      if (wg) {
        TRY(linear_wgrad(st, gin, 128, w.E, 64, dW(L_T4), 191, M, 128, kXyzCh));
        TRY(linear_wgrad(st, gin, 128, w.H[3], 128, dW(L_T4) + kXyzCh, 191, M, 128, 128));
        TRY(colsum(st, gin, 128, M, 128, dB(L_T4)));
      }
      if (need_in) TRY(linear_dgrad(st, gin, 128, Wp(L_T4), 191, s.GE, 64, M, 128, kXyzCh, nullptr, 0, 0));
      TRY(linear_dgrad(st, gin, 128, Wp(L_T4) + kXyzCh, 191, gout, 128, M, 128, 128, w.H[3], 128, 0));
    } else if (l == 0) {
      if (wg) {
        TRY(linear_wgrad(st, gin, 128, w.E, 64, dW(L_T0), kXyzCh, M, 128, kXyzCh));
        TRY(colsum(st, gin, 128, M, 128, dB(L_T0)));
      }
      if (need_in) TRY(linear_dgrad(st, gin, 128, Wp(L_T0), kXyzCh, s.GE, 64, M, 128, kXyzCh, nullptr, 0, 1));
      break;
    } else {
      if (wg) {
        TRY(linear_wgrad(st, gin, 128, w.H[l - 1], 128, dW(L_T0 + l), 128, M, 128, 128));
        TRY(colsum(st, gin, 128, M, 128, dB(L_T0 + l)));
      }
      TRY(linear_dgrad(st, gin, 128, Wp(L_T0 + l), 128, gout, 128, M, 128, 128, w.H[l - 1], 128, 0));
    }
    float* t = gin; gin = gout; gout = t;
  }
  if (need_in) TRY(nefes_encode_pe_bwd(pts, s.GE, 64, M, kXyzFreqs, d_pts, st));
  return NEFES_OK;
}

int mlp_workspace_fp32(int mode, int64_t M, int64_t N, int64_t* saved, int64_t* sf, int64_t* sb) {
  *saved = carve_saved(nullptr, M, mode).bytes;
  *sf = N * 32 * (int64_t)sizeof(float) + 256;
  *sb = carve_bwd(nullptr, M, N, mode).bytes + 256;
  return NEFES_OK;
}

int gemm_mode_set(int tf32) {
  const int prev = gemm_tf32();
  gemm_tf32() = tf32 ? 1 : 0;
  return prev;
}

}  // namespace nefes

extern "C" {
int nefes_gemm_mode(int tf32) { return nefes::gemm_mode_set(tf32); }
int nefes_linear_fwd(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc, int64_t M,
                     int N, int K, int act, void* stream) {
  if (M == 0) return NEFES_OK;
  NEFES_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, NEFES_EINVAL, "nefes_linear_fwd: bad argument");
  NEFES_REQUIRE(act == nefes::ACT_NONE || act == nefes::ACT_RELU || act == nefes::ACT_SIGMOID || act == nefes::ACT_SOFTPLUS,
                NEFES_EINVAL, "nefes_linear_fwd: bad activation %d", act);
  return nefes::linear_fwd((cudaStream_t)stream, A, lda, W, K, bias, C, ldc, M, N, K, act, 0);
}
int nefes_linear_dgrad(const float* dC, int64_t ldc, const float* W, float* dA, int64_t lda, int64_t M, int N, int K,
                       const float* mask, int64_t ldm, void* stream) {
  if (M == 0) return NEFES_OK;
  NEFES_REQUIRE(dC && W && dA && M > 0 && N > 0 && K > 0, NEFES_EINVAL, "nefes_linear_dgrad: bad argument");
  return nefes::linear_dgrad((cudaStream_t)stream, dC, ldc, W, K, dA, lda, M, N, K, mask, ldm, 0);
}
int nefes_linear_wgrad(const float* dC, int64_t ldc, const float* A, int64_t lda, float* dW, int64_t M, int N, int K,
                       void* stream) {
  if (M == 0) return NEFES_OK;
  NEFES_REQUIRE(dC && A && dW && M > 0 && N > 0 && K > 0, NEFES_EINVAL, "nefes_linear_wgrad: bad argument");
  return nefes::linear_wgrad((cudaStream_t)stream, dC, ldc, A, lda, dW, K, M, N, K);
}
}  // extern "C"
