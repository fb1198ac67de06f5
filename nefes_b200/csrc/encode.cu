// K4a frequency positional encoding (script/models/nerfh_nff.py:241-270):
//   out = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], each block 3 wide.
// sin/cos use the full-range sinf/cosf (arguments reach 512*|x|), never the fast intrinsics.
#include "common.cuh"

namespace nefes {

__global__ void pe_fwd_kernel(const float* __restrict__ x, int64_t M, int L, float* __restrict__ out, int ld) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (row, coord)
  if (idx >= M * 3) return;
  const int64_t m = idx / 3;
  const int c = (int)(idx % 3);
  const float v = x[idx];
  float* o = out + m * ld + c;
  o[0] = v;
  float f = 1.f;
  for (int l = 0; l < L; ++l, f *= 2.f) {
    float s, co;
    sincosf(v * f, &s, &co);
    o[3 + 6 * l] = s;
    o[6 + 6 * l] = co;
  }
}

// d_x = g_x + sum_l 2^l (g_sin cos(2^l x) - g_cos sin(2^l x))
__global__ void pe_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, int ld, int64_t M,
                              int L, float* __restrict__ d_x) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 3) return;
  const int64_t m = idx / 3;
  const int c = (int)(idx % 3);
  const float v = x[idx];
  const float* gi = g + m * ld + c;
  float acc = gi[0];
  float f = 1.f;
  for (int l = 0; l < L; ++l, f *= 2.f) {
    float s, co;
    sincosf(v * f, &s, &co);
    acc += f * (gi[3 + 6 * l] * co - gi[6 + 6 * l] * s);
  }
  d_x[idx] = acc;
}

}  // namespace nefes

extern "C" {

int nefes_encode_pe_fwd(const float* x, int64_t M, int n_freqs, float* out, int ld_out, void* stream) {
  NEFES_REQUIRE(x && out, NEFES_EINVAL, "nefes_encode_pe_fwd: null pointer");
  NEFES_REQUIRE(M >= 0 && n_freqs >= 0 && n_freqs <= 16 && ld_out >= 3 + 6 * n_freqs, NEFES_EINVAL,
                "nefes_encode_pe_fwd: bad shape (M=%lld L=%d ld=%d)", (long long)M, n_freqs, ld_out);
  if (M == 0) return NEFES_OK;
  nefes::pe_fwd_kernel<<<(unsigned)nefes::ceil_div(M * 3, 256), 256, 0, (cudaStream_t)stream>>>(
      x, M, n_freqs, out, ld_out);
  NEFES_CHECK_LAUNCH("pe_fwd");
  return NEFES_OK;
}

int nefes_encode_pe_bwd(const float* x, const float* d_out, int ld_out, int64_t M, int n_freqs,
                        float* d_x, void* stream) {
  NEFES_REQUIRE(x && d_out && d_x, NEFES_EINVAL, "nefes_encode_pe_bwd: null pointer");
  NEFES_REQUIRE(M >= 0 && n_freqs >= 0 && n_freqs <= 16 && ld_out >= 3 + 6 * n_freqs, NEFES_EINVAL,
                "nefes_encode_pe_bwd: bad shape");
  if (M == 0) return NEFES_OK;
  nefes::pe_bwd_kernel<<<(unsigned)nefes::ceil_div(M * 3, 256), 256, 0, (cudaStream_t)stream>>>(
      x, d_out, ld_out, M, n_freqs, d_x);
  NEFES_CHECK_LAUNCH("pe_bwd");
  return NEFES_OK;
}

}  // extern "C"
