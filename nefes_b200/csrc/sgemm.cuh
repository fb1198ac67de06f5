// fp32 SIMT GEMM used by the NEFES_PREC_FP32 (parity) path of the field MLP.
//   C[I,J] = epilogue( sum_r A(i,r) * B(r,j) )
// Operands are addressed with a "reduction-contiguous" flag so the same kernel serves
//   forward  (A = activations [I,R],   B = W [J,R] row-major -> both r-contiguous),
//   dgrad    (A = dD [I,R],            B = W [R,J]           -> A r-contig, B j-contig),
//   wgrad    (A = dD^T: dD [R,I],      B = act [R,J]         -> both i/j-contig, split over R,
//             atomicAdd epilogue).
// 128 x BJ x 16 CTA tile, 256 threads, 8 x (BJ/16) register tile, float4 shared-memory reads.
#pragma once
#include "common.cuh"

namespace nefes {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SOFTPLUS = 2, ACT_SIGMOID = 3, ACT_THEADS = 4 };

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  const float* bias;               // [J] or null
  const float* mask; int64_t ldm;  // keep where mask[i,j] > 0 (ReLU backward) or null
  int64_t I; int J; int64_t R;
  int act, accumulate, atomic;
  int64_t r_chunk;                 // > 0: blockIdx.z owns [z*r_chunk, (z+1)*r_chunk)
};

template <int BJ, bool A_RC, bool B_RC>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  constexpr int BI = 128, BR = 16, TJ = BJ / 16;
  static_assert(BJ == 128 || BJ == 64, "BJ");
  __shared__ __align__(16) float As[BR][BI + 4];
  __shared__ __align__(16) float Bs[BR][BJ + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * BI;
  const int j0 = blockIdx.y * BJ;
  int64_t r_begin = 0, r_end = g.R;
  if (g.r_chunk > 0) {
    r_begin = (int64_t)blockIdx.z * g.r_chunk;
    r_end = min(g.R, r_begin + g.r_chunk);
  }
  float acc[8][TJ];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < TJ; ++b) acc[a][b] = 0.f;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += BR) {
    if (A_RC) {
      const int r = tid & 15;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = (tid >> 4) + 16 * q;
        const bool ok = (i0 + i < g.I) && (r0 + r < r_end);
        As[r][i] = ok ? g.A[(i0 + i) * g.lda + (r0 + r)] : 0.f;
      }
    } else {
      const int i = tid & 127;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = (tid >> 7) + 2 * q;
        const bool ok = (i0 + i < g.I) && (r0 + r < r_end);
        As[r][i] = ok ? g.A[(r0 + r) * g.lda + (i0 + i)] : 0.f;
      }
    }
    if (B_RC) {
      const int r = tid & 15;
#pragma unroll
      for (int q = 0; q < BJ / 16; ++q) {
        const int j = (tid >> 4) + 16 * q;
        const bool ok = (j0 + j < g.J) && (r0 + r < r_end);
        Bs[r][j] = ok ? g.B[(int64_t)(j0 + j) * g.ldb + (r0 + r)] : 0.f;
      }
    } else {
      const int j = tid & (BJ - 1);
#pragma unroll
      for (int q = 0; q < BJ / 16; ++q) {
        const int r = tid / BJ + (256 / BJ) * q;
        const bool ok = (j0 + j < g.J) && (r0 + r < r_end);
        Bs[r][j] = ok ? g.B[(r0 + r) * g.ldb + (j0 + j)] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      float a[8], b[TJ];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[r][64 + ty * 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
      if (TJ == 8) *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[r][64 + tx * 4]);
#pragma unroll
      for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < TJ; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int x = 0; x < 8; ++x) {
    const int64_t i = i0 + (x < 4 ? ty * 4 + x : 64 + ty * 4 + (x - 4));
    if (i >= g.I) continue;
#pragma unroll
    for (int y = 0; y < TJ; ++y) {
      const int j = j0 + (y < 4 ? tx * 4 + y : 64 + tx * 4 + (y - 4));
      if (j >= g.J) continue;
      float v = acc[x][y];
      float* dst = g.C + i * g.ldc + j;
      if (g.atomic) { atomicAdd(dst, v); continue; }
      if (g.accumulate) v += *dst;
      if (g.bias) v += g.bias[j];
      switch (g.act) {
        case ACT_RELU: v = fmaxf(v, 0.f); break;
        case ACT_SOFTPLUS: v = softplus_f(v); break;
        case ACT_SIGMOID: v = sigmoid_f(v); break;
        case ACT_THEADS: v = (j < 3) ? sigmoid_f(v) : softplus_f(v); break;
        default: break;
      }
      if (g.mask) v = (g.mask[i * g.ldm + j] > 0.f) ? v : 0.f;
      *dst = v;
    }
  }
}

template <bool A_RC, bool B_RC>
inline int launch_sgemm(const GemmArgs& g, cudaStream_t st, const char* what) {
  if (g.I <= 0 || g.J <= 0 || g.R <= 0) return NEFES_OK;
  unsigned gz = g.r_chunk > 0 ? (unsigned)ceil_div(g.R, g.r_chunk) : 1u;
  if (g.J > 64) {
    dim3 grid((unsigned)ceil_div(g.I, 128), (unsigned)ceil_div(g.J, 128), gz);
    sgemm_kernel<128, A_RC, B_RC><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid((unsigned)ceil_div(g.I, 128), 1, gz);
    sgemm_kernel<64, A_RC, B_RC><<<grid, 256, 0, st>>>(g);
  }
  NEFES_CHECK_LAUNCH(what);
  return NEFES_OK;
}

}  // namespace nefes
