// Packed-layer table of the bf16 tensor-core path.
//
// Data layout ("image"): every activation / gradient tensor with C channels is stored per tile of
// 128 points as the UMMA interleaved operand image [C/8][128][8] bf16 (C*256 bytes per tile), so a tile
// moves global <-> shared with ONE linear bulk copy and is directly an MMA operand (see tc05.cuh).
// Channel concatenation == image concatenation, so the skip input [xyzPE | h4] and the direction
// input [final | dirPE] are two bulk copies into adjacent shared memory.
//
// A packed layer is what one GEMM launch computes; it may merge reference layers that share an input
// (final + sigma, dir + transient_encoding.0, the three transient heads).  Padded rows/columns are zero.
#pragma once
#include "common.cuh"

namespace nefes {

enum PackedLayer {
  PL_T0 = 0, PL_T1, PL_T2, PL_T3, PL_T4, PL_T5, PL_T6, PL_T7,
  PL_FS,      // [final(128) ; sigma(1) ; 0 x15] <- h8            N=144 K=128
  PL_SIG,     // [sigma(1) ; 0 x15]              <- h8            N=16  K=128   (sigma-only mode)
  PL_DT,      // [dir(64) ; tenc0(64)]           <- [final|dirPE] N=128 K=160   (fine)
  PL_DIR,     // [dir(64)]                       <- [final|dirPE] N=64  K=160   (coarse)
  PL_RGB,     // [rgb+feat(131) ; 0 x13]         <- dir hidden    N=144 K=64
  PL_TE1,     // transient_encoding.2                              N=64  K=64
  PL_TE2,     // transient_encoding.4                              N=64  K=64
  PL_TH,      // [t_rgb(3) ; t_sigma ; t_beta ; 0 x11] <- t3       N=16  K=64
  PL_DX,      // backward only: d xyzPE(64) <- [G5 | G1] through W_T4[:, :63] and W_T0   N=64 K=256
  PL_COUNT
};

struct PackedDims { int N, K; };
__host__ __device__ constexpr PackedDims packed_dims(int pl) {
  return pl <= PL_T7 ? (pl == PL_T0 ? PackedDims{128, 64} : pl == PL_T4 ? PackedDims{128, 192} : PackedDims{128, 128})
       : pl == PL_FS ? PackedDims{144, 128} : pl == PL_SIG ? PackedDims{16, 128}
       : pl == PL_DT ? PackedDims{128, 160} : pl == PL_DIR ? PackedDims{64, 160}
       : pl == PL_RGB ? PackedDims{144, 64} : pl == PL_TH ? PackedDims{16, 64}
       : pl == PL_DX ? PackedDims{64, 256} : PackedDims{64, 64};
}

// flat-parameter offsets needed to (un)pack: filled on the host from Layout
struct PackSrc {
  int64_t w[NEFES_MAX_LAYERS];
  int64_t b[NEFES_MAX_LAYERS];
  int fine;
};

// Map (packed layer, padded row n, padded col k) -> flat index of the weight, or -1 for padding.
__host__ __device__ inline int64_t packed_weight_index(const PackSrc& S, int pl, int n, int k) {
  switch (pl) {
    case PL_T0: return (k < kXyzCh) ? S.w[L_T0] + (int64_t)n * kXyzCh + k : -1;
    case PL_T4: {                          // input order: xyz PE (63) first, then h (nerfh_nff.py:552)
      if (k < kXyzCh) return S.w[L_T4] + (int64_t)n * 191 + k;
      if (k < 64) return -1;
      return S.w[L_T4] + (int64_t)n * 191 + kXyzCh + (k - 64);
    }
    case PL_T1: case PL_T2: case PL_T3: case PL_T5: case PL_T6: case PL_T7:
      return S.w[L_T0 + (pl - PL_T0)] + (int64_t)n * 128 + k;
    case PL_FS:
      if (n < 128) return S.w[L_FINAL] + (int64_t)n * 128 + k;
      return n == 128 ? S.w[L_SIGMA] + k : -1;
    case PL_SIG: return n == 0 ? S.w[L_SIGMA] + k : -1;
    case PL_DT: case PL_DIR:               // dir (64 rows) then tenc0 (64 rows) are adjacent in the flat buffer
      return (k < 155) ? S.w[L_DIR] + (int64_t)n * 155 + k : -1;
    case PL_RGB: return (n < kHeadCh) ? S.w[L_RGB] + (int64_t)n * 64 + k : -1;
    case PL_TE1: return S.w[L_TENC1] + (int64_t)n * 64 + k;
    case PL_TE2: return S.w[L_TENC2] + (int64_t)n * 64 + k;
    case PL_TH: return (n < 5) ? S.w[L_TRGB] + (int64_t)n * 64 + k : -1;   // t_rgb, t_sigma, t_beta adjacent
    case PL_DX:                            // n = xyz-PE channel, k = output unit of layer 5 (k<128) / layer 1
      if (n >= kXyzCh) return -1;
      return k < 128 ? S.w[L_T4] + (int64_t)k * 191 + n : S.w[L_T0] + (int64_t)(k - 128) * kXyzCh + n;
  }
  return -1;
}
__host__ __device__ inline int64_t packed_bias_index(const PackSrc& S, int pl, int n) {
  switch (pl) {
    case PL_FS: return n < 128 ? S.b[L_FINAL] + n : (n == 128 ? S.b[L_SIGMA] : -1);
    case PL_SIG: return n == 0 ? S.b[L_SIGMA] : -1;
    case PL_DT: case PL_DIR: return S.b[L_DIR] + n;
    case PL_RGB: return n < kHeadCh ? S.b[L_RGB] + n : -1;
    case PL_TE1: return S.b[L_TENC1] + n;
    case PL_TE2: return S.b[L_TENC2] + n;
    case PL_TH: return n < 5 ? S.b[L_TRGB] + n : -1;
    case PL_DX: return -1;
    default: return S.b[L_T0 + (pl - PL_T0)] + n;
  }
}

// Packed weight arena (bf16 images + fp32 padded biases), rebuilt from the flat fp32 parameters
// before every forward (the optimiser changes them every step).  For each packed layer:
//   W  image  [K/8][N][8]   B operand of forward   (D[pts,N]  = A[pts,K]  W^T)
//   WT image  [N/8][K][8]   B operand of dgrad     (dA[pts,K] = G[pts,N]  W)
struct PackedArena {
  int64_t w_off[PL_COUNT], wt_off[PL_COUNT];    // in bf16 elements
  int64_t bias_off[PL_COUNT];                   // in floats, inside the bias block
  int64_t n_bf16, n_bias;
  int64_t bytes;                                // whole arena: [bf16 block | bias block]
};
inline PackedArena packed_arena() {
  PackedArena a = {};
  int64_t e = 0, b = 0;
  for (int pl = 0; pl < PL_COUNT; ++pl) {
    const PackedDims d = packed_dims(pl);
    a.w_off[pl] = e; e += (int64_t)d.N * d.K;
    a.wt_off[pl] = e; e += (int64_t)d.N * d.K;
    a.bias_off[pl] = b; b += d.N;
  }
  a.n_bf16 = e; a.n_bias = b;
  a.bytes = round_up(e * 2, 256) + b * 4;
  return a;
}

}  // namespace nefes
