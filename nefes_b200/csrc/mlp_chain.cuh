// Fused forward layer chain of the NeFeS field (bf16 tensor path): one persistent kernel runs PE and ALL
// layers of the MLP for a pair of 128-point tiles; activations never leave the SM between layers.
//
//   warp 0      weight producer: streams each layer's W image L2 -> shared (bulk copies, 2-slot ring)
//   warp 1      MMA issuer: for every layer, tile 0 then tile 1 (tcgen05.mma M=128, accumulators in TMEM,
//               256 columns per tile); the tensor pipe works on one tile while the other tile's epilogue runs
//   warps 2-5   epilogue of tile 0, warps 6-9 epilogue of tile 1: positional encoding -> shared operand image;
//               per layer TMEM -> registers -> bias / ReLU / bit-mask -> bf16 operand image of the NEXT layer in
//               shared memory (in place) -> the same image is bulk-stored to HBM once for backward
//
// Per-tile shared region (56 KB): [ xyzPE 16 KB | H 32 KB | dirPE 8 KB ] -- contiguous so the skip input
// [xyzPE | h4] (K=192) and the direction input [final | dirPE] (K=160) are plain operand ranges.
// Included by mlp_tc.cu (uses its helpers).
#pragma once

namespace nefes {

enum { CK_HIDDEN = 0,   // bias + ReLU + mask -> image
       CK_PLAIN = 1,    // bias -> image (no activation)
       CK_FS = 2,       // bias -> image (128 ch: xyz_encoding_final) + column 128: softplus -> raw[:, sig_col]
       CK_HEADS = 3,    // columns 0..4: sigmoid x3, softplus x2 -> raw[:, 132..136]
       CK_SIGMA = 4,    // column 0: softplus -> raw[:, 0]                       (sigma-only mode)
       CK_RGB = 5 };    // bias -> 131 fp32 columns -> raw[:, 0..130] (staged, coalesced)

struct ChainStep {
  uint32_t a_off;                // operand start inside the tile region (bytes)
  uint32_t out_off;              // image destination inside the tile region (bytes)
  uint16_t K, N, out_ch;
  uint8_t kind, pad;
  uint32_t w_bytes;
  const uint8_t* w_img;          // [K/8][N][8] bf16
  const float* bias;             // [N]
  uint8_t* gdst;                 // saved activation image (or null)
  uint32_t g_tile_stride;
  uint4* mask;                   // ReLU bit-mask destination (or null)
};
constexpr int kChainMaxSteps = 14;
struct ChainArgs {
  ChainStep step[kChainMaxSteps];
  int n_steps;
  const float* pts; const float* dirs; int S; int64_t M; int n_tiles;
  float* raw; int C; int sig_col;
  uint8_t* x_img; uint8_t* d_img;            // saved xyzPE / dirPE images (wgrad operands)
};

constexpr uint32_t kRegX = 0, kRegH = 16384, kRegD = 49152, kRegBytes = 57344;
constexpr uint32_t kChainWSlot = 49152;      // largest W image: 192 x 128 bf16
constexpr int kChainThreads = 64 + 256;
constexpr int kChainBiasStride = 160;        // floats per step in the shared bias table (N <= 160)
constexpr uint32_t kChainSmem = 2 * kRegBytes + 2 * kChainWSlot + 10240;  // + bias table (14 x 160 floats)

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// 16 accumulator columns of one row -> two 8-channel chunks of the image at `dst`
template <bool RELU>
__device__ __forceinline__ uint32_t chain_block(const uint32_t (&v)[16], const float* __restrict__ bias, uint8_t* __restrict__ dst,
                                                uint8_t* __restrict__ gdst, int c0, int row) {
  uint32_t bits = 0u;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = c0 + h * 8;
    const float4 b0 = *reinterpret_cast<const float4*>(bias + c);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + c + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      x[e] = __uint_as_float(v[h * 8 + e]) + bb[e];
      if (RELU) {
        const bool on = x[e] > 0.f;
        x[e] = on ? x[e] : 0.f;
        bits |= on ? (1u << (h * 8 + e)) : 0u;
      }
    }
    uint4 pk;
    pk.x = pack_bf16(x[0], x[1]); pk.y = pack_bf16(x[2], x[3]);
    pk.z = pack_bf16(x[4], x[5]); pk.w = pack_bf16(x[6], x[7]);
    *reinterpret_cast<uint4*>(dst + (c >> 3) * kChunkBytes + row * 16) = pk;
    // saved for backward: same image layout in HBM; a warp's 32 rows x 16 B are one contiguous 512-byte store
    if (gdst != nullptr) *reinterpret_cast<uint4*>(gdst + (c >> 3) * kChunkBytes + row * 16) = pk;
  }
  return bits;
}

__global__ void __launch_bounds__(kChainThreads, 1) chain_fwd_kernel(const ChainArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_wfull[2], bar_wempty[2], bar_act[2], bar_acc[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sW = smem + 2 * kRegBytes;
  float* sBias = reinterpret_cast<float*>(smem + 2 * kRegBytes + 2 * kChainWSlot);   // [n_steps][...] packed below

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 1);
      mbar_init(&bar_act[i], 128); mbar_init(&bar_acc[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  // all biases of the chain -> shared
  for (int s = 0; s < A.n_steps; ++s)
    for (int i = threadIdx.x; i < A.step[s].N; i += kChainThreads) sBias[s * kChainBiasStride + i] = A.step[s].bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_pairs = (A.n_tiles + 1) >> 1;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
        for (int s = 0; s < A.n_steps; ++s, ++cnt) {
          const int slot = cnt & 1;
          mbar_wait(&bar_wempty[slot], ((cnt >> 1) & 1) ^ 1);
          const uint32_t bytes = A.step[s].w_bytes;
          mbar_arrive_expect_tx(&bar_wfull[slot], bytes);
          for (uint32_t off = 0; off < bytes; off += 16384u)
            bulk_g2s(sW + slot * kChainWSlot + off, A.step[s].w_img + off, min(16384u, bytes - off), &bar_wfull[slot]);
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t cnt = 0, act_ph[2] = {0u, 0u};
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const bool valid1 = pair * 2 + 1 < A.n_tiles;
        for (int s = 0; s < A.n_steps; ++s, ++cnt) {
          const ChainStep& st = A.step[s];
          const int slot = cnt & 1;
          mbar_wait(&bar_wfull[slot], (cnt >> 1) & 1);
          const uint32_t idesc = idesc_bf16(128, st.N, 0, 0);
          const uint32_t w_lbo = (uint32_t)st.N * 16u;
          const uint64_t db0 = smem_desc(smem_u32(sW + slot * kChainWSlot), w_lbo, 128);
          for (int g = 0; g < 2; ++g) {
            if (g == 1 && !valid1) break;
            mbar_wait(&bar_act[g], act_ph[g]);
            act_ph[g] ^= 1u;
            tc_fence_after();
            const uint64_t da0 = smem_desc(smem_u32(smem + g * kRegBytes + st.a_off), kChunkBytes, 128);
            const uint32_t d = tmem + g * 256;
            for (int k = 0; k < st.K / 16; ++k)
              mma_ss(d, da0 + (uint64_t)(k * (2 * kChunkBytes >> 4)), db0 + (uint64_t)(k * (2 * w_lbo >> 4)), idesc, k > 0);
            mma_commit(&bar_acc[g]);
          }
          mma_commit(&bar_wempty[slot]);
        }
      }
    }
  } else {
    const int g = (warp - 2) >> 2;                    // tile of the pair this warp serves
    const int q = warp & 3;                           // TMEM lane quarter
    const int row = q * 32 + lane;
    const int gt = ((warp - 2) & 3) * 32 + lane;      // 0..127 inside the group
    uint8_t* reg = smem + g * kRegBytes;
    const uint32_t taddr = tmem + g * 256 + ((uint32_t)(q * 32) << 16);
    uint32_t acc_ph = 0u;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int tile = pair * 2 + g;
      if (tile >= A.n_tiles) break;
      const int64_t grow = (int64_t)tile * kTile + row;
      const bool ok = grow < A.M;
      // ---- positional encodings of this row -> operand images (also saved to HBM for wgrad) ------------------
      // sin/cos of the base frequency with the accurate sincosf, higher octaves by the double-angle recurrence
      // (error doubles per octave: <= 2^9 * 1e-7 = 5e-5, far below the bf16 operand rounding of 4e-3).
      {
        float e[64];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = ok ? A.pts[grow * 3 + c] : 0.f;
          e[c] = v;
          float sn, cs;
          sincosf(v, &sn, &cs);
#pragma unroll
          for (int l = 0; l < kXyzFreqs; ++l) {
            e[3 + 6 * l + c] = ok ? sn : 0.f;
            e[6 + 6 * l + c] = ok ? cs : 0.f;
            const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn;
            sn = s2; cs = c2;
          }
        }
        e[63] = 0.f;
        uint8_t* gx = A.x_img + (int64_t)tile * (64 * 256) + row * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 pk;
          pk.x = pack_bf16(e[8 * j], e[8 * j + 1]); pk.y = pack_bf16(e[8 * j + 2], e[8 * j + 3]);
          pk.z = pack_bf16(e[8 * j + 4], e[8 * j + 5]); pk.w = pack_bf16(e[8 * j + 6], e[8 * j + 7]);
          *reinterpret_cast<uint4*>(reg + kRegX + j * kChunkBytes + row * 16) = pk;
          *reinterpret_cast<uint4*>(gx + j * kChunkBytes) = pk;
        }
      }
      if (A.d_img != nullptr) {
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = ok ? A.dirs[(grow / A.S) * 3 + c] : 0.f;
          e[c] = v;
          float sn, cs;
          sincosf(v, &sn, &cs);
#pragma unroll
          for (int l = 0; l < kDirFreqs; ++l) {
            e[3 + 6 * l + c] = ok ? sn : 0.f;
            e[6 + 6 * l + c] = ok ? cs : 0.f;
            const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn;
            sn = s2; cs = c2;
          }
        }
        uint8_t* gd = A.d_img + (int64_t)tile * (32 * 256) + row * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          pk.x = pack_bf16(e[8 * j], e[8 * j + 1]); pk.y = pack_bf16(e[8 * j + 2], e[8 * j + 3]);
          pk.z = pack_bf16(e[8 * j + 4], e[8 * j + 5]); pk.w = pack_bf16(e[8 * j + 6], e[8 * j + 7]);
          *reinterpret_cast<uint4*>(reg + kRegD + j * kChunkBytes + row * 16) = pk;
          *reinterpret_cast<uint4*>(gd + j * kChunkBytes) = pk;
        }
      }
      fence_async_smem();
      mbar_arrive(&bar_act[g]);                        // operand of step 0 is ready

      for (int s = 0; s < A.n_steps; ++s) {
        const ChainStep& st = A.step[s];
        const float* bias = sBias + s * kChainBiasStride;
        mbar_wait(&bar_acc[g], acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
        if (st.kind <= CK_FS) {
          // the image is written in place: its last reader (the MMA that just completed) is done
          uint8_t* dst = reg + st.out_off;
          uint8_t* gdst = st.gdst ? st.gdst + (int64_t)tile * st.g_tile_stride : nullptr;
          uint16_t* mrow = st.mask ? reinterpret_cast<uint16_t*>(st.mask + grow) : nullptr;
          for (int b0 = 0; b0 < (st.out_ch >> 4); b0 += 4) {
            uint32_t v0[16], v1[16], v2[16], v3[16];
            tmem_ld16(taddr + b0 * 16, v0);
            tmem_ld16(taddr + (b0 + 1) * 16, v1);
            tmem_ld16(taddr + (b0 + 2) * 16, v2);
            tmem_ld16(taddr + (b0 + 3) * 16, v3);
            tmem_ld_wait();
            uint32_t m0, m1, m2, m3;
            if (st.kind == CK_HIDDEN) {
              m0 = chain_block<true>(v0, bias, dst, gdst, b0 * 16, row);
              m1 = chain_block<true>(v1, bias, dst, gdst, (b0 + 1) * 16, row);
              m2 = chain_block<true>(v2, bias, dst, gdst, (b0 + 2) * 16, row);
              m3 = chain_block<true>(v3, bias, dst, gdst, (b0 + 3) * 16, row);
              if (mrow != nullptr) *reinterpret_cast<uint2*>(mrow + b0) = make_uint2(m0 | (m1 << 16), m2 | (m3 << 16));
            } else {
              chain_block<false>(v0, bias, dst, gdst, b0 * 16, row);
              chain_block<false>(v1, bias, dst, gdst, (b0 + 1) * 16, row);
              chain_block<false>(v2, bias, dst, gdst, (b0 + 2) * 16, row);
              chain_block<false>(v3, bias, dst, gdst, (b0 + 3) * 16, row);
            }
          }
          if (st.kind == CK_FS) {
            uint32_t v[16];
            tmem_ld16(taddr + 128, v);
            tmem_ld_wait();
            if (ok) A.raw[grow * A.C + A.sig_col] = softplus_f(__uint_as_float(v[0]) + bias[128]);
          }
          tc_fence_before();
          fence_async_smem();
        } else if (st.kind == CK_HEADS || st.kind == CK_SIGMA) {
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tmem_ld_wait();
          if (ok) {
            if (st.kind == CK_SIGMA) {
              A.raw[grow * A.C] = softplus_f(__uint_as_float(v[0]) + bias[0]);
            } else {
#pragma unroll
              for (int e = 0; e < 5; ++e) {
                const float x = __uint_as_float(v[e]) + bias[e];
                A.raw[grow * A.C + 132 + e] = e < 3 ? sigmoid_f(x) : softplus_f(x);
              }
            }
          }
          tc_fence_before();
        } else {   // CK_RGB: 131 fp32 columns, staged through the (now dead) tile region in two halves of 66 columns
          group_barrier(g);
          float* stage = reinterpret_cast<float*>(reg);
          constexpr int kHalf = 66, kPitch = 67;
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {
            const int c_lo = hf * kHalf, c_hi = hf == 0 ? kHalf : kHeadCh;
#pragma unroll 1
            for (int b = (c_lo >> 4); b * 16 < c_hi; ++b) {
              uint32_t v[16];
              tmem_ld16(taddr + b * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int c = b * 16 + e;
                if (c >= c_lo && c < c_hi) stage[row * kPitch + (c - c_lo)] = __uint_as_float(v[e]) + bias[c];
              }
            }
            group_barrier(g);
            const int ncol = c_hi - c_lo;
            for (int rr = (warp - 2) & 3; rr < kTile; rr += 4) {
              const int64_t gr = (int64_t)tile * kTile + rr;
              if (gr < A.M)
                for (int c = lane; c < ncol; c += 32) A.raw[gr * A.C + c_lo + c] = stage[rr * kPitch + c];
            }
            group_barrier(g);
          }
          tc_fence_before();
        }
        if (s + 1 < A.n_steps) mbar_arrive(&bar_act[g]);   // operand of the next step is ready, accumulator drained
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace nefes
