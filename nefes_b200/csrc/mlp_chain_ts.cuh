// Forward layer chain with the ACTIVATIONS IN TENSOR MEMORY (tcgen05.mma with the A operand in TMEM): shared memory
// holds nothing but the weight ring.
//
// Why: in the shared-memory-operand chain (mlp_chain.cuh) every tile-layer moves 32 KB of activations into shared
// memory (epilogue stores), reads them back twice (MMA operand fetch, bulk store of the saved copy) and reads 32 KB of
// weights -- the shared-memory pipe is the measured limiter (MMAs retire at ~175 cycles instead of 107 while the
// epilogue of the other tile runs).  Here the epilogue writes the bf16 activations of a row straight back into the
// row's TMEM lane (tcgen05.st), the next layer's MMA reads them from there, and the saved copy leaves from registers.
//
//   TMEM of one tile (256 columns): [ accumulator 0..143 | xyzPE 144..175 | H 176..239 | dirPE 240..255 ]  (bf16 pairs)
//   so the skip input [xyzPE | h4] (K = 192) and the direction input [final | dirPE] (K = 160) are column ranges.
//   warp 0 weight producer (4-slot ring), warps 1 / 18 MMA issuers of tile 0 / 1, warps 2..17 epilogue (8 per tile).
// Included by mlp_tc.cu.   script/models/nerfh_nff.py:525-576.
#pragma once

namespace nefes {

constexpr int kTsSlots = 4;
constexpr uint32_t kTsWSlot = 49152;
constexpr uint32_t kTsSmem = kTsSlots * kTsWSlot + kChainBiasBytes;
constexpr uint32_t kTsX = 144, kTsH = 176, kTsD = 240;          // TMEM columns of the operand regions

struct TsArgs {
  ChainStep step[kChainMaxSteps];      // a_off / out_off are TMEM COLUMNS here
  int n_steps;
  int64_t M; int n_tiles;
  float* raw; int C;
  const uint8_t* x_img; const uint8_t* d_img;     // encodings (bf16 images) written by encode_images_kernel
  int x_dead_step;                                // the xyzPE columns may be refilled once this step's MMAs retired
  long long* dbg;                                 // optional clock stamps of CTA 0 (NEFES_CHAIN_DBG)
  int xflags;                                     // timing experiments (NEFES_CHAIN_X): 1 no saves
};

__global__ void __launch_bounds__(kChainThreads, 1) chain_fwd_ts_kernel(const __grid_constant__ TsArgs A) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_wfull[kTsSlots], bar_wempty[kTsSlots], bar_act[2], bar_acc[2];
  __shared__ uint32_t tmem_slot;
  __shared__ ChainStep s_step[kChainMaxSteps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sW = smem;
  float* sBias = reinterpret_cast<float*>(smem + kTsSlots * kTsWSlot);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTsSlots; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_act[i], kChainEpiWarps * 32); mbar_init(&bar_acc[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  for (int s = threadIdx.x; s < A.n_steps; s += kChainThreads) s_step[s] = A.step[s];
  for (int s = 0; s < A.n_steps; ++s)
    if (A.step[s].bias != nullptr)
      for (int i = threadIdx.x; i < A.step[s].N; i += kChainThreads) sBias[A.step[s].bias_off + i] = A.step[s].bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_pairs = (A.n_tiles + 1) >> 1;
  const int n_steps = A.n_steps;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
        for (int s = 0; s < n_steps; ++s, ++cnt) {
          const int slot = cnt % kTsSlots;
          mbar_wait(&bar_wempty[slot], ((cnt / kTsSlots) & 1) ^ 1);
          const uint32_t bytes = s_step[s].w_bytes;
          mbar_arrive_expect_tx(&bar_wfull[slot], bytes);
          const uint8_t* src = s_step[s].w_img;
          for (uint32_t off = 0; off < bytes; off += 16384u)
            bulk_g2s(sW + slot * kTsWSlot + off, src + off, min(16384u, bytes - off), &bar_wfull[slot]);
        }
    }
  } else if (warp == 1 || warp == 2 + 2 * kChainEpiWarps) {
    if (lane == 0) {
      const int g = warp == 1 ? 0 : 1;
      uint32_t cnt = 0, act_ph = 0u;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const bool valid = pair * 2 + g < A.n_tiles;
        for (int s = 0; s < n_steps; ++s, ++cnt) {
          const ChainStep& st = s_step[s];
          const int slot = cnt % kTsSlots;
          mbar_wait(&bar_wfull[slot], (cnt / kTsSlots) & 1);
          if (valid) {
            const uint32_t idesc = idesc_bf16(128, st.N, 0, 0);
            const uint64_t db0 = smem_desc(smem_u32(sW + slot * kTsWSlot), st.w_lbo, 128);
            mbar_wait(&bar_act[g], act_ph);
            act_ph ^= 1u;
            tc_fence_after();
            if (A.dbg && blockIdx.x == 0 && cnt < 32) A.dbg[cnt * 16 + g * 2] = clock64();
            const uint32_t d = tmem + g * 256, a0 = tmem + g * 256 + st.a_off;
            for (int k = 0; k < st.K / 16; ++k)
              mma_ts(d, a0 + k * 8, db0 + (uint64_t)(k * (2 * st.w_lbo >> 4)), idesc, k > 0 ? 1u : 0u);
            mma_commit(&bar_acc[g]);
            if (A.dbg && blockIdx.x == 0 && cnt < 32) A.dbg[cnt * 16 + g * 2 + 1] = clock64();
          }
          mma_commit(&bar_wempty[slot]);
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int g = ew >> 3;
    const int half = (ew >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tbase = tmem + g * 256 + ((uint32_t)(q * 32) << 16);     // this thread's TMEM lane, tile g
    uint32_t acc_ph = 0u, ecnt = 0;
    const bool dbg = A.dbg != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0;
    uint4 xv[4], dv[2];
    auto fetch_inputs = [&](int tile) {             // this row's halves of the encoding images -> registers
      const uint8_t* xp = A.x_img + (int64_t)tile * (64 * 256) + (half * 4) * kChunkBytes + row * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = __ldg(reinterpret_cast<const uint4*>(xp + j * kChunkBytes));
      if (A.d_img != nullptr) {
        const uint8_t* dp = A.d_img + (int64_t)tile * (32 * 256) + (half * 2) * kChunkBytes + row * 16;
#pragma unroll
        for (int j = 0; j < 2; ++j) dv[j] = __ldg(reinterpret_cast<const uint4*>(dp + j * kChunkBytes));
      }
    };
    auto store_inputs = [&]() {                     // registers -> the operand columns of this row
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) { w[4 * j] = xv[j].x; w[4 * j + 1] = xv[j].y; w[4 * j + 2] = xv[j].z; w[4 * j + 3] = xv[j].w; }
      tmem_st16(tbase + kTsX + half * 16, w);
      if (A.d_img != nullptr) {
        uint32_t w8[8] = {dv[0].x, dv[0].y, dv[0].z, dv[0].w, dv[1].x, dv[1].y, dv[1].z, dv[1].w};
        tmem_st8(tbase + kTsD + half * 8, w8);
      }
      tmem_st_wait();
      tc_fence_before();
    };
    {
      const int tile0 = blockIdx.x * 2 + g;
      if (tile0 < A.n_tiles) fetch_inputs(tile0);
    }
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int tile = pair * 2 + g;
      if (tile >= A.n_tiles) break;
      const int64_t grow = (int64_t)tile * kTile + row;
      const bool ok = grow < A.M;
      float* rawt = A.raw + (int64_t)tile * A.C * kTile + row;
      store_inputs();
      mbar_arrive(&bar_act[g]);
      const int next_tile = (pair + (int)gridDim.x) * 2 + g;

      for (int s = 0; s < n_steps; ++s, ++ecnt) {
        const ChainStep& st = s_step[s];
        const float* bias = sBias + st.bias_off;
        if (s == n_steps - 1 && next_tile < A.n_tiles) fetch_inputs(next_tile);   // overlap with the last epilogue
        mbar_wait(&bar_acc[g], acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
        if (dbg && ecnt < 32) A.dbg[ecnt * 16 + 4] = clock64();
        if (st.kind == CK_HIDDEN || st.kind == CK_FS) {
          const int ncol = st.out_ch >> 1;             // accumulator columns of this warp: 32 or 64
          const int c_base = half * ncol;
          uint8_t* gdst_row = (st.gdst && !(A.xflags & 1)) ? st.gdst + (int64_t)tile * st.g_tile_stride + (c_base >> 3) * kChunkBytes + row * 16 : nullptr;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (h2 * 32 < ncol) {
              uint32_t v[32], w[16];
              tmem_ld32(tbase + c_base + h2 * 32, v);
              tmem_ld_wait();
              if (dbg && ecnt < 32) A.dbg[ecnt * 16 + 5 + h2 * 2] = clock64();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias + c_base + h2 * 32 + 8 * j);
                const float4 b1 = *reinterpret_cast<const float4*>(bias + c_base + h2 * 32 + 8 * j + 4);
                if (st.kind == CK_HIDDEN) {
                  w[4 * j + 0] = bias_pack<true>(v[8 * j + 0], v[8 * j + 1], b0.x, b0.y);
                  w[4 * j + 1] = bias_pack<true>(v[8 * j + 2], v[8 * j + 3], b0.z, b0.w);
                  w[4 * j + 2] = bias_pack<true>(v[8 * j + 4], v[8 * j + 5], b1.x, b1.y);
                  w[4 * j + 3] = bias_pack<true>(v[8 * j + 6], v[8 * j + 7], b1.z, b1.w);
                } else {
                  w[4 * j + 0] = bias_pack<false>(v[8 * j + 0], v[8 * j + 1], b0.x, b0.y);
                  w[4 * j + 1] = bias_pack<false>(v[8 * j + 2], v[8 * j + 3], b0.z, b0.w);
                  w[4 * j + 2] = bias_pack<false>(v[8 * j + 4], v[8 * j + 5], b1.x, b1.y);
                  w[4 * j + 3] = bias_pack<false>(v[8 * j + 6], v[8 * j + 7], b1.z, b1.w);
                }
                if (gdst_row != nullptr)
                  *reinterpret_cast<uint4*>(gdst_row + (h2 * 4 + j) * kChunkBytes) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
              }
              // the operand of the next layer: packed pairs back into this row's lane (its last reader, the MMA, is done)
              tmem_st16(tbase + st.out_off + (c_base >> 1) + h2 * 16, w);
              if (dbg && ecnt < 32) A.dbg[ecnt * 16 + 6 + h2 * 2] = clock64();
            }
          }
          if (st.kind == CK_FS && half == 1) {         // sigma pre-activation rides in column 128 of this GEMM
            uint32_t v[16];
            tmem_ld16(tbase + 128, v);
            tmem_ld_wait();
            if (ok) rawt[131 * kTile] = softplus_f(__uint_as_float(v[0]) + bias[128]);
          }
          tmem_st_wait();
        } else if (st.kind == CK_HEADS || st.kind == CK_SIGMA) {
          if (half == 0) {
            uint32_t v[16];
            tmem_ld16(tbase, v);
            tmem_ld_wait();
            if (ok) {
              if (st.kind == CK_SIGMA) {
                rawt[0] = softplus_f(__uint_as_float(v[0]) + bias[0]);
              } else {
#pragma unroll
                for (int e = 0; e < 5; ++e) {
                  const float x = __uint_as_float(v[e]) + bias[e];
                  rawt[(132 + e) * kTile] = e < 3 ? sigmoid_f(x) : softplus_f(x);
                }
              }
            }
          }
        } else {                                       // CK_RGB: 131 fp32 columns; half 0 -> [0, 80), half 1 -> [80, 131)
          const int cb = half * 80;
          const int nblk = half == 0 ? 5 : 4;
#pragma unroll 1
          for (int b2 = 0; b2 < nblk; ++b2) {
            uint32_t v[16];
            tmem_ld16(tbase + cb + b2 * 16, v);
            tmem_ld_wait();
            const int c0 = cb + b2 * 16;
            if (ok) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (c0 + e < kHeadCh) rawt[(c0 + e) * kTile] = __uint_as_float(v[e]) + bias[c0 + e];
            }
          }
        }
        tc_fence_before();
        if (dbg && ecnt < 32) A.dbg[ecnt * 16 + 9] = clock64();
        if (s + 1 < n_steps) mbar_arrive(&bar_act[g]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace nefes
