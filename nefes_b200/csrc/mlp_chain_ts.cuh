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
//
// Round 2 (measured with tools/tc_issue.cu and the NEFES_CHAIN_DBG stamps): (1) the issuer and producer warps stay
// CONVERGED and issue from one elected lane with uniform operands -- a stream of N=128 MMAs with A in TMEM then retires at
// 64.7 cycles/MMA, the tensor-pipe floor, against 79 from a lane-0 branch (ptxas wraps each MMA in an ELECT / R2UR /
// BRA.U.ANY "waterfall" there) and 107 with A in shared memory (operand-fetch bound); (2) the epilogue is straight-line:
// step fields in registers, st.global spelled out, the second 32-column block loaded while the first is packed, the
// operand store ahead of the HBM copy; (3) the two tiles of a pair run in anti-phase (bar_stag).
// Included by mlp_tc.cu.   script/models/nerfh_nff.py:525-576.
#pragma once

namespace nefes {

constexpr int kTsSlots = 3;
constexpr uint32_t kTsWSlot = 49152;
constexpr uint32_t kTsStage = 32768;             // per tile: the bf16 image of the layer just computed, on its way to HBM
constexpr uint32_t kTsSmem = kTsSlots * kTsWSlot + 2 * kTsStage + kChainBiasBytes;
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

// ---- epilogue helpers -------------------------------------------------------------------------------------------------
// st.global spelled out: a generic store through a pointer the compiler cannot place makes it assume the store may alias
// shared memory and re-load every shared value (step table fields, biases) after each one -- measured in round 1's
// variant as four dependent LDS + branch chains per 32-column block.
__device__ __forceinline__ void stg128(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ void stg32f(float* p, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v)); }
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void lds_bias32(const float* bp, float4 (&b)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(bp + 4 * j);
}
template <bool RELU>
__device__ __forceinline__ void pack32(const uint32_t (&v)[32], const float4 (&b)[8], uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    w[4 * j + 0] = bias_pack<RELU>(v[8 * j + 0], v[8 * j + 1], b[2 * j].x, b[2 * j].y);
    w[4 * j + 1] = bias_pack<RELU>(v[8 * j + 2], v[8 * j + 3], b[2 * j].z, b[2 * j].w);
    w[4 * j + 2] = bias_pack<RELU>(v[8 * j + 4], v[8 * j + 5], b[2 * j + 1].x, b[2 * j + 1].y);
    w[4 * j + 3] = bias_pack<RELU>(v[8 * j + 6], v[8 * j + 7], b[2 * j + 1].z, b[2 * j + 1].w);
  }
}
__device__ __forceinline__ void save32(uint8_t* g, const uint32_t (&w)[16]) {       // four 8-channel chunks of one row
#pragma unroll
  for (int j = 0; j < 4; ++j) stg128(g + j * kChunkBytes, w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

// One hidden-layer epilogue of a warp: its 32 rows x ncol accumulator columns [c_base, c_base + ncol) -> bias (+ ReLU) ->
// bf16 pairs -> the operand columns of the next layer (tcgen05.st).  The second 32-column block is loaded while the first
// is packed and stored.  The packed words come back in w0 / w1: the caller writes the saved copy to HBM AFTER it has
// handed the operand over -- mbarrier.arrive has release semantics, so an arrive issued behind the global stores waits
// until they have drained (measured: ~950 cycles between the last epilogue warp and the issuer under HBM load).
template <bool RELU>
__device__ __forceinline__ void epi_hidden(uint32_t tbase, int c_base, int ncol, uint32_t out_col, const float* bp,
                                           uint32_t (&w0)[16], uint32_t (&w1)[16]) {
  float4 b[8];
  uint32_t v[32];
  lds_bias32(bp, b);
  tmem_ld32(tbase + c_base, v);
  tmem_ld_wait();
  pack32<RELU>(v, b, w0);
  if (ncol == 64) tmem_ld32(tbase + c_base + 32, v);
  tmem_st16(tbase + out_col + (c_base >> 1), w0);
  if (ncol == 64) {
    lds_bias32(bp + 32, b);
    tmem_ld_wait();
    pack32<RELU>(v, b, w1);
    tmem_st16(tbase + out_col + (c_base >> 1) + 16, w1);
  }
  tmem_st_wait();
}

__global__ void __launch_bounds__(kChainThreads, 1) chain_fwd_ts_kernel(const __grid_constant__ TsArgs A) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_wfull[kTsSlots], bar_wempty[kTsSlots], bar_act[2], bar_acc[2], bar_stag;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sW = smem;
  uint8_t* sStage = smem + kTsSlots * kTsWSlot;
  float* sBias = reinterpret_cast<float*>(smem + kTsSlots * kTsWSlot + 2 * kTsStage);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTsSlots; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_act[i], kChainEpiWarps * 32); mbar_init(&bar_acc[i], 1); }
    mbar_init(&bar_stag, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
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
    // producer: the whole warp runs the loop (converged), one elected lane issues the bulk copies -- see tc05.cuh uni()
    const bool leader = elect_one();
    uint32_t cnt = 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
      for (int s = 0; s < n_steps; ++s, ++cnt) {
        const int slot = cnt % kTsSlots;
        mbar_wait(&bar_wempty[slot], ((cnt / kTsSlots) & 1) ^ 1);
        const uint32_t bytes = A.step[s].w_bytes;
        const uint8_t* src = A.step[s].w_img;
        if (leader) {
          mbar_arrive_expect_tx(&bar_wfull[slot], bytes);
          for (uint32_t off = 0; off < bytes; off += 16384u)
            bulk_g2s(sW + slot * kTsWSlot + off, src + off, min(16384u, bytes - off), &bar_wfull[slot]);
        }
        __syncwarp();
      }
  } else if (warp == 1 || warp == 2 + 2 * kChainEpiWarps) {
    // MMA issuer of tile g: converged warp, uniform operands, one elected lane issues (tc05.cuh uni())
    const int g = uni(warp == 1 ? 0 : 1);
    const uint32_t tm = uni(tmem);
    const bool leader = elect_one();
    const bool dbg = A.dbg != nullptr && blockIdx.x == 0;
    uint32_t cnt = 0, act_ph = 0u;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const bool valid = pair * 2 + g < A.n_tiles;
      for (int s = 0; s < n_steps; ++s, ++cnt) {
        const ChainStep& st = A.step[s];               // kernel parameter, uniform index: uniform loads
        const int slot = cnt % kTsSlots;
        mbar_wait(&bar_wfull[slot], (cnt / kTsSlots) & 1);
        if (dbg && leader && cnt < 32) A.dbg[cnt * 48 + 40 + g] = clock64();      // weights landed
        if (valid) {
          const uint32_t idesc = idesc_bf16(128, st.N, 0, 0);
          const uint64_t db0 = smem_desc(smem_u32(sW + slot * kTsWSlot), st.w_lbo, 128);
          mbar_wait(&bar_act[g], act_ph);
          act_ph ^= 1u;
          tc_fence_after();
          if (dbg && leader && cnt < 32) A.dbg[cnt * 48 + g * 2] = clock64();       // operand ready: issue starts
          const uint32_t d = tm + g * 256, a0 = tm + g * 256 + st.a_off;
          const int ksteps = st.K / 16;
          const uint64_t dbk = (uint64_t)(2 * st.w_lbo >> 4);
          if (leader) {
            for (int k = 0; k < ksteps; ++k)
              mma_ts(d, a0 + k * 8, db0 + (uint64_t)k * dbk, idesc, k > 0 ? 1u : 0u);
            mma_commit(&bar_acc[g]);
          }
          if (dbg && leader && cnt < 32) A.dbg[cnt * 48 + g * 2 + 1] = clock64();   // issued + committed
        }
        // anti-phase: tile 1 of a pair starts when tile 0's first layer has retired, so that from then on one tile is in
        // its epilogue while the other owns the tensor pipe (left alone the two fall into lock-step)
        if (leader && g == 0 && s == 0) mma_commit(&bar_stag);
        if (leader) mma_commit(&bar_wempty[slot]);
        __syncwarp();
      }
    }
  } else {
    const int ew = warp - 2;
    const int g = ew >> 3;
    const int half = (ew >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tbase = tmem + g * 256 + ((uint32_t)(q * 32) << 16);     // this thread's TMEM lane, tile g
    uint32_t acc_ph = 0u, ecnt = 0, pair_it = 0;
    const int gt = (ew & 7) * 32 + lane;            // 0..255 inside the tile's epilogue group
    bool store_pending = false;                     // a bulk store of this group may still be reading the staging image
    const bool dbg = A.dbg != nullptr && blockIdx.x == 0 && lane == 0;
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
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pair_it) {
      const int tile = pair * 2 + g;
      if (tile >= A.n_tiles) break;
      const int64_t grow = (int64_t)tile * kTile + row;
      const bool ok = grow < A.M;
      float* rawt = A.raw + (int64_t)tile * A.C * kTile + row;
      const int next_tile = (pair + (int)gridDim.x) * 2 + g;
      if (pair_it == 0) {                              // later tiles are handed over inside the last step of their predecessor
        store_inputs();
        if (g == 1 && !(A.xflags & 32)) mbar_wait(&bar_stag, 0);   // start half a period after tile 0 (see the MMA issuer)
        mbar_arrive(&bar_act[g]);
      }
      uint8_t* stage = sStage + g * kTsStage;

      for (int s = 0; s < n_steps; ++s, ++ecnt) {
        // the step's fields as locals: later asm statements clobber "memory" and would force re-loads
        const int kind = A.step[s].kind, out_ch = A.step[s].out_ch;
        const uint32_t out_col = A.step[s].out_off;
        const float* bias = sBias + A.step[s].bias_off;
        uint8_t* gdst = (A.xflags & 1) ? nullptr : A.step[s].gdst;
        const uint32_t g_stride = A.step[s].g_tile_stride;
        const bool last = s == n_steps - 1;
        if (last && next_tile < A.n_tiles) fetch_inputs(next_tile);   // overlap with the last epilogue
        mbar_wait(&bar_acc[g], acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
        if (dbg && ecnt < 32) A.dbg[ecnt * 48 + 8 + ew] = clock64();                // accumulator ready, per warp
        // Every branch first moves what the NEXT MMA needs (operand columns written, accumulator drained), hands over
        // (hand_over), and only then issues its global stores.
        auto hand_over = [&]() {
          tc_fence_before();
          if (dbg && ecnt < 32) A.dbg[ecnt * 48 + 24 + ew] = clock64();             // epilogue done, per warp
          if (!last) {
            mbar_arrive(&bar_act[g]);
          } else if (next_tile < A.n_tiles) {          // the next tile of this group: encodings in, first MMA may start
            store_inputs();
            if (g == 1 && !(A.xflags & 32)) mbar_wait(&bar_stag, (pair_it + 1) & 1);
            mbar_arrive(&bar_act[g]);
          }
        };
        if (kind == CK_HIDDEN || kind == CK_FS) {
          const int ncol = out_ch >> 1;                // accumulator columns of this warp: 32 or 64
          const int c_base = half * ncol;
          uint32_t w0[16], w1[16], sg = 0u;
          if (kind == CK_FS) {
            if (half == 1) tmem_ld1(tbase + 128, sg);  // sigma pre-activation rides in column 128 of this GEMM
            epi_hidden<false>(tbase, c_base, ncol, out_col, bias + c_base, w0, w1);
          } else {
            epi_hidden<true>(tbase, c_base, ncol, out_col, bias + c_base, w0, w1);
          }
          hand_over();
          if (gdst != nullptr) {
            // the saved copy leaves through shared memory and ONE bulk store per tile-layer (TMA engine): per-thread
            // st.global of the same bytes stalls the epilogue warps at issue once HBM is the limit (measured 0.79 ms
            // against 0.50 without saves); everything below is behind the hand-over, off the critical path
            if (store_pending) {                       // the previous bulk store must have read the staging image
              if (gt == 0) bulk_wait_read<0>();
              group_barrier(g);
            }
            uint8_t* srow = stage + (c_base >> 3) * kChunkBytes + row * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(srow + j * kChunkBytes) = make_uint4(w0[4 * j], w0[4 * j + 1], w0[4 * j + 2], w0[4 * j + 3]);
            if (ncol == 64) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(srow + (4 + j) * kChunkBytes) = make_uint4(w1[4 * j], w1[4 * j + 1], w1[4 * j + 2], w1[4 * j + 3]);
            }
            fence_async_smem();
            group_barrier(g);
            if (gt == 0) {
              bulk_s2g(gdst + (int64_t)tile * g_stride, stage, (uint32_t)out_ch * 256u);
              bulk_commit();
            }
            store_pending = true;
          }
          if (kind == CK_FS && half == 1 && ok) stg32f(rawt + 131 * kTile, softplus_f(__uint_as_float(sg) + bias[128]));
        } else if (kind == CK_HEADS || kind == CK_SIGMA) {
          uint32_t v[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (half == 0) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tbase) : "memory");
            tmem_ld_wait();
          }
          hand_over();
          if (half == 0 && ok) {
            if (kind == CK_SIGMA) {
              stg32f(rawt, softplus_f(__uint_as_float(v[0]) + bias[0]));
            } else {
#pragma unroll
              for (int e = 0; e < 5; ++e) {
                const float x = __uint_as_float(v[e]) + bias[e];
                stg32f(rawt + (132 + e) * kTile, e < 3 ? sigmoid_f(x) : softplus_f(x));
              }
            }
          }
        } else {                                       // CK_RGB: 131 fp32 columns; half 0 -> [0, 64), half 1 -> [64, 131)
          const int cb = half * 64;
          uint32_t v0[32], v1[32], t4[4] = {0u, 0u, 0u, 0u};
          tmem_ld32(tbase + cb, v0);
          tmem_ld32(tbase + cb + 32, v1);
          if (half == 1)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(t4[0]), "=r"(t4[1]), "=r"(t4[2]), "=r"(t4[3]) : "r"(tbase + 128) : "memory");
          tmem_ld_wait();
          hand_over();
          if (ok && !(A.xflags & 4)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias + cb + 4 * j);
              stg32f(rawt + (cb + 4 * j + 0) * kTile, __uint_as_float(v0[4 * j + 0]) + b4.x);
              stg32f(rawt + (cb + 4 * j + 1) * kTile, __uint_as_float(v0[4 * j + 1]) + b4.y);
              stg32f(rawt + (cb + 4 * j + 2) * kTile, __uint_as_float(v0[4 * j + 2]) + b4.z);
              stg32f(rawt + (cb + 4 * j + 3) * kTile, __uint_as_float(v0[4 * j + 3]) + b4.w);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias + cb + 32 + 4 * j);
              stg32f(rawt + (cb + 32 + 4 * j + 0) * kTile, __uint_as_float(v1[4 * j + 0]) + b4.x);
              stg32f(rawt + (cb + 32 + 4 * j + 1) * kTile, __uint_as_float(v1[4 * j + 1]) + b4.y);
              stg32f(rawt + (cb + 32 + 4 * j + 2) * kTile, __uint_as_float(v1[4 * j + 2]) + b4.z);
              stg32f(rawt + (cb + 32 + 4 * j + 3) * kTile, __uint_as_float(v1[4 * j + 3]) + b4.w);
            }
            if (half == 1) {
#pragma unroll
              for (int e = 0; e < 3; ++e) stg32f(rawt + (128 + e) * kTile, __uint_as_float(t4[e]) + bias[128 + e]);
            }
          }
        }
      }
    }
    if (gt == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace nefes
