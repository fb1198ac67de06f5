// Forward layer chain, round 2: activations in TENSOR MEMORY, every layer split into N-blocks with their own accumulator
// columns and their own hand-over, activations double-buffered -- so that inside ONE tile the epilogue of block 0 runs
// while the tensor pipe computes block 1, and the next layer's first K-steps issue while block 1 is still in its epilogue.
//
// Why (measured, tools/tc_issue.cu + NEFES_CHAIN_DBG stamps): a 128-wide layer of one tile is 8 MMAs =
// 520 cycles of tensor pipe, but the serial chain  issue -> retire (+330) -> tcgen05.ld / bias / pack / tcgen05.st (~1000)
// -> hand-over (+170)  is ~2000 cycles, and two tiles per SM (all that fits TMEM) cannot hide it: they drift into
// lock-step and the pipe idles 70-80 % of the time.  Splitting the layer shortens the chain itself.
//
//   TMEM of tile g (256 columns from g*256):  ACC [0,128)  |  HA [128,192)  |  HB [192,256)
//     ACC  fp32 accumulators, one column range per N-block (the 144-wide colour head spills into HA[0,16), dead by then)
//     HA / HB  the bf16 activation image of a layer (64 columns = 128 channels as pairs): a layer reads one and writes the
//              other, so its block-0 epilogue may store while its block-1 MMAs still read
//   the xyz / direction encodings stay in SHARED memory (bulk-copied by the producer) and enter as SS-mode K-steps
//   (4 of the 12 K-steps of the skip layer, 2 of 10 of the direction layer).
//
//   warp 0       producer: weight ring (bulk copies), the encodings of the next tile pair
//   warp 1 / 18  MMA issuer of tile 0 / 1: CONVERGED warp, uniform operands, one elected lane issues (tc05.cuh uni())
//   warps 2-9 / 10-17  epilogue of tile 0 / 1: warp = TMEM lane quarter x block parity (warps 0-3 of a tile serve block 0 and
//                      a third block, warps 4-7 block 1), all columns of the block in 32-column passes
//
// Synchronisation is by "latest completion" of per-tile mbarriers: acc_ready[g][b] (issuer -> epilogue, block b of the
// current step retired) and k_ready[g][b] (epilogue -> issuer: block b's operand columns written AND its accumulator
// columns drained).  Both sides count completions per barrier from the same step table, and a wait always names the LATEST
// completion so far (parity (n-1)&1), which can never be more than one phase behind.
//
// The saved copy of a layer leaves through a per-tile 32 KB staging image in shared memory and ONE bulk store (TMA
// engine), issued behind the hand-over: per-thread st.global of the same bytes stalls the epilogue warps at issue once
// HBM is the limit, and an mbarrier.arrive (release) behind global stores waits for them to drain.
// Included by mlp_tc.cu.   script/models/nerfh_nff.py:525-576.
#pragma once

namespace nefes {

// ---- epilogue helpers -------------------------------------------------------------------------------------------------
// st.global spelled out: a generic store through a pointer the compiler cannot place makes it assume the store may alias
// shared memory and re-load every shared value (step table fields, biases) after each one -- measured in round 1's
// variant as four dependent LDS + branch chains per 32-column block.
__device__ __forceinline__ void stg128(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ void stg32f(float* p, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v)); }
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
enum { BK_HID_RELU = 0,   // bias + ReLU -> bf16 pairs -> out_col (+ staging image)
       BK_HID = 1,        // bias        -> bf16 pairs -> out_col (+ staging image)       (xyz_encoding_final)
       BK_RAW = 2,        // bias -> fp32 raw channels [raw_c0, raw_c0 + raw_n)            (rgb + feature head)
       BK_SIGMA = 3,      // column 0: bias, softplus -> raw[raw_c0]
       BK_HEADS = 4 };    // columns 0..4: sigmoid x3, softplus x2 -> raw[raw_c0 .. +5)
enum { GS_TMEM = 0, GS_X = 1, GS_D = 2 };
enum { W_K0 = 1, W_K1 = 2, W_K2 = 4, W_X = 8, W_D = 16 };

struct Ts2Block {
  uint16_t n0, nw;          // rows [n0, n0 + nw) of the weight image; MMA N = nw (multiple of 16)
  uint16_t acc_col;         // accumulator column of the block
  uint16_t out_col;         // HID: TMEM column receiving the bf16 pairs of channels [n0, n0 + nw)
  uint16_t raw_c0, raw_n;   // RAW / SIGMA / HEADS: first raw channel, number of raw channels from this block
  uint8_t kind, wait;       // wait: W_K* completions required before the block's first MMA (beyond the K-group waits)
  uint8_t save;             // HID: 1 = write the staging image, 2 = ... and it is the last block of the step: bulk store
  uint8_t pad;
};
struct Ts2Group {           // a run of 16-wide K-steps with one operand source
  uint8_t src, wait;        // GS_*, W_* completion required before its first MMA
  uint16_t col;             // GS_TMEM: first TMEM column (8 per K-step); GS_X / GS_D: byte offset inside the encoding image
  uint16_t ksteps, k0;      // number of K-steps, first K-step index inside the weight image
};
struct Ts2Step {
  Ts2Block blk[3];
  Ts2Group grp[3];
  uint8_t n_blk, n_grp;
  uint16_t bias_off;        // offset of the step's biases in the shared bias table (floats), indexed by weight row
  uint32_t w_bytes, w_rows; // weight image [K/8][w_rows][8] bf16
  const uint8_t* w_img;
  const float* bias;        // [w_rows]
  uint8_t* gdst;            // saved image (or null)
  uint32_t g_tile_stride, save_bytes;
};
// The MMA issuer's program: one record per RUN of K-steps (one operand source, one accumulator), built on the host and
// kept in SHARED memory.  The issuer is ONE small loop over these records.  Measured (NEFES_CHAIN_DBG stamps, round 2): a
// straight-line issuer specialised per (block, group, source) -- every MMA site executed once per step -- spent 64-76
// cycles per N=64 MMA and ~340 cycles between two blocks although the pipe was idle: each step walked ~50 cold
// instruction-cache lines (the SM's 19 warps run four different roles through a 56 KB kernel; L0 is ~6 KB).  The same
// MMAs issue at the 32-cycle floor from a loop that stays resident (tools/tc_issue.cu).
enum { RF_SRC = 3,            // GS_*
       RF_FRESH = 4,          // first run of a block: its first MMA overwrites the accumulator
       RF_FIRST = 8,          // first run of a step: turn-taking wait, weights wait
       RF_END = 16,           // last run of a step: release the weight slot
       RF_HID = 128,          // a whole 128-wide hidden layer (two 64-column blocks, both commits): nks = 0 H only, 1 X only, 2 X then H
       RF_COMMIT_SHIFT = 5,   // bits 5-6: 1 + block index whose accumulator barrier is committed behind this run (0: none)
       RF_WAIT_SHIFT = 8 };   // bits 8-15: W_* completions required before the run's first MMA
struct Ts2Run {
  uint32_t d_col, idesc, a, nks;        // accumulator column, instruction descriptor, TMEM column / encoding byte offset, K-steps
  uint32_t b16, lbo_field, dbk, flags;  // B offset in the ring slot (>>4), LBO field of the B descriptor, B advance per K-step (>>4)
};
static_assert(sizeof(Ts2Run) == 32, "Ts2Run is read with two 16-byte loads");
constexpr int kTs2MaxRuns = 56;
constexpr int kTs2MaxSteps = 14;
struct Ts2Args {
  Ts2Step step[kTs2MaxSteps];
  const Ts2Run* runs;         // [n_runs] in global memory, copied to shared memory at kernel start
  int n_runs;
  int n_steps;
  int64_t M; int n_tiles;
  float* raw; int C;
  const uint8_t* x_img; const uint8_t* d_img;     // encodings (bf16 images) written by encode_images_kernel
  int x_issue, d_issue;       // producer step at which the NEXT pair's encodings are requested (>= last reader + ring depth)
  int n_slots; uint32_t off_ring;   // depth and position of the weight ring (see ts2_off_ring)
  long long* dbg;
  int xflags;                 // timing experiments (NEFES_CHAIN_X): 1 no saves, 4 no raw stores
  int save_mode;              // 0: staging image + one bulk store per block; 1: st.global.v4 from the epilogue registers; 2 (default): as 0 with an L2 evict_first policy
};

// Shared memory: [ bias table | xyz encodings x2 | direction encodings x2 | staging images x2 (only when the launch saves) |
// weight ring ].  A weight image takes ~2400 cycles from L2 under load (measured), more than a step, and a slot is only
// released when BOTH tiles have retired the step -- so the ring is as deep as the space allows: 3 slots without the staging
// images (forward-only launches), 2 with them.
constexpr int kTs2MaxSlots = 3;
constexpr uint32_t kTs2WSlot = 49152;
constexpr uint32_t kTs2X = 16384, kTs2D = 8192, kTs2Stage = 32768;
constexpr uint32_t kTs2OffBias = 0;
constexpr uint32_t kTs2OffProg = 6656;                               // >= kChainBiasBytes
constexpr uint32_t kTs2ProgBytes = kTs2MaxRuns * sizeof(Ts2Run);
constexpr uint32_t kTs2OffX = 10240;                                 // >= kTs2OffProg + kTs2ProgBytes, 512-byte aligned
static_assert(kTs2OffProg + kTs2ProgBytes <= kTs2OffX, "issue program overlaps the encodings");
constexpr uint32_t kTs2OffD = kTs2OffX + 2 * kTs2X;
constexpr uint32_t kTs2OffStage = kTs2OffD + 2 * kTs2D;
static_assert(kTs2OffProg >= kChainBiasBytes, "bias table overlaps the issue program");
inline uint32_t ts2_off_ring(bool saves) { return kTs2OffStage + (saves ? 2 * kTs2Stage : 0u); }
inline int ts2_slots(bool saves) { return saves ? 2 : 3; }
inline uint32_t ts2_smem(bool saves) { return ts2_off_ring(saves) + ts2_slots(saves) * kTs2WSlot; }
constexpr uint32_t kTs2Acc = 0, kTs2HA = 128, kTs2HB = 192;
constexpr uint32_t kTs2AccCol = 0;

// Straight-line issue of a 128-wide hidden layer of one tile: block b (b = 0, 1) accumulates weight rows [64 b, 64 b + 64)
// into accumulator columns [64 b, 64 b + 64) over VAR = 0: the 8 K-steps of the activation columns `hin`; VAR = 1: the 4
// K-steps of the xyz encoding in shared memory (first layer); VAR = 2: both (skip layer, encoding first).  Weight images
// have 128 rows: a K-step advances the B descriptor by 2 * 2048 B, block 1 starts 64 rows = 1024 B in.  ~4 instructions per
// MMA; one commit per block.
template <int VAR>
__device__ __forceinline__ void ts2_issue_hidden(uint32_t tm, uint32_t hin, uint32_t blo, uint32_t xs, uint32_t idesc,
                                                 uint64_t* acc0, uint64_t* acc1) {
  constexpr int KX = (VAR != 0) ? 4 : 0, KH = (VAR != 1) ? 8 : 0;
  const uint64_t hi = (uint64_t)0x4008u << 32;                 // SBO = 128 B, descriptor version 1
  const uint64_t xd = smem_desc(xs, kChunkBytes, 128);
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const uint32_t d = tm + kTs2AccCol + 64u * b;
#pragma unroll
    for (int k = 0; k < KX; ++k)
      mma_ss(d, xd + (uint64_t)(k * (2 * kChunkBytes >> 4)), hi | (uint64_t)(blo + 64u * b + 256u * k), idesc, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < KH; ++k)
      mma_ts(d, tm + hin + 8u * k, hi | (uint64_t)(blo + 64u * b + 256u * (k + KX)), idesc, (k + KX) > 0 ? 1u : 0u);
    mma_commit(b == 0 ? acc0 : acc1);
  }
}

__global__ void __launch_bounds__(kChainThreads, 1) chain_fwd_ts2_kernel(const __grid_constant__ Ts2Args A) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_wfull[kTs2MaxSlots], bar_wempty[kTs2MaxSlots], bar_acc[2][3], bar_k[2][3], bar_x[2], bar_d[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sW = smem + A.off_ring;
  float* sBias = reinterpret_cast<float*>(smem + kTs2OffBias);
  const int n_slots = A.n_slots;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTs2MaxSlots; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 2); }
    for (int g = 0; g < 2; ++g) {
      for (int b = 0; b < 3; ++b) { mbar_init(&bar_acc[g][b], 1); mbar_init(&bar_k[g][b], kChainEpiWarps / 2); }
      mbar_init(&bar_x[g], 1); mbar_init(&bar_d[g], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  for (int s = 0; s < A.n_steps; ++s)
    if (A.step[s].bias != nullptr)
      for (int i = threadIdx.x; i < (int)A.step[s].w_rows; i += kChainThreads) sBias[A.step[s].bias_off + i] = A.step[s].bias[i];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A.runs);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + kTs2OffProg);
    for (int i = threadIdx.x; i < (int)(A.n_runs * sizeof(Ts2Run) / 4); i += kChainThreads) dst[i] = src[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_pairs = (A.n_tiles + 1) >> 1;
  const int n_steps = A.n_steps;

  if (warp == 0) {
    // ------------------------------- producer (converged warp, one elected lane issues) --------------------------------
    const bool leader = elect_one();
    const bool has_d = A.d_img != nullptr;
    auto load_enc = [&](int pair, bool x, bool d) {
      if (!leader) return;
      for (int g = 0; g < 2; ++g) {
        const int tile = pair * 2 + g;
        if (tile >= A.n_tiles) break;
        if (x) {
          mbar_arrive_expect_tx(&bar_x[g], kTs2X);
          bulk_g2s(smem + kTs2OffX + g * kTs2X, A.x_img + (int64_t)tile * kTs2X, kTs2X, &bar_x[g]);
        }
        if (d && has_d) {
          mbar_arrive_expect_tx(&bar_d[g], kTs2D);
          bulk_g2s(smem + kTs2OffD + g * kTs2D, A.d_img + (int64_t)tile * kTs2D, kTs2D, &bar_d[g]);
        }
      }
    };
    if ((int)blockIdx.x < n_pairs) load_enc(blockIdx.x, true, true);
    uint32_t cnt = 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int next = pair + (int)gridDim.x;
      for (int s = 0; s < n_steps; ++s, ++cnt) {
        const int slot = cnt % n_slots;
        mbar_wait(&bar_wempty[slot], ((cnt / n_slots) & 1) ^ 1);
        const uint32_t bytes = A.step[s].w_bytes;
        const uint8_t* src = A.step[s].w_img;
        if (leader && (A.xflags & 16) && cnt >= 2u * n_steps) {    // timing experiment: no weight traffic after two pairs
          mbar_arrive(&bar_wfull[slot]);
        } else if (leader) {
          mbar_arrive_expect_tx(&bar_wfull[slot], bytes);
          // every SM streams the SAME image at about the same time: pieces of 4 KB, each SM starting at a different one,
          // so that the requests of the 148 SMs do not queue up on the same L2 lines (NEFES_CHAIN_X & 64: two 16 KB pieces)
          if (A.xflags & 64) {
            for (uint32_t off = 0; off < bytes; off += 16384u)
              bulk_g2s(sW + slot * kTs2WSlot + off, src + off, min(16384u, bytes - off), &bar_wfull[slot]);
          } else {
            const uint32_t np = (bytes + 4095u) >> 12; // the last piece may be short
            uint32_t i = (blockIdx.x * 5u) % np;
            for (uint32_t n = 0; n < np; ++n) {
              const uint32_t off = i << 12;
              bulk_g2s(sW + slot * kTs2WSlot + off, src + off, min(4096u, bytes - off), &bar_wfull[slot]);
              i = (i + 1 == np) ? 0u : i + 1;
            }
          }
        }
        // The wait above proves that the MMAs of step s - n_slots of BOTH tiles retired (and, the pipe being in order, all
        // earlier ones): an encoding image whose last reader is that step may be overwritten.  A request that falls behind
        // the last step of the pair is made at the start of the next pair, for that pair itself.
        if (s == A.x_issue && next < n_pairs) load_enc(next, true, false);
        if (s == A.d_issue && next < n_pairs) load_enc(next, false, true);
        if (pair != (int)blockIdx.x) {
          if (s + n_steps == A.x_issue) load_enc(pair, true, false);
          if (s + n_steps == A.d_issue) load_enc(pair, false, true);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || warp == 2 + 2 * kChainEpiWarps) {
    // ------------------------------- MMA issuer of tile g -----------------------------------------------------------------
    // The issuer is a single warp executing a serial instruction stream at ~4 cycles per instruction (measured, ncu warp
    // sampling: 70 % of its samples are fixed-latency / branch stalls inside its own code, not barrier waits).  A 128-wide
    // layer is 512 cycles of tensor pipe, so the issuer may spend ~100 instructions on it: the hidden layers (most of the
    // MMAs) go through a straight-line template with compile-time operand offsets (ts2_issue_hidden), everything else
    // through one lean loop over run records.
    const int g = uni(warp == 1 ? 0 : 1);
    const uint32_t tm = uni(tmem) + g * 256;
    const bool leader = elect_one();
#ifdef NEFES_TS2_STAMPS
    const bool dbg = A.dbg != nullptr && blockIdx.x == 0;
#endif
    const uint32_t xs = smem_u32(smem + kTs2OffX + g * kTs2X), ds = smem_u32(smem + kTs2OffD + g * kTs2D);
    const uint4* runs = reinterpret_cast<const uint4*>(smem + kTs2OffProg);
    const int n_runs = A.n_runs;
    const uint32_t w0 = smem_u32(sW) >> 4;
    uint32_t cnt = 0, nk0 = 0u, nk1 = 0u, nk2 = 0u, n_tile = 0;
    uint32_t slot = 0, wphase = 0, wslot16 = w0;
    auto wait_latest = [&](uint64_t* bar, uint32_t n) { if (n > 0) mbar_wait(bar, (n - 1) & 1); };
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const bool valid = pair * 2 + g < A.n_tiles;
      if (valid) ++n_tile;
      for (int r = 0; r < n_runs; ++r) {
        const uint4 q0 = runs[2 * r], q1 = runs[2 * r + 1];
        const uint32_t flags = q1.w;
        const uint32_t m = (flags >> RF_WAIT_SHIFT) & 0xffu;
        if (valid && m != 0u) {
          if (m & W_K0) wait_latest(&bar_k[g][0], nk0);
          if (m & W_K1) wait_latest(&bar_k[g][1], nk1);
          if (m & W_K2) wait_latest(&bar_k[g][2], nk2);
          if (m & W_X) wait_latest(&bar_x[g], n_tile);
          if (m & W_D) wait_latest(&bar_d[g], n_tile);
        }
        if (flags & RF_FIRST) {
#ifdef NEFES_TS2_STAMPS
          if (dbg && leader && valid && cnt < 32) A.dbg[cnt * 48 + 46 + g] = clock64();     // operands ready
#endif
          mbar_wait(&bar_wfull[slot], wphase);
        }
        if (valid) {
          tc_fence_after();
#ifdef NEFES_TS2_STAMPS
          if (dbg && leader && cnt < 32 && (flags & RF_FIRST)) A.dbg[cnt * 48 + g * 2] = clock64();
#endif
          if (flags & RF_HID) {
            if (leader) {
              const uint32_t blo = (wslot16 + q1.x) | q1.y;
              if (q0.w == 0u) ts2_issue_hidden<0>(tm, q0.z, blo, xs, q0.y, &bar_acc[g][0], &bar_acc[g][1]);
              else if (q0.w == 1u) ts2_issue_hidden<1>(tm, q0.z, blo, xs, q0.y, &bar_acc[g][0], &bar_acc[g][1]);
              else ts2_issue_hidden<2>(tm, q0.z, blo, xs, q0.y, &bar_acc[g][0], &bar_acc[g][1]);
            }
            ++nk0; ++nk1;
          } else {
            const uint32_t commit = (flags >> RF_COMMIT_SHIFT) & 3u;
            if (leader) {
              const uint32_t d = tm + q0.x, id = q0.y;
              const uint64_t db0 = ((uint64_t)0x4008u << 32) | (uint64_t)((wslot16 + q1.x) | q1.y);
              const uint64_t dbk = (uint64_t)q1.z;
              const int nks = (int)q0.w;                  // always even
              const uint32_t keep = (flags & RF_FRESH) ? 0u : 1u;
              if ((flags & RF_SRC) == GS_TMEM) {
                const uint32_t a0 = tm + q0.z;
#pragma unroll 1
                for (int k = 0; k < nks; k += 2) {
                  mma_ts(d, a0 + k * 8, db0 + (uint64_t)k * dbk, id, k > 0 ? 1u : keep);
                  mma_ts(d, a0 + k * 8 + 8, db0 + (uint64_t)(k + 1) * dbk, id, 1u);
                }
              } else {
                const uint64_t da0 = smem_desc(((flags & RF_SRC) == GS_X ? xs : ds) + q0.z, kChunkBytes, 128);
#pragma unroll 1
                for (int k = 0; k < nks; k += 2) {
                  mma_ss(d, da0 + (uint64_t)(k * (2 * kChunkBytes >> 4)), db0 + (uint64_t)k * dbk, id, k > 0 ? 1u : keep);
                  mma_ss(d, da0 + (uint64_t)((k + 1) * (2 * kChunkBytes >> 4)), db0 + (uint64_t)(k + 1) * dbk, id, 1u);
                }
              }
              if (commit) mma_commit(&bar_acc[g][commit - 1]);
            }
            if (commit == 1u) ++nk0; else if (commit == 2u) ++nk1; else if (commit == 3u) ++nk2;   // the epilogue completes bar_k[g][b] once per block
          }
        }
        if (flags & RF_END) {
#ifdef NEFES_TS2_STAMPS
          if (dbg && leader && valid && cnt < 32) A.dbg[cnt * 48 + g * 2 + 1] = clock64();
#endif
          if (leader) mma_commit(&bar_wempty[slot]);
          ++cnt;
          wslot16 += kTs2WSlot >> 4;
          if (++slot == (uint32_t)n_slots) { slot = 0; wphase ^= 1u; wslot16 = w0; }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- epilogue warps -------------------------------------------------------------------------
    // Block-parallel: of a tile's 8 warps, warps 0-3 (one per TMEM lane quarter) serve the EVEN blocks of a step (block 0,
    // and a third block such as the sigma column), warps 4-7 the ODD block (block 1), each warp all columns of its block in
    // 32-column passes.  Block 1's epilogue -- which the next layer's MMAs wait for -- therefore never queues behind block
    // 0's epilogue and saved-copy work on the same warps (measured before: block 1 picked up ~1000 cycles after block 0's
    // hand-over with saves on).  Each block stores its own 16 KB half of the saved image (staging + bulk store behind a
    // 4-warp barrier).
    const int ew = warp - 2;
    const int g = ew >> 3;
    const int hgrp = (ew >> 2) & 1;                   // serves blocks b with (b & 1) == hgrp
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gt4 = (ew & 3) * 32 + lane;             // 0..127 inside the 4-warp group
    const uint32_t tbase = tmem + g * 256 + ((uint32_t)(q * 32) << 16);
    uint8_t* stage = smem + kTs2OffStage + g * kTs2Stage;
    uint32_t na[3] = {0u, 0u, 0u}, ecnt = 0;
    // ONE arrival per warp: per-thread arrivals on one barrier word serialise in the SM's barrier unit (32 cycles per warp
    // instruction).  Every lane has fenced its tensor-memory accesses (tcgen05.fence::before_thread_sync) before the warp meets.
    auto warp_arrive = [&](uint64_t* bar) { __syncwarp(); if (lane == 0) mbar_arrive(bar); };
    auto half_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g * 2 + hgrp) : "memory"); };
    bool store_pending = false;
    const uint64_t evict_first = l2_evict_first_policy();
    const bool dbg = A.dbg != nullptr && blockIdx.x == 0 && lane == 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int tile = pair * 2 + g;
      if (tile >= A.n_tiles) break;
      const bool ok = (int64_t)tile * kTile + row < A.M;
      float* rawt = A.raw + (int64_t)tile * A.C * kTile + row;
      for (int s = 0; s < n_steps; ++s, ++ecnt) {
        const int n_blk = A.step[s].n_blk;
        const float* bias = sBias + A.step[s].bias_off;
        uint8_t* gdst = (A.xflags & 1) ? nullptr : A.step[s].gdst;
        const uint32_t g_stride = A.step[s].g_tile_stride;
        for (int b = hgrp; b < n_blk; b += 2) {
          // the block's fields as locals: later asm statements clobber "memory" and would force re-loads
          const int kind = A.step[s].blk[b].kind, n0 = A.step[s].blk[b].n0, nw = A.step[s].blk[b].nw;
          const uint32_t acc = tbase + A.step[s].blk[b].acc_col, out_col = A.step[s].blk[b].out_col;
          const int raw_c0 = A.step[s].blk[b].raw_c0, raw_n = A.step[s].blk[b].raw_n, save = A.step[s].blk[b].save;
          ++na[b];
          mbar_wait(&bar_acc[g][b], (na[b] - 1) & 1);
          tc_fence_after();
          if (dbg && ecnt < 32 && b == 0) A.dbg[ecnt * 48 + 8 + ew] = clock64();
          if (dbg && ecnt < 32 && b == 1 && (ew & 3) == 0) A.dbg[ecnt * 48 + 42 + g] = clock64();
          if (kind == BK_HID_RELU || kind == BK_HID) {
            const bool saving = gdst != nullptr && save != 0;
            uint8_t* srow = stage + (n0 >> 3) * kChunkBytes + row * 16;
            uint8_t* grow = saving ? gdst + (int64_t)tile * g_stride + (n0 >> 3) * kChunkBytes + row * 16 : nullptr;
            if (saving && A.save_mode != 1 && store_pending) {      // the block's previous bulk store must have read its staging half
              if (gt4 == 0) bulk_wait_read<0>();
              half_barrier();
              store_pending = false;
            }
#pragma unroll 1
            for (int c = 0; c < nw; c += 32) {                       // 32-column passes over the block's columns
              float4 bv[8];
              uint32_t v[32], w[16];
              lds_bias32(bias + n0 + c, bv);
              tmem_ld32(acc + c, v);
              tmem_ld_wait();
              if (kind == BK_HID_RELU) pack32<true>(v, bv, w); else pack32<false>(v, bv, w);
              tmem_st16(tbase + out_col + (c >> 1), w);
              if (saving) {
                if (A.save_mode == 1) {
                  // straight from the registers: a warp writes 512 contiguous bytes per 8-channel chunk (rows 32 q .. 32 q + 31)
#pragma unroll
                  for (int j = 0; j < 4; ++j) stg128(grow + ((c >> 3) + j) * kChunkBytes, w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(srow + ((c >> 3) + j) * kChunkBytes) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                }
              }
            }
            tmem_st_wait();
            tc_fence_before();
            warp_arrive(&bar_k[g][b]);
            if (dbg && ecnt < 32 && b == 0) A.dbg[ecnt * 48 + 24 + ew] = clock64();
            if (dbg && ecnt < 32 && b == 1 && (ew & 3) == 0) A.dbg[ecnt * 48 + 44 + g] = clock64();
            if (saving && A.save_mode != 1) {
              // Saved copy: this block's nw channels are nw * 256 contiguous bytes of the image.  (Measured within 4 % of each
              // other in round 2: one 32 KB bulk store per step, st.global.v4 from the registers, 512-byte bulk stores per warp.
              // The L2 evict_first policy on the store is worth 5-6 % of the launch: 0.600 against 0.635 ms, fine query, same box.)
              fence_async_smem();
              half_barrier();
              if (gt4 == 0) {
                if (A.save_mode == 2) bulk_s2g_hint(gdst + (int64_t)tile * g_stride + (n0 >> 3) * kChunkBytes, stage + (n0 >> 3) * kChunkBytes, (uint32_t)nw * 256u, evict_first);
                else bulk_s2g(gdst + (int64_t)tile * g_stride + (n0 >> 3) * kChunkBytes, stage + (n0 >> 3) * kChunkBytes, (uint32_t)nw * 256u);
                bulk_commit();
              }
              store_pending = true;
            }
          } else if (kind == BK_RAW) {
            // raw channels [raw_c0, raw_c0 + raw_n) from accumulator columns [0, raw_n) of the block: 32-column passes, then the
            // tail beyond 64 (the three channels 128..130 of the colour head)
            uint32_t t4[4] = {0u, 0u, 0u, 0u};
            if (raw_n > 64)
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(t4[0]), "=r"(t4[1]), "=r"(t4[2]), "=r"(t4[3]) : "r"(acc + 64) : "memory");
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
              uint32_t v[32];
              tmem_ld32(acc + c, v);
              tmem_ld_wait();
              if (c == 32) {                                          // every column of the block is in registers: hand over
                tc_fence_before();
                warp_arrive(&bar_k[g][b]);
                if (dbg && ecnt < 32 && b == 0) A.dbg[ecnt * 48 + 24 + ew] = clock64();
              }
              if (ok && !(A.xflags & 4)) {
                const float* bp = bias + n0 + c;
                float* rp = rawt + (raw_c0 + c) * kTile;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bp + 4 * j);
                  stg32f(rp + (4 * j + 0) * kTile, __uint_as_float(v[4 * j + 0]) + b4.x);
                  stg32f(rp + (4 * j + 1) * kTile, __uint_as_float(v[4 * j + 1]) + b4.y);
                  stg32f(rp + (4 * j + 2) * kTile, __uint_as_float(v[4 * j + 2]) + b4.z);
                  stg32f(rp + (4 * j + 3) * kTile, __uint_as_float(v[4 * j + 3]) + b4.w);
                }
              }
            }
            if (ok && !(A.xflags & 4) && raw_n > 64) {
              for (int e = 0; e < raw_n - 64 && e < 4; ++e)
                stg32f(rawt + (raw_c0 + 64 + e) * kTile, __uint_as_float(t4[e]) + bias[n0 + 64 + e]);
            }
          } else {                                     // BK_SIGMA / BK_HEADS: a few activated columns
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(acc) : "memory");
            tmem_ld_wait();
            tc_fence_before();
            warp_arrive(&bar_k[g][b]);
            if (dbg && ecnt < 32 && b == 0) A.dbg[ecnt * 48 + 24 + ew] = clock64();
            if (ok) {
              if (kind == BK_SIGMA) {
                stg32f(rawt + raw_c0 * kTile, softplus_f(__uint_as_float(v[0]) + bias[n0]));
              } else {
#pragma unroll
                for (int e = 0; e < 5; ++e) {
                  const float x = __uint_as_float(v[e]) + bias[n0 + e];
                  stg32f(rawt + (raw_c0 + e) * kTile, e < 3 ? sigmoid_f(x) : softplus_f(x));
                }
              }
            }
          }
        }
      }
    }
    if (gt4 == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace nefes
