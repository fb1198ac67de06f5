// Fused layer chains of the NeFeS field (bf16 tensor path): one persistent kernel runs ALL layers of the MLP for a
// pair of 128-point tiles; activations (forward) / data gradients (backward) never leave the SM between layers.
//
//   warp 0      producer: streams each step's weight image L2 -> shared (bulk copies, 2-slot ring) and prefetches the
//               per-tile operand images (forward: xyz / direction encodings; backward: head-gradient images) for the
//               NEXT tile pair as soon as their slot of the tile region is dead
//   warp 1, 18  MMA issuers, one per tile of the pair (tcgen05.mma M=128, fp32 accumulators in TMEM, 256 columns per
//               tile).  One thread per tile because every barrier wait / commit costs the issuing thread 200-300 cycles:
//               with a single issuer those latencies of the two tiles add up and the tensor pipe idles half the time;
//               with two, the pipe works on one tile while the other tile's issuer waits and its epilogue runs
//   warps 2-9   epilogue of tile 0, warps 10-17 epilogue of tile 1: a warp owns 32 points (its TMEM lane quarter) x
//               one half of the output columns, so every scheduler has four epilogue warps to hide tcgen05.ld / LDS
//               latency behind each other.
//     forward:  TMEM -> registers -> bias (packed fp32x2 add) -> ReLU fused into the bf16 pack -> operand image of
//               the NEXT layer in shared memory (in place) and the same 16-byte words to HBM (saved for backward);
//               the fp32 outputs go straight to the tile-major raw block [C][128] (128-byte coalesced stores).
//     backward: TMEM -> registers -> bf16 pack -> ReLU mask taken from the SAVED activation (a > 0, one compare per
//               bf16 pair) -> gradient image of the next (earlier) layer in shared memory and to HBM (operand of the
//               weight-gradient kernel).
//
// The chain is table-driven (ChainStep): operand offset inside the tile region, weight image, epilogue kind,
// destination.  Included by mlp_tc.cu (uses its helpers).   script/models/nerfh_nff.py:525-576.
#pragma once

namespace nefes {

enum { CK_HIDDEN = 0,   // fwd: bias + ReLU -> image
       CK_FS = 1,       // fwd: bias -> image (128 ch: xyz_encoding_final); column 128: softplus -> raw[131]
       CK_HEADS = 2,    // fwd: columns 0..4: sigmoid x3, softplus x2 -> raw[132..136]
       CK_SIGMA = 3,    // fwd: column 0: softplus -> raw[0]                          (sigma-only mode)
       CK_RGB = 4,      // fwd: bias -> 131 fp32 columns -> raw[0..130]
       CK_DGRAD = 5,    // bwd: optional ReLU mask from the saved activation -> image
       CK_NONE = 6 };   // first K-slice of a layer whose weights do not fit one ring slot: the next step accumulates on top

constexpr int kChainLoads = 3;
struct ChainLoad {                 // one prefetch of a per-tile operand image into the tile region
  const uint8_t* src; uint32_t tile_stride, bytes, dst_off;
  int8_t issue_step;               // producer issues it when it reaches this step ...
  int8_t next_pair;                // ... for the NEXT pair (1) or the current one (0)
  int8_t pad[2];
};
struct ChainStep {
  uint32_t a_off;                  // A operand start inside the tile region (bytes)
  uint32_t out_off;                // image destination inside the tile region (bytes)
  uint16_t K, N, out_ch;
  uint8_t kind;
  int8_t wait_load;                // index of the ChainLoad the MMA of this step must wait for (-1: none)
  int8_t wait_load2;
  int8_t acc0;                     // 1: accumulate onto the previous step's partial sums (second K-slice)
  uint16_t bias_off, pad16;        // fwd: offset of this step's biases in the shared bias table (floats)
  uint32_t w_bytes, w_lbo;         // weight image: bytes to stream (compacted), LBO = bytes between 8-wide K chunks
  uint32_t w_piece, w_src_stride;  // streamed as pieces of w_piece bytes, w_src_stride apart in the source image
  const uint8_t* w_img;
  const float* bias;               // fwd: [N]
  uint8_t* gdst;                   // image saved to HBM (or null)
  const uint8_t* act;              // bwd: saved activation image whose sign gates this gradient (or null)
  uint32_t g_tile_stride, act_tile_stride;
};
constexpr int kChainMaxSteps = 16;
struct ChainArgs {
  ChainStep step[kChainMaxSteps];
  ChainLoad load[kChainLoads];
  int n_steps, n_loads;
  int64_t M; int n_tiles;
  float* raw; int C;                         // fwd: tile-major output blocks [n_tiles][C][128] fp32
  long long* dbg;                            // optional clock stamps of CTA 0 (NEFES_CHAIN_DBG)
  int xflags;                                // timing experiments only (NEFES_CHAIN_X): 1 no saves, 2 no PE, 4 no raw store
};

// forward tile region: [ xyzPE 16 KB | H 32 KB | dirPE 8 KB ] -- contiguous so the skip input [xyzPE | h4] (K=192)
// and the direction input [final | dirPE] (K=160) are plain operand ranges.
constexpr uint32_t kRegX = 0, kRegH = 16384, kRegD = 49152;
// backward tile region: P (36 KB) | Q (36 KB) | S (4 KB)
constexpr uint32_t kRegP = 0, kRegQ = 36864, kRegS = 73728;
constexpr uint32_t kFwdRegBytes = 57344, kBwdRegBytes = 77824;
constexpr uint32_t kFwdWSlot = 49152;        // largest forward W image: 192 x 128 bf16 (a layer whose image is larger is split in two K-slices)
constexpr uint32_t kBwdWSlot = 36864;        // largest backward WT image: 144 x 128 bf16
constexpr int kFwdSlots = 2, kBwdSlots = 2;  // depth of the weight ring (measured: a third slot + split layers buys nothing)
constexpr int kChainEpiWarps = 8;             // per tile
constexpr int kChainThreads = 64 + 2 * kChainEpiWarps * 32 + 32;   // + a second MMA-issuer warp (tile 1)
constexpr uint32_t kChainBiasBytes = 6400;   // shared bias table: sum of the layer widths of one chain (<= 1600 floats)
constexpr uint32_t kFwdChainSmem = 2 * kFwdRegBytes + kFwdSlots * kFwdWSlot + kChainBiasBytes;
constexpr uint32_t kBwdChainSmem = 2 * kBwdRegBytes + kBwdSlots * kBwdWSlot;

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

// (lo + blo, hi + bhi) as one packed fp32x2 add, then one cvt to a bf16 pair (ReLU folded into the cvt)
template <bool RELU>
__device__ __forceinline__ uint32_t bias_pack(uint32_t lo, uint32_t hi, float blo, float bhi) {
  uint32_t r;
  if (RELU)
    asm("{\n\t.reg .b64 a, b, c;\n\t.reg .f32 x, y;\n\t"
        "mov.b64 a, {%1, %2};\n\tmov.b64 b, {%3, %4};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {x, y}, c;\n\t"
        "cvt.rn.relu.bf16x2.f32 %0, y, x;\n\t}"
        : "=r"(r) : "r"(lo), "r"(hi), "f"(blo), "f"(bhi));
  else
    asm("{\n\t.reg .b64 a, b, c;\n\t.reg .f32 x, y;\n\t"
        "mov.b64 a, {%1, %2};\n\tmov.b64 b, {%3, %4};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {x, y}, c;\n\t"
        "cvt.rn.bf16x2.f32 %0, y, x;\n\t}"
        : "=r"(r) : "r"(lo), "r"(hi), "f"(blo), "f"(bhi));
  return r;
}

// 32 accumulator columns [c0, c0+32) of one row -> four 8-channel chunks of the image (and of its HBM copy).
// All eight bias vectors are loaded BEFORE the first image store: the compiler will not move a shared-memory load above
// a shared-memory store (possible alias), and a load -> add -> pack -> store chain per chunk exposes the LDS latency
// four times per call with only two epilogue warps per scheduler to cover it.
template <bool RELU>
__device__ __forceinline__ void fwd_cols32(const uint32_t (&v)[32], const float* __restrict__ bias, uint8_t* __restrict__ dst_row,
                                           uint8_t* __restrict__ gdst_row, int c0) {
  float4 b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(bias + c0 + 4 * j);
  uint4 pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pk[j].x = bias_pack<RELU>(v[8 * j + 0], v[8 * j + 1], b[2 * j].x, b[2 * j].y);
    pk[j].y = bias_pack<RELU>(v[8 * j + 2], v[8 * j + 3], b[2 * j].z, b[2 * j].w);
    pk[j].z = bias_pack<RELU>(v[8 * j + 4], v[8 * j + 5], b[2 * j + 1].x, b[2 * j + 1].y);
    pk[j].w = bias_pack<RELU>(v[8 * j + 6], v[8 * j + 7], b[2 * j + 1].z, b[2 * j + 1].w);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int off = ((c0 >> 3) + j) * (int)kChunkBytes;
    *reinterpret_cast<uint4*>(dst_row + off) = pk[j];
    if (gdst_row != nullptr) *reinterpret_cast<uint4*>(gdst_row + off) = pk[j];   // a warp's 32 rows x 16 B = 512 contiguous bytes
  }
}

__device__ __forceinline__ uint32_t pack_mask(uint32_t lo, uint32_t hi, uint32_t act, bool gate) {
  uint32_t p = pack_bf16(__uint_as_float(lo), __uint_as_float(hi));
  if (gate) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&act);
    p &= __hgt2_mask(a, __floats2bfloat162_rn(0.f, 0.f));
  }
  return p;
}

template <bool BWD>
__global__ void __launch_bounds__(kChainThreads, 1) chain_kernel(const ChainArgs A) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_wfull[3], bar_wempty[3], bar_act[2], bar_acc[2], bar_ld[2][kChainLoads], bar_stag;
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t kReg = BWD ? kBwdRegBytes : kFwdRegBytes;
  constexpr uint32_t kWSlot = BWD ? kBwdWSlot : kFwdWSlot;
  constexpr int kSlots = BWD ? kBwdSlots : kFwdSlots;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sW = smem + 2 * kReg;
  float* sBias = reinterpret_cast<float*>(smem + 2 * kReg + kSlots * kWSlot);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 2); }
    mbar_init(&bar_stag, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_act[i], kChainEpiWarps * 32); mbar_init(&bar_acc[i], 1);
      for (int l = 0; l < kChainLoads; ++l) mbar_init(&bar_ld[i][l], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  if (!BWD) {
    for (int s = 0; s < A.n_steps; ++s)
      if (A.step[s].bias != nullptr)
        for (int i = threadIdx.x; i < A.step[s].N; i += kChainThreads) sBias[A.step[s].bias_off + i] = A.step[s].bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_pairs = (A.n_tiles + 1) >> 1;
  const bool dbg = A.dbg != nullptr && blockIdx.x == 0;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- producer ---------------------------------------------------------------
      uint32_t cnt = 0;                 // W ring position == steps started so far
      auto issue_load = [&](int li, int pair) {
        const ChainLoad& L = A.load[li];
        for (int g = 0; g < 2; ++g) {
          const int tile = pair * 2 + g;
          if (tile >= A.n_tiles) break;
          mbar_arrive_expect_tx(&bar_ld[g][li], L.bytes);
          bulk_g2s(smem + g * kReg + L.dst_off, L.src + (int64_t)tile * L.tile_stride, L.bytes, &bar_ld[g][li]);
        }
      };
      // loads whose issue point is step s: their destination was last read by the MMAs of step s-1 (both tiles).  That is
      // exactly what releases the weight slot of step s-1, and the producer follows those phases one by one (it never
      // lags a phase behind, so the parity wait cannot alias -- unlike a wait on the per-tile accumulator barriers, which
      // may be several phases ahead of the producer).
      auto do_loads = [&](int s, int pair) {
        for (int l = 0; l < A.n_loads; ++l) {
          const ChainLoad& L = A.load[l];
          if (L.issue_step != s) continue;
          const int tp = L.next_pair ? pair + (int)gridDim.x : pair;
          if (tp >= n_pairs) continue;
          if (cnt > 0) mbar_wait(&bar_wempty[(cnt - 1) % kSlots], ((cnt - 1) / kSlots) & 1);
          issue_load(l, tp);
        }
      };
      for (int l = 0; l < A.n_loads; ++l)
        if (A.load[l].issue_step >= 0 && A.load[l].next_pair && (int)blockIdx.x < n_pairs) issue_load(l, blockIdx.x);
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int s = 0; s < A.n_steps; ++s, ++cnt) {
          const int slot = cnt % kSlots;
          mbar_wait(&bar_wempty[slot], ((cnt / kSlots) & 1) ^ 1);
          const uint32_t bytes = A.step[s].w_bytes;
          if ((A.xflags & 16) && cnt >= 2u * A.n_steps) {   // timing experiment: no weight traffic after the first pairs
            mbar_arrive(&bar_wfull[slot]);
          } else {
            mbar_arrive_expect_tx(&bar_wfull[slot], bytes);
            const uint32_t piece = A.step[s].w_piece, sstride = A.step[s].w_src_stride;
            const uint8_t* src = A.step[s].w_img;
            for (uint32_t off = 0; off < bytes; off += piece, src += sstride)
              bulk_g2s(sW + slot * kWSlot + off, src, min(piece, bytes - off), &bar_wfull[slot]);
          }
          // slots that only die with the last step of the previous pair: after this pair's first weights are on the way
          if (s == 0 && pair != (int)blockIdx.x) do_loads(A.n_steps, pair - (int)gridDim.x);
          do_loads(s, pair);
        }
      }
    }
  } else if (warp == 1 || warp == 2 + 2 * kChainEpiWarps) {
    if (lane == 0) {
      // ------------------------------- MMA issuer of tile g --------------------------------------------------------
      const int g = warp == 1 ? 0 : 1;
      uint32_t cnt = 0, act_ph = 0u, ld_ph[kChainLoads] = {};
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const bool valid = pair * 2 + g < A.n_tiles;
        for (int s = 0; s < A.n_steps; ++s, ++cnt) {
          const ChainStep& st = A.step[s];
          const int slot = cnt % kSlots;
          mbar_wait(&bar_wfull[slot], (cnt / kSlots) & 1);
          if (dbg && g == 0 && cnt < 32) A.dbg[cnt * 48] = clock64();
          if (valid) {
            const uint32_t idesc = idesc_bf16(128, st.N, 0, 0);
            const uint64_t db0 = smem_desc(smem_u32(sW + slot * kWSlot), st.w_lbo, 128);
            mbar_wait(&bar_act[g], act_ph);
            act_ph ^= 1u;
            if (st.wait_load >= 0) {
              mbar_wait(&bar_ld[g][st.wait_load], ld_ph[st.wait_load]);
              ld_ph[st.wait_load] ^= 1u;
            }
            if (st.wait_load2 >= 0) {
              mbar_wait(&bar_ld[g][st.wait_load2], ld_ph[st.wait_load2]);
              ld_ph[st.wait_load2] ^= 1u;
            }
            if (dbg && cnt < 32) A.dbg[cnt * 48 + 40 + g] = clock64();     // barriers passed, before the tcgen05 fence
            tc_fence_after();
            if (dbg && cnt < 32) A.dbg[cnt * 48 + 1 + g] = clock64();
            const uint64_t da0 = smem_desc(smem_u32(smem + g * kReg + st.a_off), kChunkBytes, 128);
            const uint32_t d = tmem + g * 256;
            for (int k = 0; k < st.K / 16; ++k)
              mma_ss(d, da0 + (uint64_t)(k * (2 * kChunkBytes >> 4)), db0 + (uint64_t)(k * (2 * st.w_lbo >> 4)), idesc,
                     (k > 0 || st.acc0) ? 1u : 0u);
            mma_commit(&bar_acc[g]);
            // anti-phase the two tiles: tile 1 starts a pair only when tile 0's first layer has retired, so that from then
            // on one tile is in its epilogue while the other owns the tensor pipe (left alone, the two tiles fall into
            // lock-step -- both in MMA, then both in epilogue -- and every step costs MMA + epilogue instead of the larger)
            if (g == 0 && s == 0) mma_commit(&bar_stag);
            if (dbg && cnt < 32) A.dbg[cnt * 48 + 3 + g] = clock64();
          }
          mma_commit(&bar_wempty[slot]);       // the slot is free once BOTH tiles' MMAs retired (count 2)
          if (dbg && cnt < 32) A.dbg[cnt * 48 + 5 + g] = clock64();
        }
      }
    }
  } else {
    // --------------------------------- epilogue warps ----------------------------------------------------------------
    const int ew = warp - 2;                          // 0..15
    const int g = ew >> 3;                            // tile of the pair this warp serves
    const int half = (ew >> 2) & 1;                   // which half of the output columns
    const int q = warp & 3;                           // TMEM lane quarter
    const int row = q * 32 + lane;
    const int gt = (ew & 7) * 32 + lane;              // 0..255 inside the group
    uint8_t* reg = smem + g * kReg;
    const uint32_t taddr = tmem + g * 256 + ((uint32_t)(q * 32) << 16);
    uint32_t acc_ph = 0u;
    uint32_t ecnt = 0, pair_it = 0;
    bool store_pending = false;                       // a bulk store of this group may still be reading the tile region
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pair_it) {
      const int tile = pair * 2 + g;
      if (tile >= A.n_tiles) break;
      if (g == 1) mbar_wait(&bar_stag, pair_it & 1);   // start half a period after tile 0 (see the MMA issuer)
      const int64_t grow = (int64_t)tile * kTile + row;
      const bool ok = grow < A.M;
      float* rawt = BWD ? nullptr : A.raw + (int64_t)tile * A.C * kTile + row;   // + c * 128
      mbar_arrive(&bar_act[g]);                        // step 0: its operand arrives by bulk copy

      for (int s = 0; s < A.n_steps; ++s, ++ecnt) {
        const ChainStep& st = A.step[s];
        const float* bias = sBias + st.bias_off;
        const int ncol = st.out_ch >> 1;               // image columns of this warp: [c_base, c_base + ncol), 32 or 64
        const int c_base = half * ncol;
        uint4 av[8];
        if (BWD) {                                     // saved activation of this row: issue the loads before waiting
          if (st.act != nullptr) {
            const uint8_t* ap = st.act + (int64_t)tile * st.act_tile_stride + (c_base >> 3) * kChunkBytes + row * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j * 8 < ncol) av[j] = __ldg(reinterpret_cast<const uint4*>(ap + j * kChunkBytes));
          }
        }
        mbar_wait(&bar_acc[g], acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
        if (dbg && lane == 0 && ecnt < 32) A.dbg[ecnt * 48 + 8 + ew] = clock64();
        if (!BWD && store_pending) {                   // forward saves leave by bulk store: it must have read its image
          if (gt == 0) bulk_wait_read<0>();            // before anything below may overwrite the region
          group_barrier(g);
          store_pending = false;
        }
        if (!BWD && st.kind == CK_NONE) {
          // partial sums stay in TMEM; nothing to do
        } else if (BWD || st.kind == CK_HIDDEN || st.kind == CK_FS) {
          // the image is written in place: its last reader (the MMA that just completed) is done
          uint8_t* dst_row = reg + st.out_off + row * 16;
          uint8_t* gdst_row = ((BWD || (A.xflags & 8)) && st.gdst && !(A.xflags & 1)) ? st.gdst + (int64_t)tile * st.g_tile_stride + row * 16 : nullptr;
          if (BWD) {
            const bool gate = st.act != nullptr;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              if (h2 * 32 < ncol) {
                uint32_t v[32];
                tmem_ld32(taddr + c_base + h2 * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 a = av[h2 * 4 + j];
                  uint4 pk;
                  pk.x = pack_mask(v[8 * j + 0], v[8 * j + 1], a.x, gate);
                  pk.y = pack_mask(v[8 * j + 2], v[8 * j + 3], a.y, gate);
                  pk.z = pack_mask(v[8 * j + 4], v[8 * j + 5], a.z, gate);
                  pk.w = pack_mask(v[8 * j + 6], v[8 * j + 7], a.w, gate);
                  const int off = ((c_base >> 3) + h2 * 4 + j) * (int)kChunkBytes;
                  *reinterpret_cast<uint4*>(dst_row + off) = pk;
                  if (gdst_row != nullptr) *reinterpret_cast<uint4*>(gdst_row + off) = pk;
                }
              }
            }
          } else {
            uint32_t v0[32], v1[32];
            tmem_ld32(taddr + c_base, v0);
            if (ncol == 64) tmem_ld32(taddr + c_base + 32, v1);
            tmem_ld_wait();
            if (dbg && lane == 0 && ew == 0 && ecnt < 32) A.dbg[1600 + ecnt * 4 + 0] = clock64();
            if (st.kind == CK_HIDDEN) {
            fwd_cols32<true>(v0, bias, dst_row, gdst_row, c_base);
            if (ncol == 64) fwd_cols32<true>(v1, bias, dst_row, gdst_row, c_base + 32);
          } else {
            fwd_cols32<false>(v0, bias, dst_row, gdst_row, c_base);
            if (ncol == 64) fwd_cols32<false>(v1, bias, dst_row, gdst_row, c_base + 32);
            if (half == 1) {                           // sigma pre-activation rides in column 128 of this GEMM
              uint32_t v[16];
              tmem_ld16(taddr + 128, v);
              tmem_ld_wait();
              if (ok) rawt[131 * kTile] = softplus_f(__uint_as_float(v[0]) + bias[128]);
            }
          }
          }
        } else if (st.kind == CK_HEADS || st.kind == CK_SIGMA) {
          if (half == 0) {
            uint32_t v[16];
            tmem_ld16(taddr, v);
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
        } else if (!(A.xflags & 4)) {   // CK_RGB: 131 fp32 columns; half 0 -> [0, 80), half 1 -> [80, 131)
          const int cb = half * 80;
          const int nblk = half == 0 ? 5 : 4;          // 16-column blocks; the last one of half 1 is cut at 131
#pragma unroll 1
          for (int b2 = 0; b2 < nblk; ++b2) {
            uint32_t v[16];
            tmem_ld16(taddr + cb + b2 * 16, v);
            tmem_ld_wait();
            const int c0 = cb + b2 * 16;
            if (ok) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (c0 + e < kHeadCh) rawt[(c0 + e) * kTile] = __uint_as_float(v[e]) + bias[c0 + e];
            }
          }
        }
        if (dbg && lane == 0 && ew == 0 && ecnt < 32) A.dbg[1600 + ecnt * 4 + 1] = clock64();
        tc_fence_before();
        fence_async_smem();
        if (dbg && lane == 0 && ew == 0 && ecnt < 32) A.dbg[1600 + ecnt * 4 + 2] = clock64();
        if (!BWD && st.gdst != nullptr && !(A.xflags & 9)) {   // saved for backward: the image just written, one bulk store
          group_barrier(g);
          if (dbg && lane == 0 && ew == 0 && ecnt < 32) A.dbg[1600 + ecnt * 4 + 3] = clock64();
          if (gt == 0) {
            bulk_s2g(st.gdst + (int64_t)tile * st.g_tile_stride, reg + st.out_off, (uint32_t)st.out_ch * 256u);
            bulk_commit();
          }
          store_pending = true;
        }
        if (dbg && lane == 0 && ecnt < 32) A.dbg[ecnt * 48 + 24 + ew] = clock64();
        if (s + 1 < A.n_steps) mbar_arrive(&bar_act[g]);   // operand of the next step is ready, accumulator drained
      }
    }
    if (!BWD && gt == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace nefes
