// Fused backward of TWO consecutive trunk layers of the NeFeS field (training): data gradient AND weight gradient
// of both layers in one pass over the tiles, with the weight gradients accumulated in TMEM for the whole launch.
//
//   per 128-point tile, layer j = 0, 1 (layer L, then layer L-1):
//     d_j   acc[128 pts, 128 in]   = G_j[pts, out] * W_j              tcgen05.mma, K-major operands
//     w_j   dW_j[128 out, N_j in] += G_j^T[out, pts] * A_j[pts, in]   the SAME G image and the saved activation image,
//     b_j   db_j[128 out, 16]     += G_j^T[out, pts] * ones[pts, 16]  both read MN-major (K = points): no transposes
//     e_j   acc -> ReLU mask from A_j (a > 0) -> bf16 -> G_{j+1} image in shared memory (operand of the next layer)
//   G_0 arrives from HBM (bulk copy, prefetched one tile ahead), G_2 leaves to HBM (bulk store) as the next launch's
//   G_0; the activation images A_j are read from HBM exactly once and serve the mask and the weight gradient.  Both
//   weight matrices stay resident in shared memory.  After the last tile the accumulators are flushed with vector
//   reductions (red.global.add.v4.f32) into the flat gradient buffer.
//
//   warp 0: producer (bulk copies / bulk store), warp 1: MMA issuer, warps 2-9: epilogue (thread = point x column half).
//
// HBM per tile and launch: 32 KB G in + 2 x 32 KB activations + 32 KB G out = 128 KB for two layers, against
// 2 x (32 + 32 + 32 + 32) KB = 256 KB for the separate data-gradient chain + weight-gradient kernel it replaces.
// Included by mlp_tc.cu.   script/models/nerfh_nff.py:469-476, :546-553 (trunk), backward.
#pragma once

namespace nefes {

struct TrunkStep {
  const uint8_t* wt_img;           // WT image of the layer [out/8][rows][8]; rows [wt_row0, +128) are used (dgrad B operand)
  uint32_t wt_rows, wt_row0;
  int has_dgrad;                   // 0: weight gradient only (first trunk layer)
  const uint8_t* act; uint32_t act_tile_stride; int act_ch;   // saved input activation image (64 or 128 channels)
  uint8_t* g_save; uint32_t g_save_tile_stride;               // optional HBM copy of this step's OUTPUT gradient image
  int pl, k_off;                   // scatter map of the weight gradient: packed layer, first input channel
  int bias;                        // 1: this launch owns the bias gradient of the layer
};
struct TrunkArgs {
  TrunkStep step[2];
  const uint8_t* g_in; uint32_t g_in_tile_stride;   // gradient image entering the pair of layers
  uint8_t* g_out; uint32_t g_out_tile_stride;       // gradient image leaving it (null: none)
  int n_tiles;
  int bulk_flush;                                   // 1: contiguous weight-gradient blocks leave as bulk reductions
  PackSrc ps; float* d_flat;
};

constexpr uint32_t kTrOffWT = 0;                    // 2 x 32 KB weights
constexpr uint32_t kTrOffG = 65536;                 // 3 x 32 KB gradient images: in (even tiles), in (odd tiles), mid
constexpr uint32_t kTrOffA = kTrOffG + 3 * 32768;   // 2 x 32 KB activation images
constexpr uint32_t kTrOffOnes = kTrOffA + 2 * 32768;
constexpr uint32_t kTrunkSmem = kTrOffOnes + 2048;
constexpr int kTrunkThreads = 64 + 256 + 32;     // producer, dgrad issuer, 8 epilogue warps, wgrad issuer
constexpr uint32_t kTrAcc = 0, kTrDW0 = 128, kTrDB0 = 256, kTrDW1 = 272, kTrDB1 = 400;   // TMEM columns

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One launch walks up to four pairs of layers ("passes"): the gradient image a pass writes for a tile is read back by the
// SAME CTA in the next pass (tile -> CTA assignment is fixed), so no grid-wide barrier separates the passes -- only a CTA-local
// one at which the weight-gradient accumulators are flushed, the mbarriers re-initialised and the next pair of weight
// matrices loaded.  Against four launches this saves three launch gaps, TMEM allocations and pipeline drains that wait for
// the slowest CTA of the grid.
struct TrunkMulti { TrunkArgs pass[4]; int n_pass; };

__global__ void __launch_bounds__(kTrunkThreads, 1) trunk_bwd_kernel(const __grid_constant__ TrunkMulti P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_gin[2], bar_afull[2], bar_afree[2], bar_acc, bar_gmid, bar_gout, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto init_bars = [&](bool again) {
    if (again) {
      mbar_inval(&bar_w); mbar_inval(&bar_acc); mbar_inval(&bar_gmid); mbar_inval(&bar_gout); mbar_inval(&bar_done);
      for (int i = 0; i < 2; ++i) { mbar_inval(&bar_gin[i]); mbar_inval(&bar_afull[i]); mbar_inval(&bar_afree[i]); }
    }
    mbar_init(&bar_w, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_gin[i], 1); mbar_init(&bar_afull[i], 1); mbar_init(&bar_afree[i], 257); }
    mbar_init(&bar_acc, 1); mbar_init(&bar_gmid, 256); mbar_init(&bar_gout, 256); mbar_init(&bar_done, 2);
    fence_mbar_init();
  };
  if (threadIdx.x == 0) init_bars(false);
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  {  // sixteen "ones" channels read MN-major: 8 channels x 128 points, both 8-channel groups alias the same 2 KB (SBO = 0)
    uint4 ones;
    ones.x = ones.y = ones.z = ones.w = 0x3F803F80u;
    for (int i = threadIdx.x; i < 2048 / 16; i += kTrunkThreads) reinterpret_cast<uint4*>(smem + kTrOffOnes)[i] = ones;
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  int n_my = 0;
  for (int t = blockIdx.x; t < P.pass[0].n_tiles; t += gridDim.x) ++n_my;

  for (int pass = 0; pass < P.n_pass; ++pass) {
  const TrunkArgs& T = P.pass[pass];
  const bool has_out = T.g_out != nullptr;          // step 1 has a data gradient (and an epilogue)
  const bool two = T.step[1].act != nullptr;        // the pass covers two layers
  if (pass > 0) {                                   // every role finished the previous pass: accumulators flushed, stores complete
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) init_bars(true);
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) {
    if (lane == 0 && n_my > 0) {
      // ------------------------------- producer ---------------------------------------------------------------------
      uint32_t wbytes = 0;
      for (int j = 0; j < 2; ++j) if (T.step[j].has_dgrad) wbytes += 32768u;
      if (wbytes) {
        mbar_arrive_expect_tx(&bar_w, wbytes);
        for (int j = 0; j < 2; ++j) {
          if (!T.step[j].has_dgrad) continue;
          const TrunkStep& s = T.step[j];
          for (int c = 0; c < 16; ++c)      // rows [row0, row0+128) of every 8-wide out-channel chunk, compacted
            bulk_g2s(smem + kTrOffWT + j * 32768 + c * 2048, s.wt_img + ((int64_t)c * s.wt_rows + s.wt_row0) * 16, 2048u, &bar_w);
        }
      }
      auto load_g = [&](int it) {
        const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
        mbar_arrive_expect_tx(&bar_gin[it & 1], 32768u);
        bulk_g2s(smem + kTrOffG + (it & 1) * 32768, T.g_in + tile * T.g_in_tile_stride, 32768u, &bar_gin[it & 1]);
      };
      auto load_a = [&](int j, int it) {
        const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
        const uint32_t bytes = (uint32_t)T.step[j].act_ch * 256u;
        mbar_arrive_expect_tx(&bar_afull[j], bytes);
        bulk_g2s(smem + kTrOffA + j * 32768, T.step[j].act + tile * T.step[j].act_tile_stride, bytes, &bar_afull[j]);
      };
      load_g(0);
      load_a(0, 0);
      if (two) load_a(1, 0);
      if (n_my > 1 && !has_out) load_g(1);          // with an output, buffer 1 is claimed below in program order
      for (int it = 0; it < n_my; ++it) {
        // (1) G buffer (it+1)&1: holds G_out of tile it-1 (being stored) -- or, without an output, G_in of tile it-1
        if (has_out) {
          if (it >= 1) {
            mbar_wait(&bar_gout, (it - 1) & 1);                     // epilogue 1 of tile it-1 wrote G_out
            const int64_t tile = blockIdx.x + (int64_t)(it - 1) * gridDim.x;
            bulk_s2g(T.g_out + tile * T.g_out_tile_stride, smem + kTrOffG + ((it - 1) & 1) * 32768, 32768u);
            bulk_commit();
            bulk_wait_read<0>();
          }
          if (it + 1 < n_my) load_g(it + 1);
        } else if (it >= 1 && it + 1 < n_my) {
          load_g(it + 1);       // G_in of tile it-1 is dead: step (2) of the previous iteration waited for its last readers
        }
        // (2) activation slots for tile it+1
        if (it + 1 < n_my) {
          mbar_wait(&bar_afree[0], it & 1);
          load_a(0, it + 1);
          if (two) {
            mbar_wait(&bar_afree[1], it & 1);
            load_a(1, it + 1);
          }
        }
      }
      if (has_out) {
        mbar_wait(&bar_gout, (n_my - 1) & 1);
        const int64_t tile = blockIdx.x + (int64_t)(n_my - 1) * gridDim.x;
        bulk_s2g(T.g_out + tile * T.g_out_tile_stride, smem + kTrOffG + ((n_my - 1) & 1) * 32768, 32768u);
        bulk_commit();
        bulk_wait_all();
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_my > 0) {
      // ------------------------------- MMA issuer 1: the data-gradient GEMMs (the serial chain) -------------------------
      // Two issuer threads because every barrier wait / commit costs its thread 200-400 cycles: the weight-gradient
      // GEMMs (own accumulators) are issued by warp 10.  Completion order ACROSS the two threads is not defined, so
      // every buffer hand-over between the two streams is an explicit barrier.
      const uint32_t idesc_d = idesc_bf16(128, 128, 0, 0);
      if (T.step[0].has_dgrad || T.step[1].has_dgrad) mbar_wait(&bar_w, 0);
      for (int it = 0; it < n_my; ++it) {
        for (int j = 0; j < (two ? 2 : 1); ++j) {
          const TrunkStep& s = T.step[j];
          if (!s.has_dgrad) continue;
          const uint32_t gs = smem_u32(smem + kTrOffG + (j == 0 ? (it & 1) : 2) * 32768);
          if (j == 0) {
            mbar_wait(&bar_gin[it & 1], (it >> 1) & 1);
            // accumulator drained by the last epilogue of the previous tile
            if (it >= 1) mbar_wait(has_out ? &bar_gout : &bar_gmid, (it - 1) & 1);
          } else {
            mbar_wait(&bar_gmid, it & 1);                                 // G_mid written, accumulator drained
          }
          tc_fence_after();
          const uint64_t da0 = smem_desc(gs, kChunkBytes, 128);
          const uint64_t db0 = smem_desc(smem_u32(smem + kTrOffWT + j * 32768), kChunkBytes, 128);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mma_ss(tmem + kTrAcc, da0 + (uint64_t)(k * (2 * kChunkBytes >> 4)), db0 + (uint64_t)(k * (2 * kChunkBytes >> 4)), idesc_d, k > 0);
          mma_commit(&bar_acc);
        }
      }
      mma_commit(&bar_done);
    }
  } else if (warp == 10) {
    if (lane == 0 && n_my > 0) {
      // ------------------------------- MMA issuer 2: the weight- and bias-gradient GEMMs -----------------------------------
      const uint32_t idesc_b = idesc_bf16(128, 16, 1, 1);
      const uint32_t ones = smem_u32(smem + kTrOffOnes);
      for (int it = 0; it < n_my; ++it) {
        for (int j = 0; j < (two ? 2 : 1); ++j) {
          const TrunkStep& s = T.step[j];
          const uint32_t gs = smem_u32(smem + kTrOffG + (j == 0 ? (it & 1) : 2) * 32768);
          const uint32_t as = smem_u32(smem + kTrOffA + j * 32768);
          if (j == 0) mbar_wait(&bar_gin[it & 1], (it >> 1) & 1);
          else mbar_wait(&bar_gmid, it & 1);                              // its gradient operand was written by epilogue 0
          mbar_wait(&bar_afull[j], it & 1);
          tc_fence_after();
          const uint32_t idesc_w = idesc_bf16(128, s.act_ch, 1, 1);
          const uint64_t da0 = smem_desc(gs, 128, kChunkBytes);
          const uint64_t db0 = smem_desc(as, 128, kChunkBytes);
          const uint64_t do0 = smem_desc(ones, 128, 0);
          const uint32_t dw = tmem + (j == 0 ? kTrDW0 : kTrDW1), dbias = tmem + (j == 0 ? kTrDB0 : kTrDB1);
#pragma unroll
          for (int k = 0; k < 8; ++k)                   // 16 points per MMA: +256 B in both images
            mma_ss(dw, da0 + (uint64_t)(k * 16), db0 + (uint64_t)(k * 16), idesc_w, (it > 0 || k > 0) ? 1u : 0u);
          if (s.bias) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_ss(dbias, da0 + (uint64_t)(k * 16), do0 + (uint64_t)(k * 16), idesc_b, (it > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(&bar_afree[j]);                    // + 256 epilogue arrivals: the activation slot may be refilled
        }
      }
      mma_commit(&bar_done);
    }
  } else {
    // --------------------------------- epilogue warps ----------------------------------------------------------------
    const int ew = warp - 2;                          // 0..7
    const int half = ew >> 2;                         // which 64 of the 128 output columns
    const int q = warp & 3;                           // TMEM lane quarter
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t acc_ph = 0u;
    for (int it = 0; it < n_my; ++it) {
      const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
      for (int j = 0; j < (two ? 2 : 1); ++j) {
        const TrunkStep& s = T.step[j];
        if (s.has_dgrad) {
          mbar_wait(&bar_afull[j], it & 1);           // the mask comes from the activation image
          // the destination image must be dead in BOTH MMA streams: G_mid was read by the weight-gradient GEMMs of step 1
          // of the previous tile, G_out lands on G_in, read by those of step 0 of this tile
          if (j == 0) { if (it > 0 && two) mbar_wait(&bar_afree[1], (it - 1) & 1); }
          else mbar_wait(&bar_afree[0], it & 1);
          mbar_wait(&bar_acc, acc_ph);                // the accumulator completes last: waited for last
          acc_ph ^= 1u;
          tc_fence_after();
          // input activation of layer j: channels [act_ch - 128, act_ch) of the slot gate the 128 gradient columns
          const uint8_t* a_row = smem + kTrOffA + j * 32768 + (uint32_t)(s.act_ch - 128) * 256u + (half * 8) * kChunkBytes + row * 16;
          uint8_t* dst_row = smem + kTrOffG + (j == 0 ? 2 : (it & 1)) * 32768 + (half * 8) * kChunkBytes + row * 16;
          uint8_t* gsv = s.g_save ? s.g_save + tile * s.g_save_tile_stride + (half * 8) * kChunkBytes + row * 16 : nullptr;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t v[32];
            tmem_ld32(taddr + kTrAcc + half * 64 + h2 * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 a = *reinterpret_cast<const uint4*>(a_row + (h2 * 4 + c) * kChunkBytes);
              uint4 pk;
              pk.x = pack_mask(v[8 * c + 0], v[8 * c + 1], a.x, true);
              pk.y = pack_mask(v[8 * c + 2], v[8 * c + 3], a.y, true);
              pk.z = pack_mask(v[8 * c + 4], v[8 * c + 5], a.z, true);
              pk.w = pack_mask(v[8 * c + 6], v[8 * c + 7], a.w, true);
              *reinterpret_cast<uint4*>(dst_row + (h2 * 4 + c) * kChunkBytes) = pk;
              if (gsv != nullptr) *reinterpret_cast<uint4*>(gsv + (h2 * 4 + c) * kChunkBytes) = pk;
            }
          }
          tc_fence_before();
          fence_async_smem();
          mbar_arrive(j == 0 ? &bar_gmid : &bar_gout);
        }
        mbar_arrive(&bar_afree[j]);
      }
    }
    // ---- flush: TMEM -> vector reductions into the flat fp32 gradient ------------------------------------------------
    if (n_my > 0) {
      mbar_wait(&bar_done, 0);
      tc_fence_after();
      const int n = row;                              // output channel of this thread
      bool any_bulk = false;
      for (int j = 0; j < (two ? 2 : 1); ++j) {
        const TrunkStep& s = T.step[j];
        const uint32_t dw = taddr + (j == 0 ? kTrDW0 : kTrDW1);
        const int c_lo = half * (s.act_ch >> 1), c_hi = c_lo + (s.act_ch >> 1);
        // A weight block that is one contiguous, 16-byte aligned range of the flat buffer (every 128 x 128 trunk layer) leaves
        // as ONE bulk reduction: the accumulator is transposed through shared memory (operand regions are dead by now; the
        // G region is not -- the last G_out store may still be reading it) and the TMA engine adds the 64 KB block.  The
        // per-thread vector reductions it replaces (148 CTAs x 128 x 128 x 2 layers) cost ~25 us per launch whatever the batch.
        const int64_t i00 = packed_weight_index(T.ps, s.pl, 0, s.k_off);
        const int64_t i11 = packed_weight_index(T.ps, s.pl, 127, s.k_off + s.act_ch - 1);
        const bool bulk = T.bulk_flush && i00 >= 0 && (i00 & 3) == 0 && i11 == i00 + (int64_t)128 * s.act_ch - 1;
        if (bulk) {
          float* stage = reinterpret_cast<float*>(smem + (j == 0 ? kTrOffWT : kTrOffA)) + (size_t)n * s.act_ch;   // 64 KB each, both dead
          for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(dw + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 4) *reinterpret_cast<uint4*>(stage + c0 + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          }
          any_bulk = true;
        } else {
          for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(dw + c0, v);
            tmem_ld_wait();
            const int64_t i0 = packed_weight_index(T.ps, s.pl, n, s.k_off + c0);
            const int64_t i15 = packed_weight_index(T.ps, s.pl, n, s.k_off + c0 + 15);
            if (i0 >= 0 && i15 == i0 + 15 && (i0 & 3) == 0) {
#pragma unroll
              for (int e = 0; e < 16; e += 4)
                red_add_v4(T.d_flat + i0 + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int64_t idx = packed_weight_index(T.ps, s.pl, n, s.k_off + c0 + e);
                if (idx >= 0) atomicAdd(T.d_flat + idx, __uint_as_float(v[e]));
              }
            }
          }
        }
        if (s.bias && half == 0) {
          uint32_t v[16];
          tmem_ld16(taddr + (j == 0 ? kTrDB0 : kTrDB1), v);
          tmem_ld_wait();
          const int64_t idx = packed_bias_index(T.ps, s.pl, n);
          if (idx >= 0) atomicAdd(T.d_flat + idx, __uint_as_float(v[0]));
        }
      }
      if (any_bulk) {                                  // uniform over the 256 epilogue threads
        fence_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0 && lane == 0) {
          for (int j = 0; j < (two ? 2 : 1); ++j) {
            const TrunkStep& s = T.step[j];
            const int64_t i00 = packed_weight_index(T.ps, s.pl, 0, s.k_off);
            const int64_t i11 = packed_weight_index(T.ps, s.pl, 127, s.k_off + s.act_ch - 1);
            if (!(i00 >= 0 && (i00 & 3) == 0 && i11 == i00 + (int64_t)128 * s.act_ch - 1)) continue;
            const uint32_t bytes = 128u * (uint32_t)s.act_ch * 4u;
            for (uint32_t off = 0; off < bytes; off += 16384u)
              bulk_red_add_f32(T.d_flat + i00 + off / 4, smem + (j == 0 ? kTrOffWT : kTrOffA) + off, min(16384u, bytes - off));
          }
          bulk_commit();
          bulk_wait_all();
        }
      }
      tc_fence_before();
    }
  }
  }  // passes
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace nefes
