// Fused backward of the HEAD layers of the NeFeS field (training): data gradients AND weight gradients of a group of
// layers in one pass over the 128-point tiles, weight gradients accumulated in TMEM for the whole launch -- the same idea
// as mlp_trunk_bwd.cuh, but the head layers are irregular (K = 16 / 64 / 144, M = 16 / 64 / 131 / 144, merged and
// split inputs), so the kernel is a small interpreter: the host compiles, per group of layers, three straight-line
// programs over a per-tile set of mbarriers, every one of which completes exactly ONE phase per tile:
//
//   producer (1 thread)      WAIT bar | LOAD image -> smem (bulk copy, completes on bar) | STORE smem -> image | ARRIVE bar
//   MMA issuers (2 threads)  WAIT bar | MMA group (descriptors relative to the smem base, fresh / per-launch accumulate)
//                            | COMMIT bar (tcgen05.commit: all MMAs issued so far BY THIS THREAD retired).
//                            Thread 0 issues the data-gradient GEMMs (the serial chain layer -> epilogue -> layer), thread 1
//                            the weight-gradient GEMMs (independent accumulators): every wait / commit costs its thread
//                            200-400 cycles, and with one issuer those latencies were 60 % of the tile time.  Completion
//                            order ACROSS the two threads is not defined, so every buffer hand-over between the two
//                            streams is an explicit barrier in the programs.
//   epilogue (256 threads)   WAIT bar | EPI: TMEM acc -> ReLU mask from a saved activation image in smem (a > 0) -> bf16
//                            gradient image in smem (operand of the next layer) | ARRIVE bar
//
// A wait names a barrier of the CURRENT tile (parity it & 1) or of the PREVIOUS tile (flag, skipped on the first tile).
// Loads flagged "next" fetch the operand of tile it+1 as soon as its slot of the current tile is dead, and run once for
// the first tile before the loop.  Bias gradients ride in the weight-gradient MMAs as 16 extra "ones" channels appended
// to the activation slot (or a spare padding channel of the encoding image that the encoder sets to 1).
// After the last tile the accumulators are scattered into the flat gradient buffer (vector reductions where aligned).
// Included by mlp_tc.cu.   script/models/nerfh_nff.py:478-505, :555-576 (heads), backward.
#pragma once

namespace nefes {

enum { FO_END = 0, FO_WAIT, FO_LOAD, FO_STORE, FO_ARRIVE, FO_MMA, FO_COMMIT, FO_EPI };
enum { FX_THEN = 8 };                       // MMA op: then COMMIT to `bar`; EPI op: then ARRIVE on `bar` (one interpreter step less on the critical path)
enum { FW_PREV = 1, FW_ONCE = 2 };          // wait flags: the barrier phase of the previous tile / first tile only (weights)
enum { FA_FRESH = 0, FA_LAUNCH = 1 };       // MMA accumulate mode: zero-init per group / accumulate over the launch
enum { FF_W = 0, FF_BIAS = 1, FF_WT = 2 };  // flush kinds

struct FProdOp {
  uint8_t kind, bar, flags, next;           // LOAD: next = 1 -> fetches for tile it+1; 2 -> once per launch (weights)
  uint32_t smem_off, bytes, tile_stride;
  const uint8_t* src; uint8_t* dst;         // LOAD: src image; STORE: dst image
};
struct FMmaOp {
  uint8_t kind, bar, flags, accmode;
  uint8_t ksteps, pad[3];
  uint32_t a_off, b_off;                    // operand starts (bytes from the smem base)
  uint16_t a_lbo, a_sbo, b_lbo, b_sbo;      // descriptor strides in 16-byte units
  uint16_t a_adv, b_adv;                    // descriptor advance per k-step in 16-byte units
  uint16_t tmem_col, pad2;
  uint32_t idesc;
  // host-built at finish(): the two shared-memory descriptors relative to the smem base (the kernel adds base >> 4 to the
  // address field -- 14 bits, no carry below 256 KB) and the small fields packed into one word: the issuer is a single warp
  // at ~4 cycles per instruction, and building these from the fields above was ~60 of the ~100 instructions of an MMA step
  uint64_t da_rel, db_rel;
  uint32_t misc;                            // tmem_col | ksteps << 16 | accmode << 24
  uint32_t adv;                             // a_adv | b_adv << 16
};
struct FEpiOp {
  uint8_t kind, bar, flags, has_mask;
  uint16_t acc_col, n;                      // accumulator columns [acc_col, acc_col + n), n = 64 or 128
  uint32_t mask_off, out_off;               // activation image gating the gradient / destination image (smem)
};
struct FFlush {
  uint16_t tmem_col, n_cols; int16_t pl, m0, k_off; uint8_t kind, pad;
};
constexpr int kFMaxProd = 44, kFMaxMma = 28, kFMaxEpi = 28, kFMaxFlush = 12, kFMaxBars = 32;
struct FusedArgs {
  FProdOp prod[kFMaxProd];
  FMmaOp mma[2][kFMaxMma];
  FEpiOp epi[kFMaxEpi];
  FFlush flush[kFMaxFlush];
  uint16_t bar_count[kFMaxBars];            // arrival count of every barrier (0: unused)
  int n_flush, n_tiles;
  uint32_t ones_off[4]; int n_ones;         // 4 KB blocks of bf16 ones (16 channels x 128 points) appended to activation slots
  uint32_t zero_off, zero_bytes;            // smem range zero-filled once (padding rows of short gradient images)
  PackSrc ps; float* d_flat;
  long long* dbg;                           // optional clock stamps of CTA 0: [role 0..2][tile 0..3][op 0..47]
};
constexpr int kFusedThreads = 64 + 256 + 32;     // producer, dgrad issuer, 8 epilogue warps, wgrad issuer
#ifdef NEFES_FUSED_DBG
constexpr bool kFusedDbg = true;
#else
constexpr bool kFusedDbg = false;
#endif

__global__ void __launch_bounds__(kFusedThreads, 1) fused_bwd_kernel(const __grid_constant__ FusedArgs F) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[kFMaxBars];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  // the three programs are interpreted from shared memory: a dynamically indexed read of the kernel-parameter bank costs
  // a constant-cache miss (~300 cycles) per op field once the table exceeds the cache, which dominated the tile time
  __shared__ FProdOp s_prod[kFMaxProd];
  __shared__ FMmaOp s_mma[2][kFMaxMma];
  __shared__ FEpiOp s_epi[kFMaxEpi];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kFMaxProd; i += kFusedThreads) s_prod[i] = F.prod[i];
  for (int i = threadIdx.x; i < 2 * kFMaxMma; i += kFusedThreads) s_mma[i / kFMaxMma][i % kFMaxMma] = F.mma[i / kFMaxMma][i % kFMaxMma];
  for (int i = threadIdx.x; i < kFMaxEpi; i += kFusedThreads) s_epi[i] = F.epi[i];

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFMaxBars; ++i) mbar_init(&bars[i], F.bar_count[i] ? F.bar_count[i] : 1);
    mbar_init(&bar_done, 2);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  {
    uint4 ones;
    ones.x = ones.y = ones.z = ones.w = 0x3F803F80u;
    for (int b = 0; b < F.n_ones; ++b)
      for (int i = threadIdx.x; i < 4096 / 16; i += kFusedThreads) reinterpret_cast<uint4*>(smem + F.ones_off[b])[i] = ones;
    for (uint32_t i = threadIdx.x; i < F.zero_bytes / 16; i += kFusedThreads)
      reinterpret_cast<uint4*>(smem + F.zero_off)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  int n_my = 0;
  for (int t = blockIdx.x; t < F.n_tiles; t += gridDim.x) ++n_my;

  if (warp == 0) {
    if (lane == 0 && n_my > 0) {
      // ------------------------------- producer ---------------------------------------------------------------------
      auto do_load = [&](const FProdOp& o, int it) {
        const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
        mbar_arrive_expect_tx(&bars[o.bar], o.bytes);
        bulk_g2s(smem + o.smem_off, o.src + tile * o.tile_stride, o.bytes, &bars[o.bar]);
      };
      for (int i = 0; s_prod[i].kind != FO_END; ++i)            // operands of the first tile
        if (s_prod[i].kind == FO_LOAD && s_prod[i].next) do_load(s_prod[i], 0);
      for (int it = 0; it < n_my; ++it) {
        for (int i = 0; s_prod[i].kind != FO_END; ++i) {
          const FProdOp& o = s_prod[i];
          if (o.kind == FO_WAIT) {
            if (o.flags & FW_PREV) { if (it > 0) mbar_wait(&bars[o.bar], (it - 1) & 1); }
            else if (o.flags & FW_ONCE) { if (it == 0) mbar_wait(&bars[o.bar], 0); }
            else mbar_wait(&bars[o.bar], it & 1);
          } else if (o.kind == FO_LOAD) {
            if (o.next == 1) { if (it + 1 < n_my) do_load(o, it + 1); }
            else if (o.next == 0) do_load(o, it);
          } else if (o.kind == FO_STORE) {
            const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
            bulk_s2g(o.dst + tile * o.tile_stride, smem + o.smem_off, o.bytes);
            bulk_commit();
            bulk_wait_read<0>();                                  // the slot may be overwritten once this returns
          } else if (o.kind == FO_ARRIVE) {
            mbar_arrive(&bars[o.bar]);
          }
          if (kFusedDbg && F.dbg != nullptr && blockIdx.x == 0 && it < 4 && i < 48) F.dbg[(0 * 4 + it) * 48 + i] = clock64();
        }
      }
      bulk_wait_all();
    }
  } else if (warp == 1 || warp == 10) {
    if (n_my > 0) {
      // ------------------------------- MMA issuers: warp 1 data gradients, warp 10 weight gradients ----------------------
      // CONVERGED warp: all lanes walk the program and wait, one elected lane issues with uniform operands (tc05.cuh uni()).
      // From a lane-0 branch every tcgen05.mma / commit is wrapped in a waterfall loop (50-100 cycles each), and warp
      // sampling had these two threads 65 % busy in their own code with one tile in flight (round 2).
      const bool leader = elect_one();
      const int which = uni(warp == 1 ? 0 : 1);
      const FMmaOp* prog = s_mma[which];
      const uint32_t tm = uni(tmem);
      const uint64_t sb16 = uni((uint64_t)(sbase >> 4));
      for (int it = 0; it < n_my; ++it) {
        for (int i = 0; prog[i].kind != FO_END; ++i) {
          const FMmaOp& o = prog[i];
          const int kind = uni((int)o.kind), bar = uni((int)o.bar), flags = uni((int)o.flags);
          if (kind == FO_WAIT) {
            if (flags & FW_PREV) { if (it > 0) mbar_wait(&bars[bar], (it - 1) & 1); }
            else if (flags & FW_ONCE) { if (it == 0) mbar_wait(&bars[bar], 0); }
            else mbar_wait(&bars[bar], it & 1);
          } else if (kind == FO_MMA) {
            tc_fence_after();
            const uint64_t da0 = uni(o.da_rel) + sb16, db0 = uni(o.db_rel) + sb16;
            const uint32_t misc = uni(o.misc), adv = uni(o.adv), idesc = uni(o.idesc);
            const uint32_t d = tm + (misc & 0xFFFFu);
            const uint32_t acc0 = ((misc >> 24) == FA_LAUNCH && it > 0) ? 1u : 0u;
            const int ks = (int)((misc >> 16) & 0xFFu);
            const uint64_t a_adv = adv & 0xFFFFu, b_adv = adv >> 16;
            if (leader) {
              for (int k = 0; k < ks; ++k) mma_ss(d, da0 + (uint64_t)k * a_adv, db0 + (uint64_t)k * b_adv, idesc, (k > 0) ? 1u : acc0);
              if (flags & FX_THEN) mma_commit(&bars[bar]);
            }
          } else if (kind == FO_COMMIT) {
            if (leader) mma_commit(&bars[bar]);
          }
          if (kFusedDbg && F.dbg != nullptr && blockIdx.x == 0 && warp == 1 && lane == 0 && it < 4 && i < 48) F.dbg[(1 * 4 + it) * 48 + i] = clock64();
          __syncwarp();
        }
      }
      if (leader) mma_commit(&bar_done);
      __syncwarp();
    }
  } else {
    // --------------------------------- epilogue warps ----------------------------------------------------------------
    const int ew = warp - 2;                          // 0..7
    const int half = ew >> 2;                         // which half of the output columns
    const int q = warp & 3;                           // TMEM lane quarter
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
    for (int it = 0; it < n_my; ++it) {
      for (int i = 0; s_epi[i].kind != FO_END; ++i) {
        const FEpiOp& o = s_epi[i];
        if (o.kind == FO_WAIT) {
          if (o.flags & FW_PREV) { if (it > 0) mbar_wait(&bars[o.bar], (it - 1) & 1); }
          else mbar_wait(&bars[o.bar], it & 1);
        } else if (o.kind == FO_ARRIVE) {
          mbar_arrive(&bars[o.bar]);
        } else if (o.kind == FO_EPI) {
          tc_fence_after();
          const int ncol = o.n >> 1;                   // this warp's columns: [half * ncol, +ncol), 32 or 64
          const int c0 = half * ncol;
          const uint8_t* a_row = smem + o.mask_off + (c0 >> 3) * kChunkBytes + row * 16;
          uint8_t* dst_row = smem + o.out_off + (c0 >> 3) * kChunkBytes + row * 16;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (h2 * 32 < ncol) {
              uint32_t v[32];
              tmem_ld32(taddr + o.acc_col + c0 + h2 * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint4 a = make_uint4(0u, 0u, 0u, 0u);
                if (o.has_mask) a = *reinterpret_cast<const uint4*>(a_row + (h2 * 4 + c) * kChunkBytes);
                uint4 pk;
                pk.x = pack_mask(v[8 * c + 0], v[8 * c + 1], a.x, o.has_mask);
                pk.y = pack_mask(v[8 * c + 2], v[8 * c + 3], a.y, o.has_mask);
                pk.z = pack_mask(v[8 * c + 4], v[8 * c + 5], a.z, o.has_mask);
                pk.w = pack_mask(v[8 * c + 6], v[8 * c + 7], a.w, o.has_mask);
                *reinterpret_cast<uint4*>(dst_row + (h2 * 4 + c) * kChunkBytes) = pk;
              }
            }
          }
          tc_fence_before();
          fence_async_smem();
          if (o.flags & FX_THEN) mbar_arrive(&bars[o.bar]);
        }
        if (kFusedDbg && F.dbg != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 4 && i < 48) F.dbg[(2 * 4 + it) * 48 + i] = clock64();
      }
    }
    // ---- flush: TMEM -> reductions into the flat fp32 gradient ----------------------------------------------------------
    if (n_my > 0) {
      mbar_wait(&bar_done, 0);
      tc_fence_after();
      for (int f = 0; f < F.n_flush; ++f) {
        const FFlush& fl = F.flush[f];
        if (fl.kind == FF_BIAS) {
          if (half == 0) {
            uint32_t v[16];
            tmem_ld16(taddr + fl.tmem_col, v);
            tmem_ld_wait();
            float val = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) if (e == fl.k_off) val = __uint_as_float(v[e]);   // k_off: column inside the block
            const int64_t idx = packed_bias_index(F.ps, fl.pl, fl.m0 + row);
            if (idx >= 0 && fl.m0 + row < packed_dims(fl.pl).N) atomicAdd(F.d_flat + idx, val);
          }
          continue;
        }
        // the two column halves of the accumulator go to the two warp groups (16-column granularity)
        const int nblk = fl.n_cols >> 4, b_lo = half * ((nblk + 1) >> 1), b_hi = half == 0 ? ((nblk + 1) >> 1) : nblk;
        for (int b = b_lo; b < b_hi; ++b) {
          uint32_t v[16];
          tmem_ld16(taddr + fl.tmem_col + b * 16, v);
          tmem_ld_wait();
          if (fl.kind == FF_WT) {                     // transposed accumulator: thread = input channel, column = output row
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int n = fl.m0 + b * 16 + e;
              const int64_t idx = n < packed_dims(fl.pl).N ? packed_weight_index(F.ps, fl.pl, n, fl.k_off + row) : -1;
              if (idx >= 0 && b * 16 + e < 1) atomicAdd(F.d_flat + idx, __uint_as_float(v[e]));   // only column 0 is real
            }
            continue;
          }
          const int n = fl.m0 + row;                  // output channel of this thread
          if (n >= packed_dims(fl.pl).N) continue;
          const int64_t i0 = packed_weight_index(F.ps, fl.pl, n, fl.k_off + b * 16);
          const int64_t i15 = packed_weight_index(F.ps, fl.pl, n, fl.k_off + b * 16 + 15);
          if (i0 >= 0 && i15 == i0 + 15 && (i0 & 3) == 0) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              red_add_v4(F.d_flat + i0 + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int64_t idx = packed_weight_index(F.ps, fl.pl, n, fl.k_off + b * 16 + e);
              if (idx >= 0) atomicAdd(F.d_flat + idx, __uint_as_float(v[e]));
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace nefes
