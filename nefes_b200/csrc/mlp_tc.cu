// K5, NEFES_PREC_BF16 path: the NeFeS field MLP on the 5th-generation tensor cores.
//
//   tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) issued by one thread, accumulators in TMEM,
//   operands in shared memory as UMMA "interleaved" images, tiles moved by 1-D bulk async copies
//   (TMA engine, UBLKCP) completing on mbarriers; warp-specialised persistent CTAs:
//   warp 0 = copy producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM -> registers ->
//   bias / activation / ReLU bit-mask -> bf16 image in shared memory -> bulk store).
//
// Three kernels:
//   tile_gemm_kernel   D[128 pts, N] = A_tile[128, K] * W^T        forward layers and dgrad
//                      (weights stay resident in shared memory for the whole launch)
//   wgrad_kernel       dW[N, K] += G_tile^T[N, 128 pts] * A_tile[128 pts, K] over all tiles, both
//                      operands read MN-major from the SAME images forward/dgrad wrote; bias
//                      gradients fall out of an extra MMA against a tile of ones
//   plus small SIMT kernels: weight repack, PE -> images, head gradients -> images.
// script/models/nerfh_nff.py:168-231, :525-576.
#include <cstdlib>

#include "tc05.cuh"
#include "tc_layers.cuh"

namespace nefes {
using namespace tc05;

constexpr int kTile = 128;
constexpr uint32_t kChunkBytes = kTile * 16;          // one 8-channel chunk of a tile image = 2 KB
constexpr int kEpiWarp0 = 2;                          // warps 2..5 are the epilogue
constexpr int kThreads = 192;

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
// the same with an L2 evict_first policy: data nobody re-reads before it leaves L2 (saved activation images, 2.65 GB per launch)
__device__ __forceinline__ uint64_t l2_evict_first_policy() { return l2_policy_evict_first(); }
__device__ __forceinline__ void bulk_s2g_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol) : "memory");
}
// bulk reduction shared -> global: the TMA engine adds a contiguous fp32 block into global memory (split accumulators of the
// persistent CTAs into the flat gradient buffer); part of the same bulk groups as the stores
__device__ __forceinline__ void bulk_red_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
struct ASrc { const uint8_t* base; int64_t tile_stride; uint32_t bytes; };

// =================================================================================================
// tile_gemm_kernel
// =================================================================================================
enum { RAW_ACT_NONE = 0, RAW_ACT_SOFTPLUS = 2, RAW_ACT_THEADS = 4 };
enum { EPI_HIDDEN = 0,    // bias + ReLU + ReLU bit-mask out -> bf16 image            (forward hidden layers)
       EPI_DGRAD = 1,     // ReLU bit-mask in (optional)     -> bf16 image            (backward data gradients)
       EPI_IMG = 2,       // bias -> bf16 image for columns < out_ch (may be none); plus up to 5 activated fp32
                          // columns starting at d_col0 (a multiple of 16), stored directly     (final+sigma, heads)
       EPI_RAWBULK = 3 }; // bias -> many fp32 columns, staged in shared memory, coalesced rows   (rgb+feature head)

constexpr int kEpiWarps = 8;                          // two warps per TMEM lane quarter, interleaved 16-column blocks
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGemmThreads = 64 + kEpiThreads;
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct TileGemmArgs {
  ASrc a[2]; int n_src;
  int K, N;                        // MMA K (multiple of 16) and N (multiple of 16, <= 256)
  const uint8_t* w_img;            // B image [K/8][w_rows][8] bf16
  int w_rows, w_row0;              // image row count, first row used
  const float* bias;               // [N] or null
  int n_tiles; int64_t M;
  // bf16 image output: D columns [0, out_ch)
  uint8_t* out_img; int64_t out_tile_stride; int out_ch; int relu;
  uint4* mask_out; const uint4* mask_in; int mask_shift;      // mask_shift: first mask bit of D column 0 (multiple of 16)
  // fp32 output: D columns [d_col0, d_col0 + raw_ncol) -> raw[row*raw_ld + raw_col0 + i]
  float* raw; int raw_ld, raw_col0, d_col0, raw_ncol, raw_act;
  // EPI_RAWBULK reductions of the staged fp32 rows instead of storing them (input gradients of the refinement path):
  //   pe_x != null: the rows are cotangents of the frequency encoding of pe_x [M,3] -> raw = d_x [M,3]
  //   ray_S > 0   : rows summed over each ray's ray_S consecutive points -> raw [n_rays, raw_ld]
  const float* pe_x; int pe_L; int ray_S; int64_t n_rays;
  // shared-memory carve-up (bytes from the dynamic base), computed on the host
  uint32_t off_a, a_stage_stride, off_out, out_bytes, off_raw, raw_pitch, off_bias;
};

__device__ __forceinline__ float raw_activation(float x, int act, int col) {
  if (act == RAW_ACT_SOFTPLUS) return softplus_f(x);
  if (act == RAW_ACT_THEADS) return col < 3 ? sigmoid_f(x) : softplus_f(x);
  return x;
}

// one 16-column block of one row: v = accumulators, c0 = first D column
template <int EPI>
__device__ __forceinline__ void epi_block(const TileGemmArgs& g, const uint32_t (&v)[16], int c0, int row,
                                          const float* __restrict__ sBias, uint8_t* __restrict__ sOut,
                                          float* __restrict__ sRaw, uint32_t in16, uint32_t& out16) {
  out16 = 0u;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = c0 + h * 8;
    float x[8];
    if (EPI == EPI_DGRAD) {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = ((in16 >> (h * 8 + e)) & 1u) ? __uint_as_float(v[h * 8 + e]) : 0.f;
    } else {
      const float4 b0 = *reinterpret_cast<const float4*>(sBias + c);
      const float4 b1 = *reinterpret_cast<const float4*>(sBias + c + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[h * 8 + e]) + bb[e];
    }
    if (EPI == EPI_HIDDEN) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool on = x[e] > 0.f;
        x[e] = on ? x[e] : 0.f;
        out16 |= on ? (1u << (h * 8 + e)) : 0u;
      }
    }
    if (EPI == EPI_HIDDEN || EPI == EPI_DGRAD || (EPI == EPI_IMG && c < g.out_ch)) {
      uint4 pk;
      pk.x = pack_bf16(x[0], x[1]); pk.y = pack_bf16(x[2], x[3]);
      pk.z = pack_bf16(x[4], x[5]); pk.w = pack_bf16(x[6], x[7]);
      *reinterpret_cast<uint4*>(sOut + (c >> 3) * kChunkBytes + row * 16) = pk;
    }
    if (EPI == EPI_RAWBULK) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c + e < g.raw_ncol) sRaw[row * g.raw_pitch + c + e] = x[e];
    }
  }
}

template <int STAGES, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1) tile_gemm_kernel(const TileGemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_full[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sW = smem;
  float* sBias = reinterpret_cast<float*>(smem + g.off_bias);

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], kEpiThreads); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  for (int i = threadIdx.x; i < g.N; i += kGemmThreads) sBias[i] = g.bias ? g.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_bytes = (uint32_t)g.K * 256u;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer: weights once, then the activation tiles ----------------------
      const uint32_t w_bytes = (uint32_t)g.K * g.w_rows * 2u;
      mbar_arrive_expect_tx(&bar_w, w_bytes);
      for (uint32_t off = 0; off < w_bytes; off += 16384u)
        bulk_g2s(sW + off, g.w_img + off, min(16384u, w_bytes - off), &bar_w);
      int it = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], a_bytes);
        uint8_t* dst = smem + g.off_a + s * g.a_stage_stride;
        for (int q = 0; q < g.n_src; ++q) {
          bulk_g2s(dst, g.a[q].base + (int64_t)tile * g.a[q].tile_stride, g.a[q].bytes, &bar_full[s]);
          dst += g.a[q].bytes;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer -------------------------------------------------------------
      const uint32_t idesc = idesc_bf16(128, g.N, 0, 0);
      const uint32_t w_base = smem_u32(sW) + (uint32_t)g.w_row0 * 16u;
      const uint32_t w_lbo = (uint32_t)g.w_rows * 16u;
      mbar_wait(&bar_w, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&bar_tempty[acc], aph ^ 1);
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + g.off_a + s * g.a_stage_stride);
        const uint32_t d = tmem + acc * 256;
        const uint64_t da0 = smem_desc(a_base, kChunkBytes, 128);
        const uint64_t db0 = smem_desc(w_base, w_lbo, 128);
        for (int k = 0; k < g.K / 16; ++k)              // one K step = two 8-channel chunks of both images
          mma_ss(d, da0 + (uint64_t)(k * (2 * kChunkBytes >> 4)), db0 + (uint64_t)(k * (2 * w_lbo >> 4)), idesc, k > 0);
        mma_commit(&bar_empty[s]);      // the A stage may be refilled once these MMAs retire
        mma_commit(&bar_tfull[acc]);    // ... and the accumulator may be drained
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; thread = one point (TMEM lane) x every other 16-column block ----
    const int q = warp & 3;                           // TMEM lane quarter this warp may touch
    const int half = (warp - kEpiWarp0) >> 2;         // which 16-column blocks: b & 1 == half
    const int row = q * 32 + lane;
    const int et = (warp - kEpiWarp0) * 32 + lane;    // 0..255
    float* sRaw = reinterpret_cast<float*>(smem + g.off_raw);
    const bool raw_staged = (EPI == EPI_RAWBULK);
    const bool has_img = (EPI == EPI_HIDDEN) || (EPI == EPI_DGRAD) || (EPI == EPI_IMG && g.out_img != nullptr);
    const int n_blk = g.N >> 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int64_t grow = (int64_t)tile * kTile + row;
      uint8_t* sOut = smem + g.off_out + (it & 1) * g.out_bytes;
      if (has_img) {
        if (et == 0) bulk_wait_read<1>();             // the store that used this staging buffer two tiles ago
        epi_barrier();
      }
      const uint16_t* min_row = (EPI == EPI_DGRAD && g.mask_in != nullptr) ? reinterpret_cast<const uint16_t*>(g.mask_in + grow) : nullptr;
      uint16_t* mout_row = (g.mask_out != nullptr) ? reinterpret_cast<uint16_t*>(g.mask_out + grow) : nullptr;
      mbar_wait(&bar_tfull[acc], aph);
      tc_fence_after();
      const uint32_t taddr = tmem + acc * 256 + ((uint32_t)(q * 32) << 16);
      // my blocks: half, half+2, ...; loaded four at a time so the TMEM reads overlap
      for (int b0 = half; b0 < n_blk; b0 += 8) {
        uint32_t v0[16], v1[16], v2[16], v3[16];
        const bool h1 = b0 + 2 < n_blk, h2 = b0 + 4 < n_blk, h3 = b0 + 6 < n_blk;
        tmem_ld16(taddr + b0 * 16, v0);
        if (h1) tmem_ld16(taddr + (b0 + 2) * 16, v1);
        if (h2) tmem_ld16(taddr + (b0 + 4) * 16, v2);
        if (h3) tmem_ld16(taddr + (b0 + 6) * 16, v3);
        uint32_t in0 = 0xffffu, in1 = 0xffffu, in2 = 0xffffu, in3 = 0xffffu;
        if (min_row != nullptr) {
          const int mb = (g.mask_shift >> 4) + b0;
          in0 = min_row[mb];
          if (h1) in1 = min_row[mb + 2];
          if (h2) in2 = min_row[mb + 4];
          if (h3) in3 = min_row[mb + 6];
        }
        tmem_ld_wait();
        uint32_t o0, o1 = 0, o2 = 0, o3 = 0;
        epi_block<EPI>(g, v0, b0 * 16, row, sBias, sOut, sRaw, in0, o0);
        if (h1) epi_block<EPI>(g, v1, (b0 + 2) * 16, row, sBias, sOut, sRaw, in1, o1);
        if (h2) epi_block<EPI>(g, v2, (b0 + 4) * 16, row, sBias, sOut, sRaw, in2, o2);
        if (h3) epi_block<EPI>(g, v3, (b0 + 6) * 16, row, sBias, sOut, sRaw, in3, o3);
        if (mout_row != nullptr) {
          const int mb = (g.mask_shift >> 4) + b0;
          mout_row[mb] = (uint16_t)o0;
          if (h1) mout_row[mb + 2] = (uint16_t)o1;
          if (h2) mout_row[mb + 4] = (uint16_t)o2;
          if (h3) mout_row[mb + 6] = (uint16_t)o3;
        }
      }
      if (EPI == EPI_IMG && g.raw != nullptr) {
        // a handful of activated fp32 head outputs (sigma; transient rgb/sigma/beta): one copy of the
        // activation code, outside the unrolled block loop
        const int bF = g.d_col0 >> 4;
        if ((bF & 1) == half) {
          uint32_t v[16];
          tmem_ld16(taddr + bF * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 5; ++e)
            if (e < g.raw_ncol && grow < g.M)
              g.raw[grow * g.raw_ld + g.raw_col0 + e] = raw_activation(__uint_as_float(v[e]) + sBias[g.d_col0 + e], g.raw_act, e);
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_tempty[acc]);                  // accumulator drained
      if (has_img) {
        fence_async_smem();                           // generic smem writes -> async proxy
        epi_barrier();
        if (et == 0) {
          bulk_s2g(g.out_img + (int64_t)tile * g.out_tile_stride, sOut, g.out_bytes);
          bulk_commit();
        }
      }
      if (raw_staged) {
        epi_barrier();
        if (g.pe_x != nullptr) {
          // rows = d(encoding) of one point each: chain through [x, sin(2^l x), cos(2^l x)] (nerfh_nff.py:241-270), one
          // (point, coordinate) per thread; x and d_x are [M,3], so consecutive threads touch consecutive floats
          for (int i = et; i < kTile * 3; i += kEpiThreads) {
            const int rr = i / 3, c = i - 3 * rr;
            const int64_t gr = (int64_t)tile * kTile + rr;
            if (gr < g.M) {
              const float v = g.pe_x[gr * 3 + c];
              const float* gi = sRaw + rr * g.raw_pitch + c;
              // sin / cos of 2^l x by angle doubling from one sincosf: ten full-range sincosf per (point, coordinate) made this
              // epilogue -- not the GEMM or its 315 MB of operands -- the bound of the launch (0.12 ms at 4800 rays).  The
              // doubling error (~2^l ulp) is far below the bf16 rounding of the gradients that arrive here.
              float acc = gi[0], f = 1.f, sn, co;
              sincosf(v, &sn, &co);
              for (int l = 0; l < g.pe_L; ++l, f *= 2.f) {
                acc += f * (gi[3 + 6 * l] * co - gi[6 + 6 * l] * sn);
                const float s2 = 2.f * sn * co;
                co = 1.f - 2.f * sn * sn;
                sn = s2;
              }
              g.raw[gr * 3 + c] = acc;
            }
          }
        } else if (g.ray_S > 0) {
          // per-ray sums over the ray's ray_S consecutive rows (a multiple of 16 dividing 128): 16-row partial sums by
          // (column, part) threads, then one thread per (ray, column)
          float* sPart = sRaw + kTile * g.raw_pitch;                   // [8][32]
          const int c = et & 31, part = et >> 5;
          if (c < g.raw_ncol) {
            float a = 0.f;
#pragma unroll 4
            for (int r2 = 0; r2 < 16; ++r2) a += sRaw[(part * 16 + r2) * g.raw_pitch + c];
            sPart[part * 32 + c] = a;
          }
          epi_barrier();
          const int rays_per_tile = kTile / g.ray_S, parts_per_ray = g.ray_S >> 4;
          if (part < rays_per_tile && c < g.raw_ncol) {
            const int64_t ray = (int64_t)tile * rays_per_tile + part;
            if (ray < g.n_rays) {
              float a = 0.f;
              for (int q2 = 0; q2 < parts_per_ray; ++q2) a += sPart[(part * parts_per_ray + q2) * 32 + c];
              g.raw[ray * g.raw_ld + g.raw_col0 + c] = a;
            }
          }
        } else {
          for (int rr = warp - kEpiWarp0; rr < kTile; rr += kEpiWarps) {
            const int64_t gr = (int64_t)tile * kTile + rr;
            if (gr < g.M)
              for (int c = lane; c < g.raw_ncol; c += 32) g.raw[gr * g.raw_ld + g.raw_col0 + c] = sRaw[rr * g.raw_pitch + c];
          }
        }
        epi_barrier();
      }
    }
    if (et == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// =================================================================================================
// wgrad_kernel: persistent CTAs over a cost-balanced list of (layer job, tile range) pieces
// =================================================================================================
constexpr int kMaxJobs = 18;
struct WgradJob {
  ASrc g; int g_ch;                // gradient image (MN = output channel)
  ASrc a[2]; int n_src; int a_ch;  // activation image(s) (MN = input channel)
  int pl;                          // packed layer (scatter map)
  int no_bias;                     // 1: the bias gradient of this layer is owned by another kernel
  int64_t cost_begin;              // prefix sum of tiles*bytes_per_tile over jobs
  uint32_t tile_bytes;
};
struct WgradArgs {
  WgradJob job[kMaxJobs]; int n_jobs;
  int64_t total_cost;
  int n_tiles;
  PackSrc ps; float* d_flat;
  uint32_t off_ones, off_stage, ring_bytes;
};

constexpr int kWgradMaxStages = 4;
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const WgradArgs w) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kWgradMaxStages], bar_empty[kWgradMaxStages], bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgradMaxStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // this CTA's share of the cost line
  const int64_t c_lo = w.total_cost * blockIdx.x / gridDim.x;
  const int64_t c_hi = w.total_cost * (blockIdx.x + 1) / gridDim.x;

  uint32_t ph_bits = 0;   // per-role phase bit of every ring slot (producer: empty barriers, MMA: full barriers)
  int piece = 0;          // flush counter
  for (int j = 0; j < w.n_jobs; ++j) {
    const WgradJob& J = w.job[j];
    const int64_t jb = J.cost_begin, je = jb + (int64_t)w.n_tiles * J.tile_bytes;
    const int64_t lo = max(c_lo, jb), hi = min(c_hi, je);
    if (lo >= hi) continue;
    // tiles whose START lies in [lo, hi)
    const int t0 = (int)((lo - jb + J.tile_bytes - 1) / J.tile_bytes);
    const int t1 = (int)((hi - jb + J.tile_bytes - 1) / J.tile_bytes);
    if (t0 >= t1) continue;
    const int n_mblk = (J.g_ch > 128) ? 2 : 1;
    const uint32_t g_bytes = (uint32_t)J.g_ch * 256u;
    // ring geometry of this job: as many slots as fit (the M=128 gradient operand may read past a short image)
    // slot = [gradient image | activation image(s) | 16 channels of ones]: the ones make the bias gradient
    // (column sums of G) fall out of the same MMA as 16 extra output columns.
    const uint32_t stride = (J.tile_bytes + 2 * kChunkBytes + 1023u) & ~1023u;
    const uint32_t reach = (uint32_t)n_mblk * 32768u;
    const uint32_t over = reach > stride ? reach - stride : 0u;
    int ns = (int)((w.ring_bytes - over) / stride);
    ns = ns > kWgradMaxStages ? kWgradMaxStages : ns;
    {
      uint4 ones;
      ones.x = ones.y = ones.z = ones.w = 0x3F803F80u;
      for (int sl = 0; sl < ns; ++sl) {
        uint4* dst = reinterpret_cast<uint4*>(smem + w.off_stage + sl * stride + J.tile_bytes);
        for (int i = threadIdx.x; i < 2 * (int)kChunkBytes / 16; i += kThreads) dst[i] = ones;
      }
      fence_async_smem();
      __syncthreads();
    }
    if (warp == 0) {
      if (lane == 0) {
        int s = 0;
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&bar_empty[s], ((ph_bits >> s) & 1u) ^ 1u);
          ph_bits ^= 1u << s;
          mbar_arrive_expect_tx(&bar_full[s], J.tile_bytes);
          uint8_t* dst = smem + w.off_stage + s * stride;
          bulk_g2s(dst, J.g.base + (int64_t)t * J.g.tile_stride, J.g.bytes, &bar_full[s]);
          dst += g_bytes;
          for (int q = 0; q < J.n_src; ++q) {
            bulk_g2s(dst, J.a[q].base + (int64_t)t * J.a[q].tile_stride, J.a[q].bytes, &bar_full[s]);
            dst += J.a[q].bytes;
          }
          s = (s + 1 == ns) ? 0 : s + 1;
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc_w = idesc_bf16(128, J.a_ch + 16, 1, 1);
        int s = 0;
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&bar_full[s], (ph_bits >> s) & 1u);
          ph_bits ^= 1u << s;
          tc_fence_after();
          const uint32_t gs = smem_u32(smem + w.off_stage + s * stride);
          const uint32_t as = gs + g_bytes;
          const uint64_t db0 = smem_desc(as, 128, kChunkBytes);
          const uint64_t da0 = smem_desc(gs, 128, kChunkBytes);
          const uint64_t da1 = smem_desc(gs + 16 * kChunkBytes, 128, kChunkBytes);
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k) {        // 16 points per MMA: +256 B in both images
            const uint32_t accum = (t > t0 || k > 0) ? 1u : 0u;
            mma_ss(tmem, da0 + (uint64_t)(k * 16), db0 + (uint64_t)(k * 16), idesc_w, accum);
            if (n_mblk == 2) mma_ss(tmem + 256, da1 + (uint64_t)(k * 16), db0 + (uint64_t)(k * 16), idesc_w, accum);
          }
          mma_commit(&bar_empty[s]);
          s = (s + 1 == ns) ? 0 : s + 1;
        }
        mma_commit(&bar_done);
      }
    }
    // ---- flush this piece: TMEM -> atomicAdd into the flat fp32 gradient ------------------------
    if (warp >= kEpiWarp0) {
      mbar_wait(&bar_done, piece & 1);
      tc_fence_after();
      const int q = warp & 3;
      const PackedDims dims = packed_dims(J.pl);
      for (int mb = 0; mb < n_mblk; ++mb) {
        const int n = mb * 128 + q * 32 + lane;          // output channel of this thread
        const uint32_t taddr = tmem + mb * 256 + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < J.a_ch + 16; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (n < dims.N) {
            if (c0 < J.a_ch) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int64_t idx = packed_weight_index(w.ps, J.pl, n, c0 + e);
                if (idx >= 0) atomicAdd(w.d_flat + idx, __uint_as_float(v[e]));
              }
            } else if (!J.no_bias) {
              const int64_t idx = packed_bias_index(w.ps, J.pl, n);
              if (idx >= 0) atomicAdd(w.d_flat + idx, __uint_as_float(v[0]));
            }
          }
        }
      }
      tc_fence_before();
    }
    ++piece;
    __syncthreads();      // accumulators are reused by the next piece
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// =================================================================================================
// SIMT helpers
// =================================================================================================
// flat fp32 parameters -> bf16 W / WT images + padded fp32 biases.  One flat index space over all layers; an item is one
// 16-byte chunk of an image (8 consecutive k of W [K/8][N][8], 8 consecutive n of WT [N/8][K][8]) or one bias, ordered so that
// consecutive threads write consecutive chunks.  (Per layer, per element and with 2-byte stores this took 21 us per call.)
__global__ void prepack_kernel(const float* __restrict__ P, PackSrc ps, PackedArena ar, uint8_t* __restrict__ arena, int fine) {
  __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(arena);
  float* bb = reinterpret_cast<float*>(arena + round_up(ar.n_bf16 * 2, 256));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t total = 0;
  for (int pl = 0; pl < PL_COUNT; ++pl) {
    const PackedDims d = packed_dims(pl);
    total += (int64_t)d.N * d.K / 4 + d.N;
  }
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int64_t i = t;
    int pl = 0;
    PackedDims d = packed_dims(0);
    for (; pl < PL_COUNT; ++pl) {
      d = packed_dims(pl);
      const int64_t n_items = (int64_t)d.N * d.K / 4 + d.N;
      if (i < n_items) break;
      i -= n_items;
    }
    const bool fine_only = (pl == PL_DT || pl == PL_TE1 || pl == PL_TE2 || pl == PL_TH);
    if (fine_only && !fine) continue;
    const int64_t n_chunks = (int64_t)d.N * d.K / 8;
    float v[8];
    if (i < n_chunks) {                                         // W image: chunk (k8, n)
      const int k8 = (int)(i / d.N), n = (int)(i % d.N);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t idx = packed_weight_index(ps, pl, n, 8 * k8 + e);
        v[e] = idx >= 0 ? P[idx] : 0.f;
      }
      *reinterpret_cast<uint4*>(wb + ar.w_off[pl] + i * 8) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    } else if (i < 2 * n_chunks) {                              // WT image: chunk (n8, k)
      const int64_t j = i - n_chunks;
      const int n8 = (int)(j / d.K), k = (int)(j % d.K);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t idx = packed_weight_index(ps, pl, 8 * n8 + e, k);
        v[e] = idx >= 0 ? P[idx] : 0.f;
      }
      *reinterpret_cast<uint4*>(wb + ar.wt_off[pl] + j * 8) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    } else {
      const int n = (int)(i - 2 * n_chunks);
      const int64_t idx = packed_bias_index(ps, pl, n);
      bb[ar.bias_off[pl] + n] = idx >= 0 ? P[idx] : 0.f;
    }
  }
}

// positional encoding of one point -> bf16 chunks: [x, sin(2^l x), cos(2^l x)] with the base sin/cos from the accurate
// sincosf and the octaves by the double-angle recurrence (error doubles per octave: <= 2^9 * 1e-7 = 5e-5, far below
// the bf16 operand rounding of 4e-3).   script/models/nerfh_nff.py:241-270
template <int FREQS, int CHUNKS>
__device__ __forceinline__ void pe_row(const float* __restrict__ p3, bool ok, uint8_t* __restrict__ g_row, float last_ch) {
  float e[CHUNKS * 8];
#pragma unroll
  for (int i = 0; i < CHUNKS * 8; ++i) e[i] = 0.f;
  e[CHUNKS * 8 - 1] = last_ch;       // padding channel (its weights are zero); 1 makes it a bias-gradient carrier
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = ok ? p3[c] : 0.f;
    e[c] = v;
    float sn, cs;
    sincosf(v, &sn, &cs);
#pragma unroll
    for (int l = 0; l < FREQS; ++l) {
      e[3 + 6 * l + c] = ok ? sn : 0.f;
      e[6 + 6 * l + c] = ok ? cs : 0.f;
      const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn;
      sn = s2; cs = c2;
    }
  }
#pragma unroll
  for (int j = 0; j < CHUNKS; ++j) {
    uint4 pk;
    pk.x = pack_bf16(e[8 * j], e[8 * j + 1]); pk.y = pack_bf16(e[8 * j + 2], e[8 * j + 3]);
    pk.z = pack_bf16(e[8 * j + 4], e[8 * j + 5]); pk.w = pack_bf16(e[8 * j + 6], e[8 * j + 7]);
    *reinterpret_cast<uint4*>(g_row + j * kChunkBytes) = pk;
  }
}

// pts [M,3] -> X image (64 ch: PE 63 + 0); dirs [N,3] -> DIRPE image (32 ch: PE 27 + 0 x5), one thread per point.
// The images are the first operands of the forward chain AND operands of the weight-gradient kernel.
__global__ void encode_images_kernel(const float* __restrict__ pts, const float* __restrict__ dirs, int S, int64_t M,
                                     int64_t Mp, uint8_t* __restrict__ ximg, uint8_t* __restrict__ dimg) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mp) return;
  const int64_t tile = m / kTile;
  const int row = (int)(m % kTile);
  const bool ok = m < M;
  pe_row<kXyzFreqs, 8>(pts + (ok ? m : 0) * 3, ok, ximg + tile * (64 * 256) + row * 16, 0.f);
  // channel 31 of the direction image = 1: its column of the [dir | tenc0] weight gradient is that layer's bias gradient
  if (dimg != nullptr) pe_row<kDirFreqs, 4>(dirs + (ok ? m / S : 0) * 3, ok, dimg + tile * (32 * 256) + row * 16, 1.f);
}

// tile-major raw blocks [T][C][128] -> row-major [M][C]   (public row-major API of the field query).
// One CTA per tile: coalesced read channel by channel into a row-major shared-memory copy of the tile (pitch C is odd or
// 4 mod 32 -> conflict-free enough), then the tile's rows, contiguous in the output, leave fully coalesced.
__global__ void tiles_to_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t M, int C) {
  extern __shared__ float t_rows[];                                    // [128][C]
  const int64_t tile = blockIdx.x;
  const int64_t n_valid = min((int64_t)kTile, M - tile * kTile);
  const float* s = src + tile * C * kTile;
  float* d = dst + tile * kTile * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = warp; c < C; c += nw)
    for (int r = lane; r < kTile; r += 32) t_rows[r * C + c] = s[(int64_t)c * kTile + r];
  __syncthreads();
  for (int i = threadIdx.x; i < (int)n_valid * C; i += blockDim.x) d[i] = t_rows[i];
}

// d_raw (fp32) -> gradient images of the head pre-activations.  Element (row r of tile t, column c) of raw / d_raw sits
// at  t*C*128 + r*rs + c*cs : row-major rs = C, cs = 1; tile-major rs = 1, cs = 128 (then every load is coalesced).
//   C == 137: GRGB (144 ch), GTH (16 ch), sigma grad into channel 128 of GFS (chunks 16,17)
//   C == 132: GRGB, GFS chunks 16,17            C == 1: GSIG (16 ch)
__global__ void head_grad_images_kernel(const float* __restrict__ raw, const float* __restrict__ d_raw, int C, int rs, int cs,
                                        int64_t M, int64_t Mp, uint8_t* __restrict__ grgb, uint8_t* __restrict__ gth,
                                        uint8_t* __restrict__ gfs, int64_t gfs_tile_stride, int gfs_chunk0,
                                        float* __restrict__ d_sig_bias, const float* __restrict__ compact,
                                        const float* __restrict__ g_rgb, const float* __restrict__ g_feat, int S) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mp) return;                               // Mp is a multiple of the block size: whole blocks leave together
  const int64_t tile = m / kTile;
  const int row = (int)(m % kTile);
  const bool ok = m < M;
  const int64_t base = ok ? tile * C * kTile + (int64_t)row * rs : 0;
  const float* y = raw + base;
  const float* gd = d_raw + base;
  const int sig_col = (C == 1) ? 0 : 131;
  // compact cotangent (nefes_composite_bwd_compact): d_raw[m, c] = w_s[m] * g_ray[c] (c < 131), w_t[m] * g_rgb (132..134),
  // and three per-sample scalars -- rebuilt here instead of being read back from a 137-channel fp32 block
  float w_s = 0.f, w_t = 0.f, c_dsig = 0.f, c_dsigt = 0.f, c_dbeta = 0.f;
  const float* gf = nullptr;
  float grgb3[3] = {0.f, 0.f, 0.f};
  if (compact != nullptr && ok) {
    const int64_t ray = m / S;
    const float* cr = compact + ray * 5 * S + (m % S);
    w_s = cr[0]; w_t = cr[S]; c_dsig = cr[2 * S]; c_dsigt = cr[3 * S]; c_dbeta = cr[4 * S];
    if (g_feat != nullptr) gf = g_feat + ray * kFeat;
    if (g_rgb != nullptr) { grgb3[0] = g_rgb[ray * 3]; grgb3[1] = g_rgb[ray * 3 + 1]; grgb3[2] = g_rgb[ray * 3 + 2]; }
  }
  auto gval = [&](int c) -> float {              // cotangent of raw[m, c]
    if (compact == nullptr) return gd[(int64_t)c * cs];
    if (c < 3) return grgb3[c] * w_s;
    if (c < kHeadCh) return gf ? gf[c - 3] * w_s : 0.f;
    if (c == 131) return c_dsig;
    if (c < 135) return grgb3[c - 132] * w_t;
    return c == 135 ? c_dsigt : c_dbeta;
  };
  const float dsig = ok ? gval(sig_col) * (1.f - expf(-y[(int64_t)sig_col * cs])) : 0.f;
  if (d_sig_bias != nullptr) {                       // bias gradient of the sigma head = sum of its pre-activation gradients
    __shared__ float s_part[4];
    const float ws = warp_sum(__bfloat162float(__float2bfloat16(dsig)));     // the value the tensor path sees
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ws;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(d_sig_bias, s_part[0] + s_part[1] + s_part[2] + s_part[3]);
  }
  {  // sigma pre-activation gradient: channel 0 of a 16-channel group, rest zero
    uint8_t* base_s = gfs + tile * gfs_tile_stride + (int64_t)gfs_chunk0 * kChunkBytes + row * 16;
    uint4 pk = make_uint4(pack_bf16(dsig, 0.f), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(base_s) = pk;
    *reinterpret_cast<uint4*>(base_s + kChunkBytes) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (C == 1) return;
  uint8_t* rb = grgb + tile * (144 * 256) + row * 16;
  for (int j = 0; j < 18; ++j) {
    float e[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = 8 * j + q;
      e[q] = (ok && c < kHeadCh) ? gval(c) : 0.f;
    }
    uint4 pk;
    pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
    *reinterpret_cast<uint4*>(rb + j * kChunkBytes) = pk;
  }
  if (C == 137) {
    float e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { const float yc = y[(int64_t)(132 + c) * cs]; e[c] = gval(132 + c) * yc * (1.f - yc); }
      e[3] = gval(135) * (1.f - expf(-y[(int64_t)135 * cs]));
      e[4] = gval(136) * (1.f - expf(-y[(int64_t)136 * cs]));
    }
    uint8_t* tb = gth + tile * (16 * 256) + row * 16;
    uint4 pk;
    pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
    *reinterpret_cast<uint4*>(tb) = pk;
    *reinterpret_cast<uint4*>(tb + kChunkBytes) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// out[n, c] = sum_s src[(n*S + s)*ld + c], c < ncol (per-ray sum of a per-point fp32 gradient)
__global__ void ray_reduce_kernel(const float* __restrict__ src, int ld, int ncol, int S, int64_t N, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * ncol) return;
  const int64_t n = idx / ncol;
  const int c = (int)(idx % ncol);
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += src[(n * S + s) * ld + c];
  out[idx] = a;
}

}  // namespace nefes
#include "mlp_chain.cuh"
#include "mlp_chain_ts2.cuh"
#include "mlp_trunk_bwd.cuh"
#include "mlp_fused_bwd.cuh"
extern "C" int nefes_encode_pe_bwd(const float*, const float*, int, int64_t, int, float*, void*);
namespace nefes {

// =================================================================================================
// host orchestration
// =================================================================================================
namespace {

constexpr uint32_t kSmemBudget = 216 * 1024;
constexpr int kSmemAttr = 224 * 1024;         // opt-in dynamic shared memory per CTA (227 KB max incl. static)
constexpr uint32_t kMinSmem = 120 * 1024;     // keep one CTA per SM (each CTA allocates all 512 TMEM columns)

inline uint32_t r1k(uint32_t x) { return (x + 1023u) & ~1023u; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// image bookkeeping: all images of a workspace are carved from one allocation
struct Img { uint8_t* p; int ch; int64_t stride = 0; int64_t tile_stride() const { return stride ? stride : (int64_t)ch * 256; } };

struct Ws {             // saved-for-backward (forward writes, backward reads)
  uint8_t* arena;
  Img X, DIRPE, H[8], FIN, DT, T2, T3;
  int64_t bytes;
};
struct WsB {            // backward scratch
  Img GTH, GT3, GT2, GDT, GRGB, GFS, G[8], GSIG;
  float* dX;            // [Mp,64] fp32 gradient w.r.t. the xyz PE      (pose refinement only)
  float* dDIR;          // [Mp,32] fp32 gradient w.r.t. the dir PE, per point
  float* dDIRray;       // [N,32]  ... summed over the samples of a ray
  int64_t bytes;
};

Ws carve_ws(void* base, int64_t T, int mode) {
  Ws w = {};
  uint8_t* p = (uint8_t*)base;
  auto take = [&](int64_t n) { uint8_t* q = p; p += round_up(n, 1024); return q; };
  auto img = [&](int ch) { Img i; i.ch = ch; i.p = take(T * ch * 256); return i; };
  w.arena = take(packed_arena().bytes);
  w.X = img(64);
  if (mode != NEFES_MODE_SIGMA) w.DIRPE = img(32);
  // NEFES_WS_INTERLEAVE=1 (experiment): the saved activation images of a TILE are contiguous ([tile][image][ch * 256 B])
  // instead of one array per image -- same bytes, same total size; every consumer addresses (base, tile stride)
  static const bool interleave = getenv("NEFES_WS_INTERLEAVE") != nullptr;
  int chs[12], n = 0;
  for (int l = 0; l < 8; ++l) chs[n++] = 128;
  if (mode != NEFES_MODE_SIGMA) {
    chs[n++] = 128;                                           // FIN
    chs[n++] = mode == NEFES_MODE_FULL ? 128 : 64;            // DT
    if (mode == NEFES_MODE_FULL) { chs[n++] = 64; chs[n++] = 64; }
  }
  Img im[12];
  if (interleave) {
    int64_t per_tile = 0;
    for (int i = 0; i < n; ++i) per_tile += (int64_t)chs[i] * 256;
    uint8_t* blk = take(T * per_tile);
    int64_t off = 0;
    for (int i = 0; i < n; ++i) { im[i].ch = chs[i]; im[i].p = blk + off; im[i].stride = per_tile; off += (int64_t)chs[i] * 256; }
  } else {
    for (int i = 0; i < n; ++i) im[i] = img(chs[i]);
  }
  for (int l = 0; l < 8; ++l) w.H[l] = im[l];
  if (mode != NEFES_MODE_SIGMA) {
    w.FIN = im[8]; w.DT = im[9];
    if (mode == NEFES_MODE_FULL) { w.T2 = im[10]; w.T3 = im[11]; }
  }
  w.bytes = (int64_t)(p - (uint8_t*)base);
  return w;
}
WsB carve_wsb(void* base, int64_t T, int mode, int64_t N) {
  WsB w = {};
  uint8_t* p = (uint8_t*)base;
  auto img = [&](int ch) { Img i; i.ch = ch; i.p = p; p += round_up(T * ch * 256, 1024); return i; };
  for (int l = 0; l < 8; ++l) w.G[l] = img(128);
  if (mode == NEFES_MODE_SIGMA) {
    w.GSIG = img(16);
  } else {
    w.GFS = img(144); w.GRGB = img(144);
    w.GDT = img(mode == NEFES_MODE_FULL ? 128 : 64);
    if (mode == NEFES_MODE_FULL) { w.GTH = img(16); w.GT3 = img(64); w.GT2 = img(64); }
  }
  w.dX = (float*)p; p += round_up(T * kTile * 64 * 4, 1024);
  w.dDIR = (float*)p; p += round_up(T * kTile * 32 * 4, 1024);
  w.dDIRray = (float*)p; p += round_up(N * 32 * 4, 1024);
  w.bytes = (int64_t)(p - (uint8_t*)base);
  return w;
}

ASrc src_of(const Img& i, int ch0 = 0, int nch = -1) {
  ASrc s;
  s.base = i.p + (int64_t)ch0 * 256;
  s.tile_stride = i.tile_stride();
  s.bytes = (uint32_t)((nch < 0 ? i.ch - ch0 : nch) * 256);
  return s;
}

struct GemmDesc {
  ASrc a[2]; int n_src = 1;
  int K = 0, N = 0;
  const uint8_t* w_img = nullptr; int w_rows = 0, w_row0 = 0;
  const float* bias = nullptr;
  uint8_t* out_img = nullptr; int64_t out_tile_stride = 0; int out_ch = 0; int relu = 0;
  uint4* mask_out = nullptr; const uint4* mask_in = nullptr; int mask_shift = 0;
  float* raw = nullptr; int raw_ld = 0, raw_col0 = 0, d_col0 = 0, raw_ncol = 0, raw_act = 0;
  const float* pe_x = nullptr; int pe_L = 0; int ray_S = 0; int64_t n_rays = 0;     // see TileGemmArgs
};

int launch_tile_gemm(const GemmDesc& d, int n_tiles, int64_t M, cudaStream_t st, const char* what) {
  TileGemmArgs g = {};
  g.a[0] = d.a[0]; g.a[1] = d.a[1]; g.n_src = d.n_src;
  g.K = d.K; g.N = d.N; g.w_img = d.w_img; g.w_rows = d.w_rows; g.w_row0 = d.w_row0; g.bias = d.bias;
  g.n_tiles = n_tiles; g.M = M;
  g.out_img = d.out_img; g.out_tile_stride = d.out_tile_stride; g.out_ch = d.out_img ? d.out_ch : 0; g.relu = d.relu;
  g.mask_out = d.mask_out; g.mask_in = d.mask_in; g.mask_shift = d.mask_shift;
  g.raw = d.raw; g.raw_ld = d.raw_ld; g.raw_col0 = d.raw_col0; g.d_col0 = d.d_col0; g.raw_ncol = d.raw_ncol; g.raw_act = d.raw_act;
  g.pe_x = d.pe_x; g.pe_L = d.pe_L; g.ray_S = d.ray_S; g.n_rays = d.n_rays;
  NEFES_REQUIRE(!d.pe_x || (d.raw_ncol >= 3 + 6 * d.pe_L && d.raw_ncol > 8 && !d.ray_S), NEFES_EINVAL, "%s: bad encoding-backward epilogue", what);
  NEFES_REQUIRE(!d.ray_S || (d.ray_S % 16 == 0 && kTile % d.ray_S == 0 && d.raw_ncol > 8 && d.raw_ncol <= 32), NEFES_EINVAL,
                "%s: bad per-ray reduction epilogue (S=%d)", what, d.ray_S);
  uint32_t src_bytes = 0;
  for (int q = 0; q < d.n_src; ++q) src_bytes += d.a[q].bytes;
  NEFES_REQUIRE(src_bytes == (uint32_t)d.K * 256u, NEFES_EINVAL, "%s: A sources (%u B) do not add up to K=%d", what, src_bytes, d.K);
  NEFES_REQUIRE(d.K % 16 == 0 && d.N % 16 == 0 && d.N <= 256 && d.N >= 16, NEFES_EINVAL, "%s: bad K/N %d/%d", what, d.K, d.N);
  const uint32_t w_bytes = (uint32_t)d.K * d.w_rows * 2u;
  const uint32_t a_stage = r1k((uint32_t)d.K * 256u);
  g.out_bytes = (uint32_t)g.out_ch * 256u;
  const uint32_t out_total = 2 * r1k(g.out_bytes);
  g.raw_pitch = (uint32_t)(d.raw_ncol | 1);
  const uint32_t raw_total = (d.raw && d.raw_ncol > 8) ? r1k(kTile * g.raw_pitch * 4u + (d.ray_S ? 1024u : 0u)) : 0u;   // EPI_RAWBULK staging
  const uint32_t bias_total = 1024;
  const uint32_t fixed = r1k(w_bytes) + out_total + raw_total + bias_total;
  NEFES_REQUIRE(fixed + 2 * a_stage <= kSmemBudget, NEFES_EINVAL, "%s: shared memory budget exceeded", what);
  int stages = (int)((kSmemBudget - fixed) / a_stage);
  if (stages > 4) stages = 4;
  g.off_a = r1k(w_bytes);
  g.a_stage_stride = a_stage;
  g.off_out = g.off_a + stages * a_stage;
  g.off_raw = g.off_out + out_total;
  g.off_bias = g.off_raw + raw_total;
  uint32_t smem = g.off_bias + bias_total;
  if (smem < kMinSmem) smem = kMinSmem;
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  int epi = EPI_IMG;
  if (d.raw && d.raw_ncol > 8) {
    NEFES_REQUIRE(!d.out_img && d.d_col0 == 0 && d.raw_act == RAW_ACT_NONE, NEFES_EINVAL, "%s: unsupported bulk fp32 epilogue", what);
    epi = EPI_RAWBULK;
  } else if (d.out_img && !d.raw && d.out_ch == d.N && d.N <= 128 && d.relu && d.bias && d.mask_out && !d.mask_in) {
    epi = EPI_HIDDEN;
  } else if (d.out_img && !d.raw && d.out_ch == d.N && d.N <= 128 && !d.relu && !d.bias && !d.mask_out) {
    epi = EPI_DGRAD;
  } else {
    NEFES_REQUIRE(!d.relu && !d.mask_out && !d.mask_in && (!d.raw || (d.raw_ncol <= 5 && d.d_col0 % 16 == 0)), NEFES_EINVAL,
                  "%s: unsupported epilogue combination", what);
  }
  static bool attr_done = false;
  if (!attr_done) {
#define NEFES_SET(ST, EP) NEFES_CUDA(cudaFuncSetAttribute(tile_gemm_kernel<ST, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAttr))
#define NEFES_SET3(EP) NEFES_SET(2, EP); NEFES_SET(3, EP); NEFES_SET(4, EP)
    NEFES_SET3(EPI_HIDDEN); NEFES_SET3(EPI_DGRAD); NEFES_SET3(EPI_IMG); NEFES_SET3(EPI_RAWBULK);
#undef NEFES_SET3
#undef NEFES_SET
    attr_done = true;
  }
#define NEFES_GO(ST, EP) tile_gemm_kernel<ST, EP><<<grid, kGemmThreads, smem, st>>>(g)
#define NEFES_GO_EPI(ST) do { if (epi == EPI_HIDDEN) NEFES_GO(ST, EPI_HIDDEN); else if (epi == EPI_DGRAD) NEFES_GO(ST, EPI_DGRAD); \
                              else if (epi == EPI_RAWBULK) NEFES_GO(ST, EPI_RAWBULK); else NEFES_GO(ST, EPI_IMG); } while (0)
  switch (stages) {
    case 2: NEFES_GO_EPI(2); break;
    case 3: NEFES_GO_EPI(3); break;
    default: NEFES_GO_EPI(4); break;
  }
#undef NEFES_GO_EPI
#undef NEFES_GO
  NEFES_CHECK_LAUNCH(what);
  return NEFES_OK;
}

PackSrc pack_src(int net) {
  const Layout& L = layout_for(net);
  PackSrc s;
  for (int i = 0; i < NEFES_MAX_LAYERS; ++i) { s.w[i] = L.w[i]; s.b[i] = L.b[i]; }
  s.fine = net == NEFES_NET_FINE;
  return s;
}

struct Arena {
  const uint8_t* base; PackedArena ar;
  const uint8_t* W(int pl) const { return base + ar.w_off[pl] * 2; }
  const uint8_t* WT(int pl) const { return base + ar.wt_off[pl] * 2; }
  const float* bias(int pl) const { return reinterpret_cast<const float*>(base + round_up(ar.n_bf16 * 2, 256)) + ar.bias_off[pl]; }
};

#define TRY(x) do { if (int e__ = (x)) return e__; } while (0)

}  // namespace

namespace {
long long* chain_dbg_buf() {
  static long long* d = nullptr;
  static int on = -1;
  if (on < 0) { const char* e = getenv("NEFES_CHAIN_DBG"); on = (e && e[0] == '1') ? 1 : 0; }
  if (on && d == nullptr) { cudaMalloc(&d, 2048 * sizeof(long long)); }
  return on ? d : nullptr;
}
// Timing experiments of the chain kernels (NEFES_CHAIN_X bit mask: drop the saved copies, the weight traffic, the raw
// stores ...).  They change RESULTS, so the mask is honoured only together with NEFES_UNSAFE_EXPERIMENTS=1.
int chain_xflags() {
  static const int flags = [] {
    const char* u = getenv("NEFES_UNSAFE_EXPERIMENTS");
    const char* e = getenv("NEFES_CHAIN_X");
    return (u && u[0] == '1' && e) ? atoi(e) : 0;
  }();
  return flags;
}

void chain_dbg_dump(const char* what, const ChainArgs& c, cudaStream_t st) {
  if (c.dbg == nullptr) return;
  static int dumps = 0;
  if (dumps >= 6) return;
  ++dumps;
  cudaStreamSynchronize(st);
  static long long h[2048];
  cudaMemcpy(h, c.dbg, sizeof(h), cudaMemcpyDeviceToHost);
  const long long t0 = h[0];
  fprintf(stderr, "[chain dbg] %s: n_steps=%d tiles=%d  (cycles since the first weight tile landed)\n", what, c.n_steps, c.n_tiles);
  for (int i = 0; i < 2 * c.n_steps && i < 32; ++i) {
    const long long* r = h + i * 48;
    long long lo[2] = {1ll << 62, 1ll << 62}, hi[2] = {0, 0}, rd[2] = {0, 0};
    for (int w = 0; w < 16; ++w) {
      lo[w >> 3] = r[24 + w] < lo[w >> 3] ? r[24 + w] : lo[w >> 3];
      hi[w >> 3] = r[24 + w] > hi[w >> 3] ? r[24 + w] : hi[w >> 3];
      rd[w >> 3] = r[8 + w] > rd[w >> 3] ? r[8 + w] : rd[w >> 3];
    }
    fprintf(stderr, "  [epi warp 0: ready %lld | tmem loaded +%lld | computed+stored +%lld | fenced +%lld | group barrier +%lld]\n", r[8] - t0,
            h[1600 + i * 4] - r[8], h[1600 + i * 4 + 1] - r[8], h[1600 + i * 4 + 2] - r[8], h[1600 + i * 4 + 3] - r[8]);
    fprintf(stderr, "  [issuer 0: waits passed %lld fence +%lld acc commit +%lld wempty commit +%lld | issuer 1: waits passed %lld fence +%lld acc commit +%lld wempty commit +%lld]\n",
            r[40] - t0, r[1] - r[40], r[3] - r[40], r[5] - r[40], r[41] - t0, r[2] - r[41], r[4] - r[41], r[6] - r[41]);
    fprintf(stderr, "  seq %2d step %2d | W ok %7lld | t0: A ok %7lld commit %7lld | t1: A ok %7lld commit %7lld | epi0 ready %7lld done %7lld..%7lld | epi1 ready %7lld done %7lld..%7lld\n",
            i, i % c.n_steps, r[0] - t0, r[1] - t0, r[3] - t0, r[2] - t0, r[4] - t0, rd[0] - t0, lo[0] - t0, hi[0] - t0,
            rd[1] - t0, lo[1] - t0, hi[1] - t0);
  }
}

// weight image streamed in pieces: the whole image, or rows [row0, row0 + 128) of every 8-wide chunk (compacted)
void set_weights(ChainStep& s, const uint8_t* img, int chunks, int rows, int row0, int take_rows) {
  s.w_img = img + (int64_t)row0 * 16;
  s.w_lbo = (uint32_t)take_rows * 16u;
  s.w_bytes = (uint32_t)chunks * take_rows * 16u;
  if (take_rows == rows) { s.w_piece = 16384u; s.w_src_stride = 16384u; }
  else { s.w_piece = (uint32_t)take_rows * 16u; s.w_src_stride = (uint32_t)rows * 16u; }
}

// Build the step table of the fused forward chain for (net, mode) and launch it.  raw_t: tile-major [T][C][128].
int launch_chain_fwd(const Ws& w, const Arena& A, int mode, int64_t M, float* raw_t, cudaStream_t st) {
  const int T = (int)ceil_div(M, kTile);
  ChainArgs c = {};
  int n = 0, bias_floats = 0;
  enum { LD_X = 0, LD_D = 1 };
  // one layer = one step, or two when its weight image exceeds a ring slot: K-slice [0, k_split) without epilogue
  // (CK_NONE), then [k_split, K) accumulating on top.  a_off2: operand offset of the second slice.
  auto add = [&](int pl, uint32_t a_off, int kind, uint32_t out_off, int out_ch, const Img* save, int wait_load,
                 int k_split = 0, uint32_t a_off2 = 0, int wait_load_b = -1) {
    const PackedDims pd = packed_dims(pl);
    const int boff = bias_floats;
    bias_floats += (pd.N + 3) & ~3;
    for (int part = 0; part < (k_split ? 2 : 1); ++part) {
      ChainStep& s = c.step[n++];
      const int k0 = part == 0 ? 0 : k_split, k1 = (k_split && part == 0) ? k_split : pd.K;
      s.a_off = part == 0 ? a_off : a_off2; s.out_off = out_off; s.K = (uint16_t)(k1 - k0); s.N = (uint16_t)pd.N;
      s.out_ch = (uint16_t)out_ch; s.wait_load2 = -1; s.acc0 = (int8_t)part;
      const bool last = !k_split || part == 1;
      s.kind = (uint8_t)(last ? kind : CK_NONE);
      s.wait_load = (int8_t)(part == 0 ? wait_load : wait_load_b);
      s.bias = last ? A.bias(pl) : nullptr; s.bias_off = (uint16_t)boff;
      set_weights(s, A.W(pl) + (int64_t)(k0 / 8) * pd.N * 16, (k1 - k0) / 8, pd.N, 0, pd.N);
      const bool keep = save && last && !forward_only();     // no saved copies when nobody will run the backward
      s.gdst = keep ? save->p : nullptr; s.g_tile_stride = keep ? (uint32_t)save->tile_stride() : 0u;
    }
  };
  auto load = [&](int idx, const Img& img, uint32_t dst_off, int issue_step) {
    ChainLoad& L = c.load[idx];
    L.src = img.p; L.tile_stride = (uint32_t)img.tile_stride(); L.bytes = (uint32_t)img.ch * 256u;
    L.dst_off = dst_off; L.issue_step = (int8_t)issue_step; L.next_pair = 1;
  };
  for (int l = 0; l < kChainLoads; ++l) c.load[l].issue_step = -1;
  add(PL_T0, kRegX, CK_HIDDEN, kRegH, 128, &w.H[0], LD_X);
  for (int l = 1; l < 8; ++l) {
    const uint32_t kbytes = (uint32_t)packed_dims(PL_T0 + l).K * packed_dims(PL_T0 + l).N * 2u;
    if (l == 4 && kbytes > kFwdWSlot) add(PL_T4, kRegX, CK_HIDDEN, kRegH, 128, &w.H[4], -1, 64, kRegH);   // xyz slice, then h slice
    else add(PL_T0 + l, l == 4 ? kRegX : kRegH, CK_HIDDEN, kRegH, 128, &w.H[l], -1);
  }
  const int x_dead = n;                                            // first step after the skip layer retired
  if (mode == NEFES_MODE_SIGMA) {
    add(PL_SIG, kRegH, CK_SIGMA, 0, 0, nullptr, -1);
    load(LD_X, w.X, kRegX, x_dead);
  } else {
    add(PL_FS, kRegH, CK_FS, kRegH, 128, &w.FIN, -1);
    if (mode == NEFES_MODE_FULL) {
      if ((uint32_t)packed_dims(PL_DT).K * packed_dims(PL_DT).N * 2u > kFwdWSlot)
        add(PL_DT, kRegH, CK_HIDDEN, kRegH, 128, &w.DT, -1, 128, kRegD, LD_D);   // [final | dirPE]: final slice, then dirPE slice
      else
        add(PL_DT, kRegH, CK_HIDDEN, kRegH, 128, &w.DT, LD_D);
      load(LD_D, w.DIRPE, kRegD, n);
      add(PL_TE1, kRegH + 16384, CK_HIDDEN, kRegX, 64, &w.T2, -1);        // t1 -> t2 (parked in the xyzPE slot)
      add(PL_TE2, kRegX, CK_HIDDEN, kRegH + 16384, 64, &w.T3, -1);        // t2 -> t3 (over t1)
      add(PL_TH, kRegH + 16384, CK_HEADS, 0, 0, nullptr, -1);
      // the xyz slot held t2: free once the MMAs of the heads step retired (its epilogue waited for t2's bulk store)
      load(LD_X, w.X, kRegX, n);
    } else {
      add(PL_DIR, kRegH, CK_HIDDEN, kRegH, 64, &w.DT, LD_D);
      load(LD_X, w.X, kRegX, x_dead);
      load(LD_D, w.DIRPE, kRegD, n);
    }
    add(PL_RGB, kRegH, CK_RGB, 0, 0, nullptr, -1);
  }
  NEFES_REQUIRE(n <= kChainMaxSteps && bias_floats * 4 <= (int)kChainBiasBytes, NEFES_EINVAL, "chain_fwd: step table overflow");
  for (int i = 0; i < n; ++i)
    NEFES_REQUIRE(c.step[i].w_bytes <= kFwdWSlot, NEFES_EINVAL, "chain_fwd: weight slice of step %d exceeds the ring slot", i);
  c.n_steps = n; c.n_loads = kChainLoads;
  c.M = M; c.n_tiles = T;
  c.raw = raw_t; c.C = (mode == NEFES_MODE_SIGMA) ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137);
  c.dbg = chain_dbg_buf();
  c.xflags = chain_xflags();
  static bool attr_done = false;
  if (!attr_done) {
    NEFES_CUDA(cudaFuncSetAttribute(chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdChainSmem));
    attr_done = true;
  }
  const int n_pairs = (T + 1) / 2;
  const int grid = n_pairs < num_sms() ? n_pairs : num_sms();
  {
    // algorithmic HBM bytes per point: encodings in, every saved activation image out (bf16), raw out (fp32); flops: 2 x MACs
    const double save_ch = forward_only() ? 0 : ((mode == NEFES_MODE_SIGMA) ? 8 * 128 : (mode == NEFES_MODE_STATIC ? 8 * 128 + 128 + 64 : 8 * 128 + 128 + 128 + 64 + 64));
    const double in_ch = (mode == NEFES_MODE_SIGMA) ? 64 : 96;
    const double macs = (mode == NEFES_MODE_SIGMA) ? 130944 : (mode == NEFES_MODE_STATIC ? 165632 : 184064);
    prof_begin(mode == NEFES_MODE_FULL ? "chain_fwd_fine" : (mode == NEFES_MODE_STATIC ? "chain_fwd_coarse" : "chain_fwd_sigma"), st,
               (double)M * (2.0 * (save_ch + in_ch) + 4.0 * c.C), (double)M * 2.0 * macs);
  }
  chain_kernel<false><<<grid, kChainThreads, kFwdChainSmem, st>>>(c);
  prof_end(st);
  NEFES_CHECK_LAUNCH("chain_fwd");
  chain_dbg_dump("fwd", c, st);
  return NEFES_OK;
}

// clock stamps of CTA 0 of the TMEM-operand chains (NEFES_CHAIN_DBG=1): stride 48 per sequence number
void chain_ts_dbg_dump(long long* dbg, int n_steps, cudaStream_t st) {
  if (dbg == nullptr) return;
  static int dumps = 0;
  if (dumps++ >= 2) return;
  cudaStreamSynchronize(st);
  static long long h[2048];
  cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemset(dbg, 0, sizeof(h));
  const long long t0 = h[0];
  fprintf(stderr, "[chain_ts dbg] n_steps=%d  (cycles since tile 0's first issue)\n", n_steps);
  for (int i = 0; i < 32 && i < 2 * n_steps; ++i) {
    const long long* r = h + i * 48;
    long long rd[2] = {0, 0}, lo[2] = {1ll << 62, 1ll << 62}, hi[2] = {0, 0};
    for (int w2 = 0; w2 < 16; ++w2) {
      rd[w2 >> 3] = r[8 + w2] > rd[w2 >> 3] ? r[8 + w2] : rd[w2 >> 3];
      lo[w2 >> 3] = r[24 + w2] < lo[w2 >> 3] ? r[24 + w2] : lo[w2 >> 3];
      hi[w2 >> 3] = r[24 + w2] > hi[w2 >> 3] ? r[24 + w2] : hi[w2 >> 3];
    }
    fprintf(stderr, "  seq %2d step %2d | prep %7lld %7lld | t0: K %7lld W %7lld A %7lld issued %7lld  epi0 ready %7lld done %7lld..%7lld epi1 %7lld..%7lld | t1: K %7lld W %7lld A %7lld issued %7lld  epi0 ready %7lld done %7lld..%7lld epi1 %7lld..%7lld\n",
            i, i % n_steps, r[4] - t0, r[5] - t0, r[46] - t0, r[40] - t0, r[0] - t0, r[1] - t0, rd[0] - t0, lo[0] - t0, hi[0] - t0, r[42] - t0, r[44] - t0,
            r[47] - t0, r[41] - t0, r[2] - t0, r[3] - t0, rd[1] - t0, lo[1] - t0, hi[1] - t0, r[43] - t0, r[45] - t0);
    if (i < 8)
      for (int g = 0; g < 2; ++g)
        for (int rr = 0; rr < 4; ++rr) {
          const long long* q = h + 1600 + ((i * 4 + rr) * 2 + g) * 7;
          if (q[0] == 0) continue;
          fprintf(stderr, "      issuer t%d run %d: top %lld | record +%lld | waits +%lld | weights +%lld | fence +%lld | MMAs issued +%lld | commit +%lld\n",
                  g, rr, q[0] - t0, q[1] - q[0], q[2] - q[0], q[3] - q[0], q[4] - q[0], q[5] - q[0], q[6] - q[0]);
        }
  }
}

// Round-2 forward chain (mlp_chain_ts2.cuh): N-split layers, double-buffered activations in tensor memory.
int launch_chain_fwd_ts2(const Ws& w, const Arena& A, int mode, int64_t M, float* raw_t, cudaStream_t st) {
  const int T = (int)ceil_div(M, kTile);
  Ts2Args c = {};
  int n = 0, bias_floats = 0;
  const bool keep_all = !forward_only();
  auto begin = [&](int pl, const Img* save) -> Ts2Step& {
    Ts2Step& s = c.step[n++];
    const PackedDims pd = packed_dims(pl);
    s.w_img = A.W(pl); s.w_rows = (uint32_t)pd.N; s.w_bytes = (uint32_t)pd.K * pd.N * 2u;
    s.bias = A.bias(pl); s.bias_off = (uint16_t)bias_floats;
    bias_floats += (pd.N + 3) & ~3;
    const bool keep = save && keep_all;
    s.gdst = keep ? save->p : nullptr; s.g_tile_stride = keep ? (uint32_t)save->tile_stride() : 0u;
    s.save_bytes = keep ? (uint32_t)save->ch * 256u : 0u;
    return s;
  };
  auto grp = [&](Ts2Step& s, int src, int wait, int col, int ksteps, int k0) {
    Ts2Group& g = s.grp[s.n_grp++];
    g.src = (uint8_t)src; g.wait = (uint8_t)wait; g.col = (uint16_t)col; g.ksteps = (uint16_t)ksteps; g.k0 = (uint16_t)k0;
  };
  auto blk = [&](Ts2Step& s, int kind, int n0, int nw, int acc_col, int wait, int out_col, int save, int raw_c0 = 0, int raw_n = 0) {
    Ts2Block& b = s.blk[s.n_blk++];
    b.kind = (uint8_t)kind; b.n0 = (uint16_t)n0; b.nw = (uint16_t)nw; b.acc_col = (uint16_t)acc_col; b.wait = (uint8_t)wait;
    b.out_col = (uint16_t)out_col; b.save = (uint8_t)save; b.raw_c0 = (uint16_t)raw_c0; b.raw_n = (uint16_t)raw_n;
  };
  // a 128-wide hidden layer: two 64-column blocks, reading H buffer `in` (both halves), writing buffer `out`
  auto hidden128 = [&](Ts2Step& s, int kind, uint32_t out, int wait0, int wait1) {
    blk(s, kind, 0, 64, kTs2Acc, wait0, out, 1);
    blk(s, kind, 64, 64, kTs2Acc + 64, wait1, out + 32, 2);
  };
  auto h_groups = [&](Ts2Step& s, uint32_t in, int k0) {
    grp(s, GS_TMEM, W_K0, in, 4, k0);
    grp(s, GS_TMEM, W_K1, in + 32, 4, k0 + 4);
  };
  uint32_t hin = kTs2HB, hout = kTs2HA;               // layer l reads `hin`, writes `hout`; swapped after every 128-wide layer
  {                                                    // T0: xyz encoding (shared memory) -> HA.  A new tile: both accumulator halves
    Ts2Step& s = begin(PL_T0, &w.H[0]);                // must have been drained by the previous tile's last users
    grp(s, GS_X, W_X, 0, 4, 0);
    hidden128(s, BK_HID_RELU, hout, W_K0, W_K1);
  }
  for (int l = 1; l < 8; ++l) {
    uint32_t t = hin; hin = hout; hout = t;
    Ts2Step& s = begin(PL_T0 + l, &w.H[l]);
    if (l == 4) { grp(s, GS_X, W_X, 0, 4, 0); h_groups(s, hin, 4); }        // skip layer: [xyzPE | h4]
    else h_groups(s, hin, 0);
    hidden128(s, BK_HID_RELU, hout, 0, 0);
  }
  { uint32_t t = hin; hin = hout; hout = t; }          // hin = h8
  int x_last = 4, d_last = -1;
  if (mode == NEFES_MODE_SIGMA) {
    Ts2Step& s = begin(PL_SIG, nullptr);
    h_groups(s, hin, 0);
    blk(s, BK_SIGMA, 0, 16, kTs2Acc, 0, 0, 0, 0, 1);
  } else {
    {                                                  // final (128, no activation) -> hout; sigma (row 128) as a third block once
      Ts2Step& s = begin(PL_FS, &w.FIN);               // block 0's accumulator columns are drained (W_K0 then names THIS step's)
      h_groups(s, hin, 0);
      hidden128(s, BK_HID, hout, 0, 0);
      blk(s, BK_SIGMA, 128, 16, kTs2Acc, W_K0, 0, 0, 131, 1);
    }
    { uint32_t t = hin; hin = hout; hout = t; }        // hin = final, hout = the buffer that held h8
    if (mode == NEFES_MODE_FULL) {
      {                                                // [final | dirPE] -> [dir hidden | t1]
        Ts2Step& s = begin(PL_DT, &w.DT);
        h_groups(s, hin, 0);
        grp(s, GS_D, W_D, 0, 2, 8);
        blk(s, BK_HID_RELU, 0, 64, kTs2Acc, W_K2, hout, 1);
        blk(s, BK_HID_RELU, 64, 64, kTs2Acc + 64, 0, hout + 32, 2);
        d_last = n - 1;
      }
      {                                                // t1 (hout[32,64)) -> t2, parked in hin[32,64) (final is dead)
        Ts2Step& s = begin(PL_TE1, &w.T2);
        grp(s, GS_TMEM, W_K1, hout + 32, 4, 0);
        blk(s, BK_HID_RELU, 0, 64, kTs2Acc, W_K0, hin + 32, 2);
      }
      {                                                // dir hidden (hout[0,32)) -> 131 raw channels; columns 128..143 spill into HA[0,16)
        Ts2Step& s = begin(PL_RGB, nullptr);
        NEFES_REQUIRE(hin == kTs2HA, NEFES_EINVAL, "chain_fwd_ts2: the colour head's spill columns must be the dead buffer");
        grp(s, GS_TMEM, W_K0, hout, 4, 0);
        blk(s, BK_RAW, 0, 64, kTs2Acc, 0, 0, 0, 0, 64);
        blk(s, BK_RAW, 64, 80, kTs2Acc + 64, W_K1, 0, 0, 64, 67);
      }
      {                                                // t2 -> t3 over the dir hidden
        Ts2Step& s = begin(PL_TE2, &w.T3);
        grp(s, GS_TMEM, W_K0, hin + 32, 4, 0);
        blk(s, BK_HID_RELU, 0, 64, kTs2Acc, 0, hout, 2);
      }
      {
        Ts2Step& s = begin(PL_TH, nullptr);
        grp(s, GS_TMEM, W_K0, hout, 4, 0);
        blk(s, BK_HEADS, 0, 16, kTs2Acc, 0, 0, 0, 132, 5);
      }
    } else {
      {
        Ts2Step& s = begin(PL_DIR, &w.DT);
        h_groups(s, hin, 0);
        grp(s, GS_D, W_D, 0, 2, 8);
        blk(s, BK_HID_RELU, 0, 64, kTs2Acc, W_K2, hout, 2);
        d_last = n - 1;
      }
      {
        Ts2Step& s = begin(PL_RGB, nullptr);
        NEFES_REQUIRE(hin == kTs2HA, NEFES_EINVAL, "chain_fwd_ts2: the colour head's spill columns must be the dead buffer");
        grp(s, GS_TMEM, W_K0, hout, 4, 0);
        blk(s, BK_RAW, 0, 64, kTs2Acc, 0, 0, 0, 0, 64);
        blk(s, BK_RAW, 64, 80, kTs2Acc + 64, W_K1, 0, 0, 64, 67);
      }
    }
  }
  NEFES_REQUIRE(n <= kTs2MaxSteps && bias_floats * 4 <= (int)kChainBiasBytes, NEFES_EINVAL, "chain_fwd_ts2: step table overflow");
  for (int i = 0; i < n; ++i)
    NEFES_REQUIRE(c.step[i].w_bytes <= kTs2WSlot, NEFES_EINVAL, "chain_fwd_ts2: weight image of step %d exceeds the ring slot", i);
  // the issuer's program (Ts2Run records).  It depends on the geometry of the step table only, which is the same for
  // every call of a (mode, saves) combination: built and uploaded once per combination.
  static Ts2Run* d_runs[3][2] = {};
  static int n_runs_of[3][2] = {};
  const int pmode = mode == NEFES_MODE_FULL ? 0 : (mode == NEFES_MODE_STATIC ? 1 : 2);
  if (d_runs[pmode][keep_all ? 1 : 0] == nullptr) {
    static Ts2Run h_runs[kTs2MaxRuns];
    int nr = 0;
    for (int i = 0; i < n; ++i) {
      const Ts2Step& s = c.step[i];
      const uint32_t lbo = s.w_rows * 16u;
      NEFES_REQUIRE((lbo >> 4) < 0x4000u, NEFES_EINVAL, "chain_fwd_ts2: LBO overflow");
      const int first = nr;
      // a 128-wide hidden layer (two 64-column HID blocks on accumulator columns 0 / 64 over the xyz encoding and / or the
      // 8 K-steps of one activation buffer) is ONE record for the issuer's straight-line template; further blocks of the
      // step (the sigma column of the final layer) follow as ordinary runs
      int b_first = 0;
      {
        const bool two = s.n_blk >= 2 && (s.blk[0].kind == BK_HID_RELU || s.blk[0].kind == BK_HID) && s.blk[1].kind == s.blk[0].kind &&
                         s.blk[0].nw == 64 && s.blk[1].nw == 64 && s.blk[0].n0 == 0 && s.blk[1].n0 == 64 &&
                         s.blk[0].acc_col == kTs2Acc && s.blk[1].acc_col == kTs2Acc + 64 && s.w_rows == 128;
        int var = -1; uint32_t hin = 0, wait = 0;
        if (two) {
          const Ts2Group* G = s.grp;
          auto is_h = [&](int j0) { return G[j0].src == GS_TMEM && G[j0 + 1].src == GS_TMEM && G[j0].ksteps == 4 && G[j0 + 1].ksteps == 4 &&
                                           G[j0 + 1].col == G[j0].col + 32 && G[j0 + 1].k0 == G[j0].k0 + 4; };
          if (s.n_grp == 2 && is_h(0) && G[0].k0 == 0) { var = 0; hin = G[0].col; }
          else if (s.n_grp == 1 && G[0].src == GS_X && G[0].ksteps == 4 && G[0].k0 == 0 && G[0].col == 0) { var = 1; }
          else if (s.n_grp == 3 && G[0].src == GS_X && G[0].ksteps == 4 && G[0].k0 == 0 && G[0].col == 0 && is_h(1) && G[1].k0 == 4) { var = 2; hin = G[1].col; }
          for (int j = 0; j < s.n_grp; ++j) wait |= G[j].wait;
          wait |= s.blk[0].wait | s.blk[1].wait;
        }
        if (var >= 0) {
          NEFES_REQUIRE(nr < kTs2MaxRuns, NEFES_EINVAL, "chain_fwd_ts2: run table overflow");
          Ts2Run& R = h_runs[nr++];
          R = Ts2Run{};
          R.d_col = kTs2Acc; R.idesc = idesc_bf16(128, 64, 0, 0); R.a = hin; R.nks = (uint32_t)var;
          R.b16 = 0; R.lbo_field = (lbo >> 4) << 16; R.dbk = (2u * lbo) >> 4;
          R.flags = (uint32_t)RF_HID | (wait << RF_WAIT_SHIFT);
          b_first = 2;
        }
      }
      for (int b = b_first; b < s.n_blk; ++b)
        for (int j = 0; j < s.n_grp; ++j) {
          NEFES_REQUIRE(nr < kTs2MaxRuns, NEFES_EINVAL, "chain_fwd_ts2: run table overflow");
          Ts2Run& R = h_runs[nr++];
          R = Ts2Run{};
          R.d_col = s.blk[b].acc_col; R.idesc = idesc_bf16(128, s.blk[b].nw, 0, 0);
          R.a = s.grp[j].col; R.nks = s.grp[j].ksteps;
          R.b16 = (s.grp[j].k0 * 2u * lbo + s.blk[b].n0 * 16u) >> 4;
          R.lbo_field = (lbo >> 4) << 16; R.dbk = (2u * lbo) >> 4;
          // a K-group's operands are waited for where block 0 first reads them (later blocks read the same columns); a
          // block's own wait (its accumulator columns drained) sits on its first run
          uint32_t wait = (b == 0 ? s.grp[j].wait : 0u) | (j == 0 ? s.blk[b].wait : 0u);
          uint32_t flags = (uint32_t)s.grp[j].src | (j == 0 ? (uint32_t)RF_FRESH : 0u) | (wait << RF_WAIT_SHIFT);
          if (j == s.n_grp - 1) flags |= (uint32_t)(b + 1) << RF_COMMIT_SHIFT;
          R.flags = flags;
        }
      h_runs[first].flags |= RF_FIRST;
      h_runs[nr - 1].flags |= RF_END;
    }
    Ts2Run* dp = nullptr;
    NEFES_CUDA(cudaMalloc(&dp, sizeof(h_runs)));
    NEFES_CUDA(cudaMemcpy(dp, h_runs, sizeof(h_runs), cudaMemcpyHostToDevice));
    d_runs[pmode][keep_all ? 1 : 0] = dp;
    n_runs_of[pmode][keep_all ? 1 : 0] = nr;
  }
  c.runs = d_runs[pmode][keep_all ? 1 : 0];
  c.n_runs = n_runs_of[pmode][keep_all ? 1 : 0];
  c.n_steps = n; c.M = M; c.n_tiles = T;
  c.raw = raw_t; c.C = (mode == NEFES_MODE_SIGMA) ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137);
  c.x_img = w.X.p; c.d_img = (mode == NEFES_MODE_SIGMA) ? nullptr : w.DIRPE.p;
  const bool saves = keep_all && !(chain_xflags() & 1);
  c.n_slots = ts2_slots(saves); c.off_ring = ts2_off_ring(saves);
  c.x_issue = x_last + c.n_slots; c.d_issue = d_last < 0 ? -1 : d_last + c.n_slots;
  NEFES_REQUIRE(c.x_issue < 2 * n && c.d_issue < 2 * n, NEFES_EINVAL, "chain_fwd_ts2: encoding request falls two pairs ahead");
  static bool attr_done = false;
  if (!attr_done) {
    NEFES_CUDA(cudaFuncSetAttribute(chain_fwd_ts2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts2_smem(true) > (int)ts2_smem(false) ? (int)ts2_smem(true) : (int)ts2_smem(false)));
    attr_done = true;
  }
  const int n_pairs = (T + 1) / 2;
  const int grid = n_pairs < num_sms() ? n_pairs : num_sms();
  {
    const double save_ch = !keep_all ? 0 : ((mode == NEFES_MODE_SIGMA) ? 8 * 128 : (mode == NEFES_MODE_STATIC ? 8 * 128 + 128 + 64 : 8 * 128 + 128 + 128 + 64 + 64));
    const double in_ch = (mode == NEFES_MODE_SIGMA) ? 64 : 96;
    const double macs = (mode == NEFES_MODE_SIGMA) ? 130944 : (mode == NEFES_MODE_STATIC ? 165632 : 184064);
    prof_begin(mode == NEFES_MODE_FULL ? "chain_fwd_fine" : (mode == NEFES_MODE_STATIC ? "chain_fwd_coarse" : "chain_fwd_sigma"), st,
               (double)M * (2.0 * (save_ch + in_ch) + 4.0 * c.C), (double)M * 2.0 * macs);
  }
  c.dbg = chain_dbg_buf();
  c.xflags = chain_xflags();
  static const int save_mode = [] { const char* e = getenv("NEFES_TS2_STG"); return e ? atoi(e) : 2; }();
  c.save_mode = save_mode;
  chain_fwd_ts2_kernel<<<grid, kChainThreads, ts2_smem(saves), st>>>(c);
  prof_end(st);
  NEFES_CHECK_LAUNCH("chain_fwd_ts2");
  chain_ts_dbg_dump(c.dbg, c.n_steps, st);
  return NEFES_OK;
}

// Fused data-gradient chain: head-gradient images in, the gradient image of every layer's pre-activation out
// (operands of the weight-gradient kernel); ReLU masks come from the saved activations.
int launch_chain_bwd(const Ws& w, const WsB& b, const Arena& A, int mode, int64_t M, bool with_trunk, bool need_wgrad_images,
                     cudaStream_t st) {
  const int T = (int)ceil_div(M, kTile);
  ChainArgs c = {};
  int n = 0;
  // D[pts, n_in] = G[pts, K = out channels] * W: B operand = WT image [K/8][in rows][8], rows [row0, row0 + n_in)
  auto add = [&](int pl, int k_ch, int row0, int n_in, uint32_t a_off, uint32_t out_off, const Img* act, int act_ch0,
                 const Img* save, int save_ch0, int wait_load) {
    ChainStep& s = c.step[n++];
    const PackedDims pd = packed_dims(pl);
    s.a_off = a_off; s.out_off = out_off; s.K = (uint16_t)k_ch; s.N = (uint16_t)n_in; s.out_ch = (uint16_t)n_in;
    s.kind = CK_DGRAD; s.wait_load = (int8_t)wait_load; s.wait_load2 = -1; s.acc0 = 0; s.bias = nullptr; s.bias_off = 0;
    set_weights(s, A.WT(pl), k_ch / 8, pd.K, row0, n_in);
    s.act = act ? act->p + (int64_t)act_ch0 * 256 : nullptr; s.act_tile_stride = act ? (uint32_t)act->tile_stride() : 0u;
    // a gradient image leaves the SM only if somebody reads it: the weight-gradient kernel (training), or the input-
    // gradient GEMMs of the refinement path (G[4] and G[0] for the xyz encoding, GDT for the direction encoding)
    if (save != nullptr && !need_wgrad_images && save != &b.G[4] && save != &b.G[0] && save != &b.GDT) save = nullptr;
    s.gdst = save ? save->p + (int64_t)save_ch0 * 256 : nullptr; s.g_tile_stride = save ? (uint32_t)save->tile_stride() : 0u;
  };
  auto load = [&](int idx, const Img& img, int ch0, int nch, uint32_t dst_off, int issue_step, int next_pair) {
    ChainLoad& L = c.load[idx];
    L.src = img.p + (int64_t)ch0 * 256; L.tile_stride = (uint32_t)img.tile_stride(); L.bytes = (uint32_t)nch * 256u;
    L.dst_off = dst_off; L.issue_step = (int8_t)issue_step; L.next_pair = (int8_t)next_pair;
  };
  enum { LD_RGB = 0, LD_TH = 1, LD_SIG = 2 };
  if (mode == NEFES_MODE_SIGMA) {
    add(PL_SIG, 16, 0, 128, kRegS, kRegQ, &w.H[7], 0, &b.G[7], 0, LD_SIG);
    load(LD_SIG, b.GSIG, 0, 16, kRegS, 1, 1);
    c.n_loads = 3;
  } else {
    add(PL_RGB, 144, 0, 64, kRegP, kRegQ, &w.DT, 0, &b.GDT, 0, LD_RGB);
    if (mode == NEFES_MODE_FULL) {
      add(PL_TH, 16, 0, 64, kRegS, kRegP, &w.T3, 0, &b.GT3, 0, LD_TH);
      add(PL_TE2, 64, 0, 64, kRegP, kRegP + 16384, &w.T2, 0, &b.GT2, 0, -1);
      add(PL_TE1, 64, 0, 64, kRegP + 16384, kRegQ + 16384, &w.DT, 64, &b.GDT, 64, -1);
      add(PL_DT, 128, 0, 128, kRegQ, kRegP, nullptr, 0, &b.GFS, 0, -1);
      load(LD_TH, b.GTH, 0, 16, kRegS, 2, 1);
    } else {
      add(PL_DIR, 64, 0, 128, kRegQ, kRegP, nullptr, 0, &b.GFS, 0, -1);
    }
    const int fs_step = n;
    add(PL_FS, 144, 0, 128, kRegP, kRegQ, &w.H[7], 0, &b.G[7], 0, LD_SIG);
    load(LD_RGB, b.GRGB, 0, 144, kRegP, fs_step + 1, 1);          // P is free once the FS MMAs retired
    load(LD_SIG, b.GFS, 128, 16, kRegP + 32768, 1, 0);            // P[32K:36K] is free once the RGB MMAs retired
    c.n_loads = 3;
  }
  if (with_trunk)
    for (int l = 7; l >= 1; --l)   // G[l] = grad wrt pre-activation of trunk layer l  ->  G[l-1]
      add(PL_T0 + l, 128, l == 4 ? 64 : 0, 128, kRegQ, kRegQ, &w.H[l - 1], 0, &b.G[l - 1], 0, -1);
  c.n_steps = n;
  c.M = M; c.n_tiles = T;
  c.dbg = chain_dbg_buf();
  c.xflags = chain_xflags();
  for (int l = 0; l < kChainLoads; ++l)
    if (c.load[l].bytes == 0) c.load[l].issue_step = -1;
  static bool attr_done = false;
  if (!attr_done) {
    NEFES_CUDA(cudaFuncSetAttribute(chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdChainSmem));
    attr_done = true;
  }
  const int n_pairs = (T + 1) / 2;
  const int grid = n_pairs < num_sms() ? n_pairs : num_sms();
  {
    double ch = 0, macs = 0;     // channels moved per point: operand loads + mask activations + saved gradient images
    for (int i = 0; i < n; ++i) { ch += (c.step[i].act ? c.step[i].out_ch : 0) + (c.step[i].gdst ? c.step[i].out_ch : 0); macs += (double)c.step[i].K * c.step[i].N; }
    for (int l = 0; l < kChainLoads; ++l) ch += c.load[l].bytes / 256.0;
    prof_begin(with_trunk ? "chain_bwd_full" : "chain_bwd_heads", st, (double)M * 2.0 * ch, (double)M * 2.0 * macs);
  }
  chain_kernel<true><<<grid, kChainThreads, kBwdChainSmem, st>>>(c);
  prof_end(st);
  NEFES_CHECK_LAUNCH("chain_bwd");
  chain_dbg_dump("bwd", c, st);
  return NEFES_OK;
}
}  // namespace

namespace {
bool bulk_flush() {
  static const bool on = [] { const char* e = getenv("NEFES_BULK_FLUSH"); return e == nullptr || atoi(e) != 0; }();
  return on;
}
// Fused data + weight gradients of the eight trunk layers: one launch walking four passes of two layers each (mlp_trunk_bwd.cuh).
int launch_trunk_bwd(const Ws& w, const WsB& b, const Arena& A, int net, int64_t M, float* dP, cudaStream_t st) {
  const int T = (int)ceil_div(M, kTile);
  static bool attr_done = false;
  if (!attr_done) {
    NEFES_CUDA(cudaFuncSetAttribute(trunk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrunkSmem));
    attr_done = true;
  }
  const int grid = T < num_sms() ? T : num_sms();
  // one launch walking the four layer pairs (TrunkMulti), or -- NEFES_TRUNK_PASSES=1 -- four launches of one pair each
  static const int per_launch = [] { const char* e = getenv("NEFES_TRUNK_PASSES"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 4 ? 4 : v); }();
  TrunkMulti mp = {};
  double bytes = 0, flops = 0;
  for (int grp = 0; grp < 4; ++grp) {
    const int l0 = 7 - 2 * grp, l1 = l0 - 1;        // layers of this pass: T_l0 then T_l1
    TrunkArgs& t = mp.pass[mp.n_pass++];
    t = TrunkArgs{};
    t.g_in = b.G[l0].p; t.g_in_tile_stride = (uint32_t)b.G[l0].tile_stride();
    for (int j = 0; j < 2; ++j) {
      const int l = j == 0 ? l0 : l1;
      TrunkStep& s = t.step[j];
      const PackedDims pd = packed_dims(PL_T0 + l);
      s.wt_img = A.WT(PL_T0 + l); s.wt_rows = (uint32_t)pd.K; s.wt_row0 = l == 4 ? 64u : 0u;
      s.has_dgrad = l > 0;
      const Img& act = l > 0 ? w.H[l - 1] : w.X;
      s.act = act.p; s.act_tile_stride = (uint32_t)act.tile_stride(); s.act_ch = act.ch;
      s.pl = PL_T0 + l; s.k_off = l == 4 ? 64 : 0; s.bias = 1;
      if (l == 5) { s.g_save = b.G[4].p; s.g_save_tile_stride = (uint32_t)b.G[4].tile_stride(); }   // operand of the T4 xyz-part job
    }
    if (l1 > 0) { t.g_out = b.G[l1 - 1].p; t.g_out_tile_stride = (uint32_t)b.G[l1 - 1].tile_stride(); }
    t.n_tiles = T; t.ps = pack_src(net); t.d_flat = dP; t.bulk_flush = bulk_flush() ? 1 : 0;
    bytes += (double)M * 2.0 * (128 + t.step[0].act_ch + t.step[1].act_ch + (t.g_out ? 128 : 0) + (t.step[0].g_save ? 128 : 0));
    for (int j = 0; j < 2; ++j) flops += (double)M * 2.0 * ((t.step[j].has_dgrad ? 128.0 * 128 : 0) + 128.0 * t.step[j].act_ch);
    if (mp.n_pass == per_launch || grp == 3) {
      prof_begin("trunk_bwd", st, bytes, flops);
      trunk_bwd_kernel<<<grid, kTrunkThreads, kTrunkSmem, st>>>(mp);
      prof_end(st);
      NEFES_CHECK_LAUNCH("trunk_bwd");
      mp.n_pass = 0; bytes = flops = 0;
    }
  }
  return NEFES_OK;
}
}  // namespace

namespace {
// ---- program builder for fused_bwd_kernel (mlp_fused_bwd.cuh) ------------------------------------------------------
struct FusedBuilder {
  FusedArgs a = {};
  int np = 0, nm[2] = {0, 0}, ne = 0, nbar = 0, mp = 0;   // mp: MMA program being written (0 dgrad thread, 1 wgrad thread)
  uint32_t smem = 0;
  uint32_t alloc(uint32_t bytes) { const uint32_t o = smem; smem += (bytes + 127u) & ~127u; return o; }
  int bar(int count) { a.bar_count[nbar] = (uint16_t)count; return nbar++; }
  // producer
  void p_wait(int b, int flags = 0) { FProdOp& o = a.prod[np++]; o.kind = FO_WAIT; o.bar = (uint8_t)b; o.flags = (uint8_t)flags; }
  void p_load(int b, const uint8_t* src, uint32_t tile_stride, uint32_t bytes, uint32_t off, int next) {
    FProdOp& o = a.prod[np++]; o.kind = FO_LOAD; o.bar = (uint8_t)b; o.next = (uint8_t)next; o.src = src; o.tile_stride = tile_stride;
    o.bytes = bytes; o.smem_off = off;
  }
  void p_store(uint8_t* dst, uint32_t tile_stride, uint32_t bytes, uint32_t off) {
    FProdOp& o = a.prod[np++]; o.kind = FO_STORE; o.dst = dst; o.tile_stride = tile_stride; o.bytes = bytes; o.smem_off = off;
  }
  void p_arrive(int b) { FProdOp& o = a.prod[np++]; o.kind = FO_ARRIVE; o.bar = (uint8_t)b; }
  // MMA issuer
  void m_wait(int b, int flags = 0) { FMmaOp& o = a.mma[mp][nm[mp]++]; o.kind = FO_WAIT; o.bar = (uint8_t)b; o.flags = (uint8_t)flags; }
  void m_commit(int b) {                                       // folded into the MMA op it follows
    if (nm[mp] > 0 && a.mma[mp][nm[mp] - 1].kind == FO_MMA && !(a.mma[mp][nm[mp] - 1].flags & FX_THEN)) {
      FMmaOp& m = a.mma[mp][nm[mp] - 1]; m.flags |= FX_THEN; m.bar = (uint8_t)b; return;
    }
    FMmaOp& o = a.mma[mp][nm[mp]++]; o.kind = FO_COMMIT; o.bar = (uint8_t)b;
  }
  // data gradient: acc[128 pts, n] = G[pts, k_ch] (K-major image at g_off) * WT (image [k_ch/8][wt_rows][8] at w_off)
  void m_dgrad(uint32_t g_off, int k_ch, uint32_t w_off, int wt_rows, int n, int col) {
    FMmaOp& o = a.mma[mp][nm[mp]++]; o.kind = FO_MMA; o.accmode = FA_FRESH; o.ksteps = (uint8_t)(k_ch / 16);
    o.a_off = g_off; o.a_lbo = 128; o.a_sbo = 8; o.a_adv = 256;
    o.b_off = w_off; o.b_lbo = (uint16_t)wt_rows; o.b_sbo = 8; o.b_adv = (uint16_t)(2 * wt_rows);
    o.tmem_col = (uint16_t)col; o.idesc = idesc_bf16(128, n, 0, 0);
  }
  // weight gradient: D[128 rows of the image at a_off, n channels of the image at b_off] += A^T B over the 128 points
  void m_wgrad(uint32_t a_off, uint32_t b_off, int n, int col) {
    FMmaOp& o = a.mma[mp][nm[mp]++]; o.kind = FO_MMA; o.accmode = FA_LAUNCH; o.ksteps = 8;
    o.a_off = a_off; o.a_lbo = 8; o.a_sbo = 128; o.a_adv = 16;
    o.b_off = b_off; o.b_lbo = 8; o.b_sbo = 128; o.b_adv = 16;
    o.tmem_col = (uint16_t)col; o.idesc = idesc_bf16(128, n, 1, 1);
  }
  // epilogue
  void e_wait(int b, int flags = 0) { FEpiOp& o = a.epi[ne++]; o.kind = FO_WAIT; o.bar = (uint8_t)b; o.flags = (uint8_t)flags; }
  void e_arrive(int b) {                                       // folded into the EPI op it follows
    if (ne > 0 && a.epi[ne - 1].kind == FO_EPI && !(a.epi[ne - 1].flags & FX_THEN)) { a.epi[ne - 1].flags |= FX_THEN; a.epi[ne - 1].bar = (uint8_t)b; return; }
    FEpiOp& o = a.epi[ne++]; o.kind = FO_ARRIVE; o.bar = (uint8_t)b;
  }
  void e_epi(int col, int n, bool mask, uint32_t mask_off, uint32_t out_off) {
    FEpiOp& o = a.epi[ne++]; o.kind = FO_EPI; o.acc_col = (uint16_t)col; o.n = (uint16_t)n; o.has_mask = mask ? 1 : 0;
    o.mask_off = mask_off; o.out_off = out_off;
  }
  void flush(int col, int n_cols, int pl, int m0, int k_off, int kind) {
    FFlush& f = a.flush[a.n_flush++]; f.tmem_col = (uint16_t)col; f.n_cols = (uint16_t)n_cols; f.pl = (int16_t)pl; f.m0 = (int16_t)m0;
    f.k_off = (int16_t)k_off; f.kind = (uint8_t)kind;
  }
  void ones(uint32_t off) { a.ones_off[a.n_ones++] = off; }
  // weight image, rows [row0, row0 + take) of every chunk compacted, loaded once
  void p_weights(int b, const uint8_t* img, int chunks, int rows, int row0, int take, uint32_t off, int* n_loads) {
    if (take == rows) {
      const uint32_t bytes = (uint32_t)chunks * rows * 16u;
      for (uint32_t o = 0; o < bytes; o += 16384u) { p_load(b, img + o, 0, bytes - o < 16384u ? bytes - o : 16384u, off + o, 2); ++*n_loads; }
    } else {
      for (int c = 0; c < chunks; ++c) { p_load(b, img + ((int64_t)c * rows + row0) * 16, 0, (uint32_t)take * 16u, off + (uint32_t)c * take * 16u, 2); ++*n_loads; }
    }
  }
  int finish(int n_tiles, int net, float* dP, const char* what) {
    NEFES_REQUIRE(np < kFMaxProd && nm[0] < kFMaxMma && nm[1] < kFMaxMma && ne < kFMaxEpi && a.n_flush <= kFMaxFlush &&
                  nbar <= kFMaxBars && a.n_ones <= 4,
                  NEFES_EINVAL, "%s: program table overflow (%d %d %d %d %d %d)", what, np, nm[0], nm[1], ne, a.n_flush, nbar);
    a.prod[np].kind = FO_END; a.mma[0][nm[0]].kind = FO_END; a.mma[1][nm[1]].kind = FO_END; a.epi[ne].kind = FO_END;
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i < nm[t]; ++i) {
        FMmaOp& o = a.mma[t][i];
        if (o.kind != FO_MMA) continue;
        auto rel = [](uint32_t off, uint32_t lbo16, uint32_t sbo16) {      // tc05.cuh smem_desc with the address relative to the base
          return (uint64_t)((off >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
        };
        o.da_rel = rel(o.a_off, o.a_lbo, o.a_sbo);
        o.db_rel = rel(o.b_off, o.b_lbo, o.b_sbo);
        o.misc = (uint32_t)o.tmem_col | ((uint32_t)o.ksteps << 16) | ((uint32_t)o.accmode << 24);
        o.adv = (uint32_t)o.a_adv | ((uint32_t)o.b_adv << 16);
      }
    a.n_tiles = n_tiles; a.ps = pack_src(net); a.d_flat = dP;
    return NEFES_OK;
  }
};

int launch_fused(FusedBuilder& B, int n_tiles, cudaStream_t st, const char* what) {
  static uint32_t attr_bytes = 0;
  NEFES_REQUIRE(B.smem <= 228000u, NEFES_EINVAL, "%s: shared memory plan exceeds the SM (%u B)", what, B.smem);
  if (B.smem > attr_bytes) {
    NEFES_CUDA(cudaFuncSetAttribute(fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.smem));
    attr_bytes = B.smem;
  }
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  B.a.dbg = chain_dbg_buf();
  if (B.a.dbg) cudaMemsetAsync(B.a.dbg, 0, 2048 * sizeof(long long), st);
  {
    double bytes = 0, macs = 0;  // per tile: every per-tile load and store; MACs of every MMA group
    for (int i = 0; i < B.np; ++i) if ((B.a.prod[i].kind == FO_LOAD && B.a.prod[i].next != 2) || B.a.prod[i].kind == FO_STORE) bytes += B.a.prod[i].bytes;
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i < B.nm[t]; ++i)
        if (B.a.mma[t][i].kind == FO_MMA) macs += 128.0 * (((B.a.mma[t][i].idesc >> 17) & 0x3F) * 8) * B.a.mma[t][i].ksteps * 16;
    prof_begin(what, st, bytes * n_tiles, 2.0 * macs * n_tiles);
  }
  fused_bwd_kernel<<<grid, kFusedThreads, B.smem, st>>>(B.a);
  prof_end(st);
  NEFES_CHECK_LAUNCH(what);
  if (B.a.dbg) {
    static int dumps = 0;
    if (dumps++ < 4) {
      cudaStreamSynchronize(st);
      static long long h[2048];
      cudaMemcpy(h, B.a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
      const long long t0 = h[(1 * 4 + 0) * 48 + 0];
      static const char* kn[] = {"END", "WAIT", "LOAD", "STORE", "ARRIVE", "MMA", "COMMIT", "EPI"};
      fprintf(stderr, "[fused dbg] %s  (cycles since the MMA issuer's first op)\n", what);
      for (int it = 1; it < 3; ++it) {
        fprintf(stderr, " tile %d producer:", it);
        for (int i = 0; i < 48 && B.a.prod[i].kind != FO_END; ++i) fprintf(stderr, " %s%d@%lld", kn[B.a.prod[i].kind], B.a.prod[i].bar, h[(0 * 4 + it) * 48 + i] - t0);
        fprintf(stderr, "\n tile %d mma:", it);
        for (int i = 0; i < 48 && B.a.mma[0][i].kind != FO_END; ++i) fprintf(stderr, " %s%d@%lld", kn[B.a.mma[0][i].kind], B.a.mma[0][i].bar, h[(1 * 4 + it) * 48 + i] - t0);
        fprintf(stderr, "\n tile %d epi:", it);
        for (int i = 0; i < 48 && B.a.epi[i].kind != FO_END; ++i) fprintf(stderr, " %s%d@%lld", kn[B.a.epi[i].kind], B.a.epi[i].bar, h[(2 * 4 + it) * 48 + i] - t0);
        fprintf(stderr, "\n");
      }
    }
  }
  return NEFES_OK;
}

// Head group 1: rgb+feature head, and (fine net) the transient heads and the two transient hidden layers.
//   in : GRGB (144 ch), GTH (16 ch) gradient images; saved DT (dir | tenc0 hidden), T3, T2
//   out: GDT (gradient wrt the pre-activations of [dir | tenc0]) -> HBM, operand of head group 2
int launch_heads1(const Ws& w, const WsB& b, const Arena& A, int net, int mode, int64_t M, float* dP, cudaStream_t st) {
  const int T = (int)ceil_div(M, kTile);
  const bool fine = mode == NEFES_MODE_FULL;
  FusedBuilder B;
  const uint32_t W0 = B.alloc(18 * 64 * 16), W1 = B.alloc(2 * 64 * 16), W2 = B.alloc(8 * 64 * 16), W3 = B.alloc(8 * 64 * 16);
  const uint32_t G0 = B.alloc(36864), G1 = B.alloc(4096), G2 = B.alloc(16384), G3 = B.alloc(16384), G4 = B.alloc(32768);
  const uint32_t Adir = B.alloc(20480), Atenc = B.alloc(20480), At3 = B.alloc(20480), At2 = B.alloc(20480);
  B.ones(Adir + 16384);
  if (fine) { B.ones(Atenc + 16384); B.ones(At3 + 16384); B.ones(At2 + 16384); }
  int n_w = 0;
  const int bW = B.bar(0);
  B.p_weights(bW, A.WT(PL_RGB), 18, 64, 0, 64, W0, &n_w);
  if (fine) {
    B.p_weights(bW, A.WT(PL_TH), 2, 64, 0, 64, W1, &n_w);
    B.p_weights(bW, A.WT(PL_TE2), 8, 64, 0, 64, W2, &n_w);
    B.p_weights(bW, A.WT(PL_TE1), 8, 64, 0, 64, W3, &n_w);
  }
  B.a.bar_count[bW] = (uint16_t)n_w;
  const int L_GRGB = B.bar(1), L_DIR = B.bar(1), ACC0 = B.bar(1), E0 = B.bar(256), D0 = B.bar(1), OUTFREE = B.bar(1);
  int L_GTH = -1, L_TENC = -1, L_T3 = -1, L_T2 = -1, ACC1 = -1, ACC2 = -1, ACC3 = -1, E1 = -1, E2 = -1, E3 = -1, D1 = -1, D2 = -1, D3 = -1;
  if (fine) {
    L_GTH = B.bar(1); L_TENC = B.bar(1); L_T3 = B.bar(1); L_T2 = B.bar(1);
    ACC1 = B.bar(1); ACC2 = B.bar(1); ACC3 = B.bar(1); E1 = B.bar(256); E2 = B.bar(256); E3 = B.bar(256);
    D1 = B.bar(1); D2 = B.bar(1); D3 = B.bar(1);
  }
  const int Elast = fine ? E3 : E0;
  const uint32_t dt_stride = (uint32_t)w.DT.tile_stride();
  // ---- producer
  B.p_load(L_GRGB, b.GRGB.p, (uint32_t)b.GRGB.tile_stride(), 36864, G0, 1);
  B.p_load(L_DIR, w.DT.p, dt_stride, 16384, Adir, 1);
  if (fine) {
    B.p_load(L_GTH, b.GTH.p, (uint32_t)b.GTH.tile_stride(), 4096, G1, 1);
    B.p_load(L_TENC, w.DT.p + 64 * 256, dt_stride, 16384, Atenc, 1);
    B.p_load(L_T3, w.T3.p, (uint32_t)w.T3.tile_stride(), 16384, At3, 1);
    B.p_load(L_T2, w.T2.p, (uint32_t)w.T2.tile_stride(), 16384, At2, 1);
  }
  // (the loads above run for the first tile before the loop; inside the loop each one refills its slot for tile it+1
  //  right after the waits that precede it in program order)
  {
    FusedBuilder P;   // reorder: waits first, then the matching load -- rebuild the producer program
    P = B; P.np = 0;
    auto reload = [&](int idx) { P.a.prod[P.np++] = B.a.prod[idx]; };
    int li = 0;
    for (; B.a.prod[li].kind == FO_LOAD && B.a.prod[li].next == 2; ++li) reload(li);      // once-loads (weights)
    const int iGRGB = li, iDIR = li + 1, iGTH = li + 2, iTENC = li + 3, iT3 = li + 4, iT2 = li + 5;
    P.p_wait(D0); P.p_wait(E0); reload(iGRGB); reload(iDIR);       // D: weight-gradient reads done, E: data-gradient + mask reads done
    if (fine) {
      P.p_wait(D1); P.p_wait(E1); reload(iGTH); reload(iT3);
      P.p_wait(D2); P.p_wait(E2); reload(iT2);
      P.p_wait(D3); P.p_wait(E3); reload(iTENC);
    }
    P.p_store(b.GDT.p, (uint32_t)b.GDT.tile_stride(), fine ? 32768u : 16384u, G4);
    P.p_arrive(OUTFREE);
    B = P;
  }
  // ---- data-gradient issuer: the serial chain  RGB -> (fine) TH -> TE2 -> TE1
  B.mp = 0;
  B.m_wait(bW, FW_ONCE);
  B.m_wait(L_GRGB); B.m_wait(Elast, FW_PREV);                 // accumulator drained by the last epilogue of the previous tile
  B.m_dgrad(G0, 144, W0, 64, 64, 0); B.m_commit(ACC0);
  if (fine) {
    B.m_wait(L_GTH); B.m_wait(E0);
    B.m_dgrad(G1, 16, W1, 64, 64, 0); B.m_commit(ACC1);
    B.m_wait(E1);
    B.m_dgrad(G2, 64, W2, 64, 64, 0); B.m_commit(ACC2);
    B.m_wait(E2);
    B.m_dgrad(G3, 64, W3, 64, 64, 0); B.m_commit(ACC3);
  }
  // ---- weight-gradient issuer (own accumulators; its operands are handed over by explicit barriers)
  B.mp = 1;
  B.m_wait(L_GRGB); B.m_wait(L_DIR);
  B.m_wgrad(G0, Adir, 80, 64); B.m_wgrad(G0 + 32768, Adir, 80, 144); B.m_commit(D0);
  if (fine) {
    B.m_wait(L_GTH); B.m_wait(L_T3);
    B.m_wgrad(G1, At3, 80, 224); B.m_commit(D1);
    B.m_wait(L_T2); B.m_wait(E1);                              // GT3 written by epilogue 1
    B.m_wgrad(G2, At2, 80, 304); B.m_commit(D2);
    B.m_wait(L_TENC); B.m_wait(E2);                            // GT2 written by epilogue 2
    B.m_wgrad(G3, Atenc, 80, 384); B.m_commit(D3);
  }
  // ---- epilogue (a gradient image may be overwritten only when BOTH streams finished reading its previous content)
  // (every wait costs the interpreter ~300 cycles even when its barrier completed long ago -- stamps, profiles/ -- so the
  //  barrier that completes LAST, the round's accumulator, is waited for last: the others are checked while its MMAs run)
  B.e_wait(L_DIR); B.e_wait(OUTFREE, FW_PREV); B.e_wait(ACC0); B.e_epi(0, 64, true, Adir, G4); B.e_arrive(E0);
  if (fine) {
    B.e_wait(L_T3); B.e_wait(D2, FW_PREV); B.e_wait(ACC1); B.e_epi(0, 64, true, At3, G2); B.e_arrive(E1);
    B.e_wait(L_T2); B.e_wait(D3, FW_PREV); B.e_wait(ACC2); B.e_epi(0, 64, true, At2, G3); B.e_arrive(E2);
    B.e_wait(L_TENC); B.e_wait(ACC3); B.e_epi(0, 64, true, Atenc, G4 + 16384); B.e_arrive(E3);
  }
  // ---- flush
  B.flush(64, 64, PL_RGB, 0, 0, FF_W); B.flush(128, 16, PL_RGB, 0, 0, FF_BIAS);
  B.flush(144, 64, PL_RGB, 128, 0, FF_W); B.flush(208, 16, PL_RGB, 128, 0, FF_BIAS);
  if (fine) {
    B.flush(224, 64, PL_TH, 0, 0, FF_W); B.flush(288, 16, PL_TH, 0, 0, FF_BIAS);
    B.flush(304, 64, PL_TE2, 0, 0, FF_W); B.flush(368, 16, PL_TE2, 0, 0, FF_BIAS);
    B.flush(384, 64, PL_TE1, 0, 0, FF_W); B.flush(448, 16, PL_TE1, 0, 0, FF_BIAS);
  }
  TRY(B.finish(T, net, dP, "heads1"));
  return launch_fused(B, T, st, "fused_bwd heads1");
}

// Head group 2: [dir | tenc0] layer (input [final | dirPE]) and final+sigma layer (input h8).
//   in : GDT (from group 1), sigma gradient chunk (GFS image channels 128..143); saved FIN, DIRPE, H[7]
//   out: G[7] (gradient wrt the pre-activation of the last trunk layer) -> HBM, operand of the trunk launches
int launch_heads2(const Ws& w, const WsB& b, const Arena& A, int net, int mode, int64_t M, float* dP, cudaStream_t st) {
  const int T = (int)ceil_div(M, kTile);
  const bool fine = mode == NEFES_MODE_FULL;
  const int pl_dt = fine ? PL_DT : PL_DIR;
  const int dt_ch = fine ? 128 : 64;
  FusedBuilder B;
  const uint32_t W0 = B.alloc((uint32_t)(dt_ch / 8) * 128 * 16), W1 = B.alloc(18 * 128 * 16);
  const uint32_t G0 = B.alloc(32768), G1 = B.alloc(36864), A0 = B.alloc(40960), A1 = B.alloc(36864);
  B.ones(A1 + 32768);
  int n_w = 0;
  const int bW = B.bar(0);
  B.p_weights(bW, A.WT(pl_dt), dt_ch / 8, 160, 0, 128, W0, &n_w);
  B.p_weights(bW, A.WT(PL_FS), 18, 128, 0, 128, W1, &n_w);
  B.a.bar_count[bW] = (uint16_t)n_w;
  const int L_GDT = B.bar(1), L_GSIG = B.bar(1), L_A0 = B.bar(2), L_H7 = B.bar(1);
  const int ACC0 = B.bar(1), ACC1 = B.bar(1), E0 = B.bar(256), E1 = B.bar(256), D0 = B.bar(1), D1 = B.bar(1);
  static const bool g7_over_gfin = [] { const char* e = getenv("NEFES_HEADS2_OUT"); return e == nullptr || atoi(e) != 0; }();
  const int OUTFREE = g7_over_gfin ? B.bar(1) : -1;
  const uint32_t gdt_bytes = (uint32_t)dt_ch * 256u;
  // ---- producer: first-tile loads, then per tile
  B.p_load(L_GDT, b.GDT.p, (uint32_t)b.GDT.tile_stride(), gdt_bytes, G0, 1);
  B.p_load(L_GSIG, b.GFS.p + 128 * 256, (uint32_t)b.GFS.tile_stride(), 4096, G1 + 32768, 1);
  B.p_load(L_A0, w.FIN.p, (uint32_t)w.FIN.tile_stride(), 32768, A0, 1);
  B.p_load(L_A0, w.DIRPE.p, (uint32_t)w.DIRPE.tile_stride(), 8192, A0 + 32768, 1);
  B.p_load(L_H7, w.H[7].p, (uint32_t)w.H[7].tile_stride(), 32768, A1, 1);
  {
    FusedBuilder P;
    P = B; P.np = 0;
    auto reload = [&](int idx) { P.a.prod[P.np++] = B.a.prod[idx]; };
    int li = 0;
    for (; B.a.prod[li].kind == FO_LOAD && B.a.prod[li].next == 2; ++li) reload(li);
    const int iGDT = li, iGSIG = li + 1, iFIN = li + 2, iDIRPE = li + 3, iH7 = li + 4;
    if (g7_over_gfin) {
      // G7 (output) is written over the gFIN image in G1 once BOTH streams are done with it (ACC1, D1 -- they finish about
      // together): the [final | dirPE] slot A0 is then free as soon as the first weight-gradient GEMM retired, and the next
      // tile's FIN / DIRPE -- which gate that GEMM, and through D0 nothing else any more -- are fetched early in the tile
      // instead of behind the store at its very end.
      P.p_wait(D0); reload(iFIN); reload(iDIRPE);
      P.p_wait(E0); reload(iGDT);
      P.p_wait(D1); P.p_wait(E1); reload(iH7); reload(iGSIG);
      P.p_store(b.G[7].p, (uint32_t)b.G[7].tile_stride(), 32768, G1);
      P.p_arrive(OUTFREE);
    } else {
      // G7 (output) is written over the [final | dirPE] slot A0, which only the weight-gradient GEMM of the first layer
      // reads: the gradient input slot G0 is then free early in the tile and the NEXT tile's GDT -- the head of the
      // serial data-gradient chain -- is prefetched; what waits for the store is A0, needed only by the weight-gradient stream
      P.p_wait(D0); P.p_wait(E0); reload(iGDT);
      P.p_wait(D1); P.p_wait(E1); reload(iH7); reload(iGSIG);
      P.p_store(b.G[7].p, (uint32_t)b.G[7].tile_stride(), 32768, A0);
      reload(iFIN); reload(iDIRPE);
    }
    B = P;
  }
  // ---- data-gradient issuer
  B.mp = 0;
  B.m_wait(bW, FW_ONCE);
  B.m_wait(L_GDT); B.m_wait(E1, FW_PREV);
  B.m_dgrad(G0, dt_ch, W0, 128, 128, 0); B.m_commit(ACC0);
  B.m_wait(L_GSIG); B.m_wait(E0);
  B.m_dgrad(G1, 144, W1, 128, 128, 0); B.m_commit(ACC1);
  // ---- weight-gradient issuer
  B.mp = 1;
  B.m_wait(L_GDT); B.m_wait(L_A0);
  B.m_wgrad(G0, A0, 160, 128); B.m_commit(D0);
  B.m_wait(L_GSIG); B.m_wait(L_H7); B.m_wait(E0);
  B.m_wgrad(G1, A1, 144, 288);
  B.m_wgrad(A1, G1 + 32768, 16, 432);                 // sigma row, transposed: D[h8 channel, 0] = sum_p h8[p, ch] * gsig[p]
  B.m_commit(D1);
  // ---- epilogue
  if (g7_over_gfin) {
    B.e_wait(D1, FW_PREV); B.e_wait(OUTFREE, FW_PREV); B.e_wait(ACC0); B.e_epi(0, 128, false, 0, G1); B.e_arrive(E0);   // gFIN over the previous G7 (stored)
    B.e_wait(L_H7); B.e_wait(ACC1); B.e_wait(D1); B.e_epi(0, 128, true, A1, G1); B.e_arrive(E1);                        // G7 over gFIN (read by ACC1, D1)
  } else {
    B.e_wait(ACC0); B.e_wait(D1, FW_PREV); B.e_epi(0, 128, false, 0, G1); B.e_arrive(E0);        // gFIN over the previous gFS
    B.e_wait(ACC1); B.e_wait(L_H7); B.e_wait(D0); B.e_epi(0, 128, true, A1, A0); B.e_arrive(E1); // G7 over [final|dirPE] (read by D0)
  }
  // ---- flush: the dirPE image carries a constant 1 in its last padding channel (column 128 + 31): bias of [dir | tenc0]
  B.flush(128, 160, pl_dt, 0, 0, FF_W); B.flush(128 + 144, 16, pl_dt, 0, 15, FF_BIAS);
  B.flush(288, 128, PL_FS, 0, 0, FF_W); B.flush(288 + 128, 16, PL_FS, 0, 0, FF_BIAS);
  B.flush(432, 16, PL_FS, 128, 0, FF_WT);
  TRY(B.finish(T, net, dP, "heads2"));
  return launch_fused(B, T, st, "fused_bwd heads2");
}
}  // namespace

int mlp_workspace_bf16(int net, int mode, int64_t M, int64_t N, int64_t* saved, int64_t* sf, int64_t* sb) {
  (void)net;
  const int64_t T = ceil_div(M, kTile);
  const int C = (mode == NEFES_MODE_SIGMA) ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137);
  *saved = carve_ws(nullptr, T, mode).bytes + 1024;
  *sf = T * C * kTile * 4 + 1024;                  // tile-major raw block when the caller wants rows
  *sb = carve_wsb(nullptr, T, mode, N).bytes + 1024;
  return NEFES_OK;
}

// the weight re-pack of mlp_fwd_bf16 on its own: fp32 parameters -> bf16 operand images in the `saved` workspace of an (N, S, mode) query
int mlp_prepack_bf16(const float* P, int net, int mode, int64_t N, int S, void* saved, cudaStream_t st) {
  const int T = (int)ceil_div(N * S, kTile);
  Ws w = carve_ws(saved, T, mode);
  Arena A = {w.arena, packed_arena()};
  prepack_kernel<<<256, 256, 0, st>>>(P, pack_src(net), A.ar, w.arena, net == NEFES_NET_FINE ? 1 : 0);
  NEFES_CHECK_LAUNCH("prepack");
  return NEFES_OK;
}

int mlp_fwd_bf16(const float* P, int net, int mode, const float* pts, const float* dirs, int64_t N, int S, float* raw,
                 void* saved, void* scratch, int layout, cudaStream_t st) {
  const int64_t M = N * S;
  const int T = (int)ceil_div(M, kTile);
  const int64_t Mp = (int64_t)T * kTile;
  Ws w = carve_ws(saved, T, mode);
  const int C = (mode == NEFES_MODE_SIGMA) ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137);
  const bool fine = net == NEFES_NET_FINE;
  Arena A = {w.arena, packed_arena()};

  if (!weights_packed()) {                           // frozen-weight callers (refinement) pack once per query: mlp_prepack_bf16
    prepack_kernel<<<256, 256, 0, st>>>(P, pack_src(net), A.ar, w.arena, fine ? 1 : 0);
    NEFES_CHECK_LAUNCH("prepack");
  }
  encode_images_kernel<<<(unsigned)ceil_div(Mp, 128), 128, 0, st>>>(pts, dirs, S, M, Mp, w.X.p,
                                                                    mode == NEFES_MODE_SIGMA ? nullptr : w.DIRPE.p);
  NEFES_CHECK_LAUNCH("encode_images");
  const bool direct = (layout == NEFES_RAW_TILES) || C == 1;      // C == 1: the two layouts coincide
  float* raw_t = direct ? raw : reinterpret_cast<float*>(scratch);
  // Two forward chains.  Default since round 2: mlp_chain_ts2.cuh (activations in tensor memory, N-split layers, straight-
  // line issuer) -- fine query at 6144 rays 0.60 ms with saved copies / 0.476 without, against 0.65 / 0.55 for the shared-
  // memory-operand chain (mlp_chain.cuh, NEFES_FWD_SS=1; its template also serves the data-gradient chains).  Round 1's
  // first tensor-memory chain (0.79 / 0.52 ms) was removed once TS2 superseded it.  Both are parity-tested against the oracle.
  static const bool force_ss = getenv("NEFES_FWD_SS") != nullptr;
  if (force_ss) TRY(launch_chain_fwd(w, A, mode, M, raw_t, st));
  else TRY(launch_chain_fwd_ts2(w, A, mode, M, raw_t, st));
  if (!direct) {
    static bool t2r_attr = false;
    if (!t2r_attr) {
      NEFES_CUDA(cudaFuncSetAttribute(tiles_to_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTile * 137 * 4));
      t2r_attr = true;
    }
    tiles_to_rows_kernel<<<T, 512, kTile * C * 4, st>>>(raw_t, raw, M, C);
    NEFES_CHECK_LAUNCH("tiles_to_rows");
  }
  return NEFES_OK;
}

int mlp_bwd_bf16(const float* P, int net, int mode, const float* pts, const float* dirs, int64_t N, int S,
                 const float* raw, const float* d_raw, const void* saved, void* scratch, float* dP, float* d_pts,
                 float* d_dirs, int layout, cudaStream_t st, const float* compact, const float* g_rgb, const float* g_feat) {
  (void)P;
  const int64_t M = N * S;
  const int T = (int)ceil_div(M, kTile);
  const int64_t Mp = (int64_t)T * kTile;
  const Ws w = carve_ws(const_cast<void*>(saved), T, mode);
  WsB b = carve_wsb(scratch, T, mode, N);
  const int C = (mode == NEFES_MODE_SIGMA) ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137);
  Arena A = {w.arena, packed_arena()};

  const Img& gsig_img = (mode == NEFES_MODE_SIGMA) ? b.GSIG : b.GFS;
  const bool tiles = layout == NEFES_RAW_TILES;
  // training (weight gradients, no gradient to the sample positions): fused data+weight-gradient launches -- two for
  // the heads, one (four passes) for the trunk; otherwise (pose refinement) the data-gradient chain runs all layers and keeps every
  // gradient image for the input-gradient GEMMs
  const bool fused_trunk = dP != nullptr && d_pts == nullptr && getenv("NEFES_NO_FUSED_TRUNK") == nullptr;
  // (measured: a gain for the fine net's six head layers, none for the coarse net's three)
  const bool fused_heads = fused_trunk && (mode == NEFES_MODE_FULL || (mode == NEFES_MODE_STATIC && getenv("NEFES_FUSED_HEADS_COARSE") != nullptr)) &&
                           getenv("NEFES_NO_FUSED_HEADS") == nullptr;
  prof_begin("head_grad_images", st, (double)M * ((compact ? 20.0 : 4.0 * C) + 4.0 * (C == 137 ? 6 : 1) + 2.0 * (C == 1 ? 16 : (C == 137 ? 176 : 160))), 0.0);
  head_grad_images_kernel<<<(unsigned)ceil_div(Mp, 128), 128, 0, st>>>(
      raw, d_raw, C, tiles ? 1 : C, tiles ? kTile : 1, M, Mp, b.GRGB.p, b.GTH.p, gsig_img.p, gsig_img.tile_stride(),
      mode == NEFES_MODE_SIGMA ? 0 : 16, fused_heads ? dP + layout_for(net).b[L_SIGMA] : nullptr, compact, g_rgb, g_feat, S);
  prof_end(st);
  NEFES_CHECK_LAUNCH("head_grad_images");
  if (fused_heads) {
    TRY(launch_heads1(w, b, A, net, mode, M, dP, st));
    TRY(launch_heads2(w, b, A, net, mode, M, dP, st));
  } else {
    TRY(launch_chain_bwd(w, b, A, mode, M, !fused_trunk, dP != nullptr, st));
  }
  if (fused_trunk) TRY(launch_trunk_bwd(w, b, A, net, M, dP, st));

  // ---- gradients to the inputs (pose refinement): fp32 out of the GEMM, then the SIMT PE backward ------
  if (d_pts != nullptr) {        // d xyzPE = G5 W_T4[:, :63] + G1 W_T0 as ONE GEMM over the concatenated K = [G5 | G1]
    GemmDesc d;
    ASrc g1 = src_of(b.G[0]);
    d.a[0] = src_of(b.G[4]); d.a[1] = g1; d.n_src = 2;
    d.K = 256; d.N = 64; d.w_img = A.W(PL_DX); d.w_rows = 64;
    // ... and the encoding's own backward on the staged rows, so the [M,64] fp32 cotangent never goes to HBM
    d.raw = d_pts; d.raw_ld = 3; d.raw_col0 = 0; d.d_col0 = 0; d.raw_ncol = 64; d.raw_act = RAW_ACT_NONE;
    d.pe_x = pts; d.pe_L = kXyzFreqs;
    TRY(launch_tile_gemm(d, T, M, st, "dgrad xyz PE"));
  }
  if (d_dirs != nullptr && mode != NEFES_MODE_SIGMA) {   // d dirPE = GDT W_DT[:, 128:155], summed over each ray's samples
    GemmDesc d;
    const int pl = (mode == NEFES_MODE_FULL) ? PL_DT : PL_DIR;
    d.a[0] = src_of(b.GDT); d.K = w.DT.ch; d.N = 32; d.w_img = A.WT(pl); d.w_rows = packed_dims(pl).K; d.w_row0 = 128;
    d.raw_col0 = 0; d.d_col0 = 0; d.raw_ncol = 32; d.raw_act = RAW_ACT_NONE;
    if (S % 16 == 0 && kTile % S == 0) {   // the per-ray sum happens on the staged rows: [M,32] fp32 never goes to HBM
      d.raw = b.dDIRray; d.raw_ld = 32; d.ray_S = S; d.n_rays = N;
      TRY(launch_tile_gemm(d, T, Mp, st, "dgrad dir PE"));
    } else {
      d.raw = b.dDIR; d.raw_ld = 32;
      TRY(launch_tile_gemm(d, T, Mp, st, "dgrad dir PE"));
      ray_reduce_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(b.dDIR, 32, 32, S, N, b.dDIRray);
      NEFES_CHECK_LAUNCH("ray_reduce");
    }
    TRY(nefes_encode_pe_bwd(dirs, b.dDIRray, 32, N, kDirFreqs, d_dirs, st));
  }

  if (dP == nullptr) return NEFES_OK;

  // ---- weight + bias gradients: one persistent launch over all layers ---------------------------
  WgradArgs wa = {};
  int nj = 0;
  uint32_t max_tile_bytes = 0;
  auto job = [&](int pl, const Img& gimg, int g_ch0, int g_ch, ASrc a0, const ASrc* a1, int no_bias = 0) {
    WgradJob& J = wa.job[nj++];
    J.no_bias = no_bias;
    J.g = src_of(gimg, g_ch0, g_ch); J.g_ch = g_ch;
    J.a[0] = a0; J.n_src = 1; J.a_ch = (int)(a0.bytes / 256);
    if (a1) { J.a[1] = *a1; J.n_src = 2; J.a_ch += (int)(a1->bytes / 256); }
    J.pl = pl;
    J.tile_bytes = (uint32_t)(J.g_ch + J.a_ch) * 256u;
    if (J.tile_bytes > max_tile_bytes) max_tile_bytes = J.tile_bytes;
  };
  if (mode == NEFES_MODE_FULL && !fused_heads) {
    job(PL_TH, b.GTH, 0, 16, src_of(w.T3), nullptr);
    job(PL_TE2, b.GT3, 0, 64, src_of(w.T2), nullptr);
    job(PL_TE1, b.GT2, 0, 64, src_of(w.DT, 64, 64), nullptr);
  }
  if (fused_heads) {
    // every head layer was handled by the fused launches
  } else if (mode != NEFES_MODE_SIGMA) {
    ASrc dp = src_of(w.DIRPE);
    job(PL_RGB, b.GRGB, 0, 144, src_of(w.DT, 0, 64), nullptr);
    job(mode == NEFES_MODE_FULL ? PL_DT : PL_DIR, b.GDT, 0, w.DT.ch, src_of(w.FIN), &dp);
    job(PL_FS, b.GFS, 0, 144, src_of(w.H[7]), nullptr);
  } else {
    job(PL_SIG, b.GSIG, 0, 16, src_of(w.H[7]), nullptr);
  }
  if (fused_trunk) {
    job(PL_T4, b.G[4], 0, 128, src_of(w.X), nullptr, 1);      // xyz columns of the skip layer (its h columns and bias: trunk launch)
  } else {
    for (int l = 7; l >= 1; --l) {
      if (l == 4) { ASrc h = src_of(w.H[3]); job(PL_T4, b.G[4], 0, 128, src_of(w.X), &h); }
      else job(PL_T0 + l, b.G[l], 0, 128, src_of(w.H[l - 1]), nullptr);
    }
    job(PL_T0, b.G[0], 0, 128, src_of(w.X), nullptr);
  }
  wa.n_jobs = nj;
  int64_t cost = 0;
  for (int j = 0; j < nj; ++j) { wa.job[j].cost_begin = cost; cost += (int64_t)T * wa.job[j].tile_bytes; }
  wa.total_cost = cost;
  wa.n_tiles = T;
  wa.ps = pack_src(net);
  wa.d_flat = dP;
  wa.off_ones = 0;
  wa.off_stage = 4096;
  wa.ring_bytes = 192 * 1024;
  const uint32_t smem = wa.off_stage + wa.ring_bytes;
  NEFES_REQUIRE(2 * r1k(max_tile_bytes + 4096) <= wa.ring_bytes, NEFES_EINVAL, "wgrad: tile too large for the ring");
  static bool attr_done = false;
  if (!attr_done) {
    NEFES_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAttr));
    attr_done = true;
  }
  int grid = num_sms();
  if (grid > T * nj) grid = T * nj;
  {
    double macs = 0;
    for (int j = 0; j < nj; ++j) macs += (double)wa.job[j].g_ch * wa.job[j].a_ch;
    prof_begin("wgrad", st, (double)cost, (double)M * 2.0 * macs);
  }
  wgrad_kernel<<<grid, kThreads, smem, st>>>(wa);
  prof_end(st);
  NEFES_CHECK_LAUNCH("wgrad");
  return NEFES_OK;
}

}  // namespace nefes
