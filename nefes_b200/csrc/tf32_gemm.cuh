// tf32 tensor-core GEMM behind the fp32 path's layer-at-a-time structure (NEFES_PREC_TF32): the same three products as
// sgemm.cuh -- forward C = act(A W^T + b), data gradient dA = mask(dD W), weight gradient dW += dD^T A -- with the fp32
// activations staying in HBM as they are, but the multiply-accumulate on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32 (10-bit mantissa operands, fp32 accumulation in tensor memory).  Operands are ROUNDED to tf32
// (cvt.rna) while they are staged -- the instruction itself would truncate, and a one-sided 2^-11 bias per operand adds up
// over K = 128 products.  Measured on the oracle with the same operand rounding: worst output 3.6e-4 of scale against the
// fp32 reference on random-init fields (bf16 operands: 3.1e-3), i.e. inside the north-star's 1e-3 with a tensor-core path.
//
// One CTA = a 128 x Jt output tile (Jt <= 256 columns = one MMA N), the reduction in chunks of 32 with two shared-memory
// stages: all 8 warps stage chunk k+1 (global fp32 -> registers -> cvt.rna.tf32 -> the UMMA K-major no-swizzle layout,
// 8 x 16-byte core matrices, [chunk column][row][4 floats]) while the tensor pipe works on chunk k; warp 0 issues (converged
// warp, elected lane), warps 0-3 run the epilogue (thread = output row).  The loaders address A and B through the same
// "reduction-contiguous" flags as sgemm.cuh, so operands that need a transpose (dgrad's W, wgrad's dD^T and A^T) get it on
// the way into shared memory and every MMA operand is K-major.
// script/models/nerfh_nff.py:525-576 (the Linear layers), :356-418 (FusionNet's convolutions as im2col GEMMs).
#pragma once
#include "sgemm.cuh"
#include "tc05.cuh"

namespace nefes {

using namespace tc05;

constexpr int kTfRK = 32;                               // reduction elements per stage
constexpr uint32_t kTfLboA = 128 * 16 + 16;             // chunk-column stride of the A stage (16 bytes of padding: bank spread)
constexpr uint32_t kTfStageA = (kTfRK / 4) * kTfLboA;   // 16 512 B
__host__ __device__ constexpr uint32_t tf_lbo_b(int jt) { return (uint32_t)jt * 16u + 16u; }
__host__ __device__ constexpr uint32_t tf_stage_b(int jt) { return (kTfRK / 4) * tf_lbo_b(jt); }
__host__ __device__ constexpr uint32_t tf_smem(int jt) {       // two operand stages, re-used by the epilogue as a [128][jt + 4] fp32 tile
  return (2u * (kTfStageA + tf_stage_b(jt)) > 128u * (uint32_t)(jt + 4) * 4u ? 2u * (kTfStageA + tf_stage_b(jt)) : 128u * (uint32_t)(jt + 4) * 4u) + 128u;
}

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {      // tf32 x tf32 -> fp32, both operands K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// stage `rows` x 32 reduction elements of an operand: element (x, r) = RC ? P[(x0 + x) * ld + r] : P[r * ld + (x0 + x)].
// Two halves so that the global loads of chunk k+1 are in flight (in registers) while the tensor pipe works on chunk k:
// tf_fetch issues the loads, tf_put rounds to tf32 and writes the UMMA layout.  NS = slots per thread (rows * 8 / 256).
template <bool RC, int NS>
__device__ __forceinline__ void tf_fetch(const float* __restrict__ P, int64_t ld, int64_t x0, int64_t x_end, int64_t r0, int64_t r_end,
                                         int rows, float4 (&v)[NS]) {
  const int slots = rows * (kTfRK / 4);
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const int idx = threadIdx.x + n * 256;
    v[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx >= slots) continue;
    int x, c;
    if (RC) { c = idx & 7; x = idx >> 3; }              // 8 consecutive threads read 128 contiguous bytes of one row
    else { x = idx % rows; c = idx / rows; }            // consecutive threads read consecutive x of one reduction row
    const int64_t gx = x0 + x, r = r0 + c * 4;
    if (gx >= x_end) continue;
    if (RC) {
      const float* p = P + gx * ld + r;
      if (r + 3 < r_end && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) v[n] = __ldg(reinterpret_cast<const float4*>(p));
      else {
        if (r < r_end) v[n].x = __ldg(p);
        if (r + 1 < r_end) v[n].y = __ldg(p + 1);
        if (r + 2 < r_end) v[n].z = __ldg(p + 2);
        if (r + 3 < r_end) v[n].w = __ldg(p + 3);
      }
    } else {
      const float* p = P + r * ld + gx;
      if (r < r_end) v[n].x = __ldg(p);
      if (r + 1 < r_end) v[n].y = __ldg(p + ld);
      if (r + 2 < r_end) v[n].z = __ldg(p + 2 * ld);
      if (r + 3 < r_end) v[n].w = __ldg(p + 3 * ld);
    }
  }
}
template <bool RC, int NS>
__device__ __forceinline__ void tf_put(const float4 (&v)[NS], int rows, uint32_t lbo, uint8_t* dst) {
  const int slots = rows * (kTfRK / 4);
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const int idx = threadIdx.x + n * 256;
    if (idx >= slots) continue;
    int x, c;
    if (RC) { c = idx & 7; x = idx >> 3; }
    else { x = idx % rows; c = idx / rows; }
    *reinterpret_cast<float4*>(dst + c * lbo + x * 16) = make_float4(to_tf32(v[n].x), to_tf32(v[n].y), to_tf32(v[n].z), to_tf32(v[n].w));
  }
}

// Per-slot state computed ONCE per CTA (warp sampling: the generic fetch / put above re-derived row, chunk column, 64-bit
// address, bounds and alignment for every slot of every chunk -- the kernel was 61 % issue-bound): the global pointer of the
// slot in chunk 0 (null: row out of range -> zeros), whether it can be read as one 16-byte load, its byte offset in a stage.
template <int NS>
struct TfSlots { const float* p[NS]; uint32_t soff[NS]; bool vec[NS], live[NS]; };
template <bool RC, int NS>
__device__ __forceinline__ void tf_slots(const float* __restrict__ P, int64_t ld, int64_t x0, int64_t x_end, int64_t r_begin, int rows,
                                         uint32_t lbo, TfSlots<NS>& S) {
  const int slots = rows * (kTfRK / 4);
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const int idx = threadIdx.x + n * 256;
    int x = 0, c = 0;
    S.live[n] = idx < slots;
    if (S.live[n]) {
      if (RC) { c = idx & 7; x = idx >> 3; }
      else { x = idx % rows; c = idx / rows; }
    }
    S.soff[n] = (uint32_t)c * lbo + (uint32_t)x * 16u;
    const bool in = S.live[n] && x0 + x < x_end;
    S.p[n] = in ? (RC ? P + (x0 + x) * ld + r_begin + c * 4 : P + (r_begin + c * 4) * ld + (x0 + x)) : nullptr;
    S.vec[n] = RC && in && ((reinterpret_cast<uintptr_t>(S.p[n]) & 15) == 0) && ((ld & 3) == 0);
  }
}
// a FULL chunk (all 32 reduction elements inside the range): chunk kc of the slots
template <bool RC, int NS>
__device__ __forceinline__ void tf_fetch_fast(const TfSlots<NS>& S, int64_t adv, int64_t ld, float4 (&v)[NS]) {
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    v[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (S.p[n] == nullptr) continue;
    const float* p = S.p[n] + adv;
    if (RC) {
      if (S.vec[n]) v[n] = __ldg(reinterpret_cast<const float4*>(p));
      else v[n] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
    } else {
      v[n] = make_float4(__ldg(p), __ldg(p + ld), __ldg(p + 2 * ld), __ldg(p + 3 * ld));
    }
  }
}
template <int NS>
__device__ __forceinline__ void tf_put_fast(const TfSlots<NS>& S, const float4 (&v)[NS], uint8_t* dst) {
#pragma unroll
  for (int n = 0; n < NS; ++n)
    if (S.live[n])
      *reinterpret_cast<float4*>(dst + S.soff[n]) = make_float4(to_tf32(v[n].x), to_tf32(v[n].y), to_tf32(v[n].z), to_tf32(v[n].w));
}

template <bool A_RC, bool B_RC, int NSB>              // NSB = B slots per thread: 4 (Jt <= 128) or 8 (Jt <= 256)
__global__ void __launch_bounds__(256) tf32_gemm_kernel(const GemmArgs g, int jt_max) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.x * 128;
  const int j0 = blockIdx.y * jt_max;
  int jt = g.J - j0;
  jt = jt > jt_max ? jt_max : ((jt + 15) & ~15);        // MMA N: a multiple of 16, rows beyond J are staged as zeros
  int64_t r_begin = 0, r_end = g.R;
  if (g.r_chunk > 0) { r_begin = (int64_t)blockIdx.z * g.r_chunk; r_end = min(g.R, r_begin + g.r_chunk); }
  const uint32_t lbo_b = tf_lbo_b(jt_max), stage_b = tf_stage_b(jt_max);
  uint8_t* sA[2] = {smem, smem + kTfStageA};
  uint8_t* sB[2] = {smem + 2 * kTfStageA, smem + 2 * kTfStageA + stage_b};
  const int tcols = jt_max <= 32 ? 32 : (jt_max <= 64 ? 64 : (jt_max <= 128 ? 128 : 256));
  if (threadIdx.x == 0) { mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1); fence_mbar_init(); }
  if (warp == 0) {
    if (tcols == 32) tmem_alloc<32>(&tmem_slot);
    else if (tcols == 64) tmem_alloc<64>(&tmem_slot);
    else if (tcols == 128) tmem_alloc<128>(&tmem_slot);
    else tmem_alloc<256>(&tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = idesc_tf32(128, jt);
  const int n_chunks = (int)((r_end - r_begin + kTfRK - 1) / kTfRK);
  const bool leader = warp == 0 ? elect_one() : false;

  float4 va[4], vb[NSB];                                // the next chunk, in flight (registers)
  TfSlots<4> SA;
  TfSlots<NSB> SB;
  tf_slots<A_RC, 4>(g.A, g.lda, i0, g.I, r_begin, 128, kTfLboA, SA);
  tf_slots<B_RC, NSB>(g.B, g.ldb, j0, g.J, r_begin, jt, lbo_b, SB);
  const int64_t adv_a = A_RC ? kTfRK : kTfRK * g.lda, adv_b = B_RC ? kTfRK : kTfRK * g.ldb;
  auto fetch = [&](int kc) {
    const int64_t r0 = r_begin + (int64_t)kc * kTfRK;
    if (r0 + kTfRK <= r_end) {
      tf_fetch_fast<A_RC, 4>(SA, kc * adv_a, g.lda, va);
      tf_fetch_fast<B_RC, NSB>(SB, kc * adv_b, g.ldb, vb);
    } else {                                                             // the ragged last chunk: bounds-checked path
      tf_fetch<A_RC, 4>(g.A, g.lda, i0, g.I, r0, r_end, 128, va);
      tf_fetch<B_RC, NSB>(g.B, g.ldb, j0, g.J, r0, r_end, jt, vb);
    }
  };
  if (n_chunks > 0) fetch(0);
  for (int kc = 0; kc < n_chunks; ++kc) {
    const int s = kc & 1;
    if (kc >= 2) mbar_wait(&bar_mma[s], ((kc >> 1) - 1) & 1);           // the MMAs that read this stage two chunks ago retired
    tf_put_fast<4>(SA, va, sA[s]);
    tf_put_fast<NSB>(SB, vb, sB[s]);
    if (kc + 1 < n_chunks) fetch(kc + 1);                                // loads of the next chunk fly during the MMAs of this one
    fence_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint32_t a0 = uni(smem_u32(sA[s])), b0 = uni(smem_u32(sB[s]));
      if (leader) {
#pragma unroll
        for (int m = 0; m < kTfRK / 8; ++m)                              // K = 8 per MMA = two chunk columns
          mma_ss_tf32(tmem, smem_desc(a0 + 2 * m * kTfLboA, kTfLboA, 128), smem_desc(b0 + 2 * m * lbo_b, lbo_b, 128), idesc,
                      (kc > 0 || m > 0) ? 1u : 0u);
        mma_commit(&bar_mma[s]);
      }
      __syncwarp();
    }
  }
  // every chunk's commit completed => the accumulator is final (commits retire in order: wait for the last one)
  if (n_chunks > 0) {
    const int last = n_chunks - 1;
    mbar_wait(&bar_mma[last & 1], (last >> 1) & 1);
  }
  tc_fence_after();
  // epilogue: tensor memory -> shared memory (the stages are dead; row-major [128][jt + 1]) -> coalesced global access:
  // a warp handles one output row at a time, 32 consecutive columns per instruction
  float* sC = reinterpret_cast<float*>(smem);
  const int ldsc = jt + 4;                               // rows 16-byte aligned; 128-bit accesses spread over all banks
  if (warp < 4 && n_chunks > 0) {
    const int row = warp * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < jt; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        *reinterpret_cast<uint4*>(sC + row * ldsc + c0 + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (n_chunks > 0) {
    // the per-element conditions are uniform over the launch: three specialised loops (warp-sampling showed the generic one
    // -- atomic? accumulate? bias? activation switch? mask? per element -- taking 45 % of the kernel's samples)
    const int n_cols = min(jt, g.J - j0);
    if (g.atomic) {
      for (int row = warp; row < 128 && i0 + row < g.I; row += 8)
        for (int c = lane; c < n_cols; c += 32) atomicAdd(g.C + (i0 + row) * g.ldc + j0 + c, sC[row * ldsc + c]);
    } else if (!g.accumulate && (g.act == ACT_NONE || g.act == ACT_RELU)) {
      const bool relu = g.act == ACT_RELU;
      const bool vec = (n_cols & 3) == 0 && (g.ldc & 3) == 0 && (j0 & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                       (g.mask == nullptr || ((g.ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0)) &&
                       (g.bias == nullptr || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
      for (int row = warp; row < 128 && i0 + row < g.I; row += 8) {
        const int64_t i = i0 + row;
        if (vec) {
          for (int c = lane * 4; c < n_cols; c += 128) {
            float4 x = *reinterpret_cast<const float4*>(sC + row * ldsc + c);
            if (g.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + j0 + c)); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
            if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            if (g.mask) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(g.mask + i * g.ldm + j0 + c));
              x.x = m.x > 0.f ? x.x : 0.f; x.y = m.y > 0.f ? x.y : 0.f; x.z = m.z > 0.f ? x.z : 0.f; x.w = m.w > 0.f ? x.w : 0.f;
            }
            *reinterpret_cast<float4*>(g.C + i * g.ldc + j0 + c) = x;
          }
        } else {
          for (int c = lane; c < n_cols; c += 32) {
            float x = sC[row * ldsc + c];
            if (g.bias) x += __ldg(g.bias + j0 + c);
            if (relu) x = fmaxf(x, 0.f);
            if (g.mask) x = (__ldg(g.mask + i * g.ldm + j0 + c) > 0.f) ? x : 0.f;
            g.C[i * g.ldc + j0 + c] = x;
          }
        }
      }
    } else {
      for (int row = warp; row < 128 && i0 + row < g.I; row += 8) {
        const int64_t i = i0 + row;
        for (int c = lane; c < n_cols; c += 32) {
          const int j = j0 + c;
          float x = sC[row * ldsc + c];
          float* dst = g.C + i * g.ldc + j;
          if (g.accumulate) x += *dst;
          if (g.bias) x += g.bias[j];
          switch (g.act) {
            case ACT_RELU: x = fmaxf(x, 0.f); break;
            case ACT_SOFTPLUS: x = softplus_f(x); break;
            case ACT_SIGMOID: x = sigmoid_f(x); break;
            case ACT_THEADS: x = (j < 3) ? sigmoid_f(x) : softplus_f(x); break;
            default: break;
          }
          if (g.mask) x = (g.mask[i * g.ldm + j] > 0.f) ? x : 0.f;
          *dst = x;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    if (tcols == 32) tmem_dealloc<32>(tmem);
    else if (tcols == 64) tmem_dealloc<64>(tmem);
    else if (tcols == 128) tmem_dealloc<128>(tmem);
    else tmem_dealloc<256>(tmem);
  }
}

template <bool A_RC, bool B_RC>
inline int launch_tf32_gemm(const GemmArgs& g, cudaStream_t st, const char* what) {
  if (g.I <= 0 || g.J <= 0 || g.R <= 0) return NEFES_OK;
  int jt = g.J >= 256 ? 256 : ((g.J + 15) & ~15);
  if (jt < 32) jt = 32;                                                   // smallest tensor-memory allocation
  const unsigned gz = g.r_chunk > 0 ? (unsigned)ceil_div(g.R, g.r_chunk) : 1u;
  dim3 grid((unsigned)ceil_div(g.I, 128), (unsigned)ceil_div(g.J, jt), gz);
  static bool attr_done = false;
  if (!attr_done) {
    NEFES_CUDA(cudaFuncSetAttribute(tf32_gemm_kernel<A_RC, B_RC, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf_smem(128)));
    NEFES_CUDA(cudaFuncSetAttribute(tf32_gemm_kernel<A_RC, B_RC, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf_smem(256)));
    attr_done = true;
  }
  if (jt <= 128) tf32_gemm_kernel<A_RC, B_RC, 4><<<grid, 256, tf_smem(jt), st>>>(g, jt);
  else tf32_gemm_kernel<A_RC, B_RC, 8><<<grid, 256, tf_smem(jt), st>>>(g, jt);
  NEFES_CHECK_LAUNCH(what);
  return NEFES_OK;
}

// ---- the three products of a Linear layer, on the SIMT fp32 GEMM or on the tf32 tensor-core GEMM --------------------------
// gemm_tf32(): process-wide switch, set by the NEFES_PREC_TF32 entry points (and nefes_gemm_mode) around their calls.
inline int& gemm_tf32() { static int on = 0; return on; }
struct GemmTf32Scope {
  int prev;
  explicit GemmTf32Scope(int on) : prev(gemm_tf32()) { gemm_tf32() = on; }
  ~GemmTf32Scope() { gemm_tf32() = prev; }
};

// C[M,N] = act( (accumulate ? C : 0) + A[M,K] W[N,K]^T + bias )
inline int linear_fwd(cudaStream_t st, const float* A, int64_t lda, const float* W, int64_t ldw,
                      const float* bias, float* C, int64_t ldc, int64_t M, int N, int K, int act,
                      int accumulate) {
  GemmArgs g = {A, lda, W, ldw, C, ldc, bias, nullptr, 0, M, N, K, act, accumulate, 0, 0};
  if (gemm_tf32()) return launch_tf32_gemm<true, true>(g, st, "linear_fwd(tf32)");
  return launch_sgemm<true, true>(g, st, "linear_fwd");
}
// dA[M,K] = relu_mask( (accumulate ? dA : 0) + dD[M,N] W[N,K] )
inline int linear_dgrad(cudaStream_t st, const float* dD, int64_t ldd, const float* W, int64_t ldw,
                        float* dA, int64_t lda, int64_t M, int N, int K, const float* mask,
                        int64_t ldm, int accumulate) {
  GemmArgs g = {dD, ldd, W, ldw, dA, lda, nullptr, mask, ldm, M, K, N, ACT_NONE, accumulate, 0, 0};
  if (gemm_tf32()) return launch_tf32_gemm<true, false>(g, st, "linear_dgrad(tf32)");
  return launch_sgemm<true, false>(g, st, "linear_dgrad");
}
// dW[N,K] += dD[M,N]^T A[M,K]      (split over M, fp32 atomics)
inline int linear_wgrad(cudaStream_t st, const float* dD, int64_t ldd, const float* A, int64_t lda,
                        float* dW, int64_t ldw, int64_t M, int N, int K) {
  int64_t chunk = round_up(ceil_div(M, 592), 16);      // ~4 CTAs per SM worth of splits
  if (chunk < 512) chunk = 512;
  if (gemm_tf32()) chunk = round_up(chunk, kTfRK);
  GemmArgs g = {dD, ldd, A, lda, dW, ldw, nullptr, nullptr, 0, N, K, M, ACT_NONE, 0, 1, chunk};
  if (gemm_tf32()) return launch_tf32_gemm<false, false>(g, st, "linear_wgrad(tf32)");
  return launch_sgemm<false, false>(g, st, "linear_wgrad");
}

}  // namespace nefes
