// K6 raw2outputs_NeRFH_NFF (script/models/nerfh_nff.py:25-166): alpha compositing of rgb (3),
// feature (128), depth / disparity / opacity, NeRF-W transient colour and beta; forward and
// backward.  One CTA per ray.  Phase 1: one thread per sample -- alphas, transmittance as a
// warp-shuffle product scan in fp64 (torch's CPU cumprod accumulates in fp64), weights.
// Phase 2: one thread per channel -- coalesced sweep over the ray's [S, C] block of `raw`.
// TILES variants read / write the engine's tile-major raw blocks [C][128] (nefes_mlp_fwd_tiles): element (sample s,
// channel c) of a ray sits at  tile*C*128 + c*128 + row0 + s, so phase 2 runs one WARP per channel with the lanes on
// consecutive samples (128-byte coalesced) and a shuffle reduction.
#include "common.cuh"

namespace nefes {

constexpr int kMaxS = 256;
constexpr float kLastDelta = 1e2f;           // nerfh_nff.py:56

struct Scan {
  double warp_tot[8];
  float red[8];
};

// exclusive product scan over samples (thread s holds v_s); returns prod_{j<s} v_j.
__device__ __forceinline__ float excl_cumprod(float v, int S, Scan& sc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double p = (threadIdx.x < S) ? (double)v : 1.0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, p, o);
    if (lane >= o) p *= n;
  }
  if (lane == 31) sc.warp_tot[warp] = p;
  double excl = __shfl_up_sync(0xffffffffu, p, 1);
  if (lane == 0) excl = 1.0;
  __syncthreads();
  for (int w = 0; w < warp; ++w) excl *= sc.warp_tot[w];
  __syncthreads();
  return (float)excl;
}

// suffix sum: returns sum_{k>s} v_k (fp32).
__device__ __forceinline__ float excl_suffix_sum(float v, int S, Scan& sc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float p = (threadIdx.x < S) ? v : 0.f;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_down_sync(0xffffffffu, p, o);
    if (lane + o < 32) p += n;
  }
  if (lane == 0) sc.red[warp] = p;           // total of this warp
  float excl = __shfl_down_sync(0xffffffffu, p, 1);
  if (lane == 31) excl = 0.f;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = warp + 1; w < nw; ++w) excl += sc.red[w];
  __syncthreads();
  return excl;
}

__device__ __forceinline__ float block_total(float v, Scan& sc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sc.red[warp] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) s += sc.red[w];
  __syncthreads();
  return s;
}

template <int MODE> struct Chan;
template <> struct Chan<NEFES_COMP_SIGMA> { static constexpr int C = 1, SIG = 0; };
template <> struct Chan<NEFES_COMP_STATIC> { static constexpr int C = 132, SIG = 131; };
template <> struct Chan<NEFES_COMP_TRANSIENT> { static constexpr int C = 137, SIG = 131; };
template <> struct Chan<NEFES_COMP_TRANSIENT_STATIC_ONLY> { static constexpr int C = 137, SIG = 131; };

// per-sample quantities shared by forward and backward
struct Samp {
  float z, delta, a, a_s, a_t, T, Ts;        // Ts: static-only transmittance (mode 3)
};

template <int MODE>
__device__ __forceinline__ Samp sample_terms(const float* __restrict__ row, int cs, const float* __restrict__ zr,
                                             const float* __restrict__ noise_r, int s, int S, Scan& sc) {
  constexpr bool TR = (MODE == NEFES_COMP_TRANSIENT || MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY);
  Samp q = {};
  const bool ok = s < S;
  float sig_s = 0.f, sig_t = 0.f;
  if (ok) {
    q.z = zr[s];
    q.delta = (s + 1 < S) ? __fsub_rn(zr[s + 1], q.z) : kLastDelta;
    sig_s = row[Chan<MODE>::SIG * cs];
    if (TR) sig_t = row[135 * cs];
    if (!TR && noise_r != nullptr) sig_s = __fadd_rn(sig_s, noise_r[s]);
  }
  if (TR) {
    q.a_s = 1.f - expf(-(q.delta * sig_s));
    q.a_t = 1.f - expf(-(q.delta * sig_t));
    q.a = 1.f - expf(-(q.delta * __fadd_rn(sig_s, sig_t)));
  } else {
    q.a = 1.f - expf(-(q.delta * sig_s));
    q.a_s = q.a;
  }
  q.T = excl_cumprod(1.f - q.a, S, sc);
  if (MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) q.Ts = excl_cumprod(1.f - q.a_s, S, sc);
  return q;
}

template <int MODE, bool TILES>
__global__ void composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                     const float* __restrict__ noise, int S, float beta_min,
                                     nefes_comp_out_t o) {
  constexpr int C = Chan<MODE>::C;
  constexpr bool TR = (MODE == NEFES_COMP_TRANSIENT || MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY);
  __shared__ Scan sc;
  __shared__ float s_ws[kMaxS], s_wt[kMaxS];
  const int r = blockIdx.x, t = threadIdx.x;
  // element (sample s, channel c) = rr[s * rs + c * cs]
  const int rs = TILES ? 1 : C, cs = TILES ? 128 : 1;
  const float* rr = TILES ? raw + ((int64_t)r * S / 128) * C * 128 + ((int64_t)r * S) % 128 : raw + (int64_t)r * S * C;
  const float* zr = z + (int64_t)r * S;
  const Samp q = sample_terms<MODE>(rr + (int64_t)t * rs, cs, zr, noise ? noise + (int64_t)r * S : nullptr, t, S, sc);
  const bool ok = t < S;
  const float w = q.a * q.T;                                 // combined weights (nerfh_nff.py:77)
  float w_static, w_out;
  if (MODE == NEFES_COMP_TRANSIENT) { w_static = q.a_s * q.T; w_out = w; }
  else if (MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) { w_static = q.a_s * q.Ts; w_out = w_static; }
  else { w_static = w; w_out = w; }
  const float w_t = (MODE == NEFES_COMP_TRANSIENT) ? q.a_t * q.T : 0.f;
  if (ok) {
    s_ws[t] = w_static;
    s_wt[t] = w_t;
    o.weights[(int64_t)r * S + t] = w_out;
    if (TR && o.tsig != nullptr) o.tsig[(int64_t)r * S + t] = rr[(int64_t)t * rs + 135 * cs];
  }
  const float acc = block_total(ok ? w : 0.f, sc);           // acc_map = sum of combined weights
  if (t == 0) o.acc[r] = acc;
  if (MODE == NEFES_COMP_SIGMA) return;

  const float depth = block_total(ok ? w_out * q.z : 0.f, sc);
  const float wsum = (MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) ? block_total(ok ? w_out : 0.f, sc) : acc;
  float beta = 0.f;
  if (MODE == NEFES_COMP_TRANSIENT) beta = block_total(ok ? w_t * rr[(int64_t)t * rs + 136 * cs] : 0.f, sc) + beta_min;
  if (t == 0) {
    o.depth[r] = depth;
    o.disp[r] = 1.f / fmaxf(1e-10f, depth / wsum);
    o.beta[r] = beta;
  }
  __syncthreads();                                           // s_ws / s_wt visible
  if (TILES) {
    const int lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    if (S == 128 || S == 64) {
      // warp per channel, lane owns S/32 CONSECUTIVE samples (one 16- or 8-byte load per channel), eight channels in
      // flight per warp: the plain one-channel-at-a-time loop keeps ~1 load per warp outstanding and is latency-bound
      // at 1.6 TB/s (profiles/); the loads of a batch are independent and overlap
      const int spl = S >> 5;
      float w[4] = {0.f, 0.f, 0.f, 0.f}, wt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < spl) { w[i] = s_ws[lane * spl + i]; wt[i] = s_wt[lane * spl + i]; }
      const float* base = rr + lane * spl;
      for (int c0 = warp * 8; c0 < kHeadCh + 3; c0 += nw * 8) {      // virtual channels 131..133: transient colour
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + u;
          x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool live = c < kHeadCh || (MODE == NEFES_COMP_TRANSIENT && c < kHeadCh + 3);
          if (live) {
            const float* col = base + (int64_t)(c < kHeadCh ? c : c + 1) * 128;       // 131..133 -> raw channels 132..134
            if (spl == 4) x[u] = *reinterpret_cast<const float4*>(col);
            else { const float2 y = *reinterpret_cast<const float2*>(col); x[u].x = y.x; x[u].y = y.y; }
          }
        }
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool tr = (c0 + u) >= kHeadCh;
          v[u] = (tr ? wt[0] : w[0]) * x[u].x + (tr ? wt[1] : w[1]) * x[u].y + (tr ? wt[2] : w[2]) * x[u].z + (tr ? wt[3] : w[3]) * x[u].w;
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o2);
        }
        if (lane < 8) {
          float val = v[0];
#pragma unroll
          for (int u = 1; u < 8; ++u) if (lane == u) val = v[u];
          const int c = c0 + lane;
          if (c < 3) s_ws[kMaxS - 8 + c] = val;       // static colour: the transient part is added below
          else if (c < kHeadCh) o.feat[(int64_t)r * kFeat + (c - 3)] = val;
          else if (MODE == NEFES_COMP_TRANSIENT && c < kHeadCh + 3) s_wt[kMaxS - 8 + (c - kHeadCh)] = val;
        }
      }
      __syncthreads();
      if (t < 3) o.rgb[(int64_t)r * 3 + t] = s_ws[kMaxS - 8 + t] + (MODE == NEFES_COMP_TRANSIENT ? s_wt[kMaxS - 8 + t] : 0.f);
      return;
    }
    for (int c = warp; c < kHeadCh; c += nw) {
      const float* col = rr + (int64_t)c * 128;
      float v = 0.f;
      for (int s2 = lane; s2 < S; s2 += 32) v += s_ws[s2] * col[s2];
      if (MODE == NEFES_COMP_TRANSIENT && c < 3) {
        const float* colt = rr + (int64_t)(132 + c) * 128;
        for (int s2 = lane; s2 < S; s2 += 32) v += s_wt[s2] * colt[s2];
      }
      v = warp_sum(v);
      if (lane == 0) {
        if (c < 3) o.rgb[(int64_t)r * 3 + c] = v;
        else o.feat[(int64_t)r * kFeat + (c - 3)] = v;
      }
    }
    return;
  }
  for (int c = t; c < kHeadCh; c += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int s = 0;
    for (; s + 4 <= S; s += 4) {
      a0 += s_ws[s] * rr[(int64_t)s * C + c];
      a1 += s_ws[s + 1] * rr[(int64_t)(s + 1) * C + c];
      a2 += s_ws[s + 2] * rr[(int64_t)(s + 2) * C + c];
      a3 += s_ws[s + 3] * rr[(int64_t)(s + 3) * C + c];
    }
    for (; s < S; ++s) a0 += s_ws[s] * rr[(int64_t)s * C + c];
    float v = (a0 + a1) + (a2 + a3);
    if (MODE == NEFES_COMP_TRANSIENT && c < 3) {             // + transient colour (:128-150)
      float b0 = 0.f;
      for (int s2 = 0; s2 < S; ++s2) b0 += s_wt[s2] * rr[(int64_t)s2 * C + 132 + c];
      v += b0;
    }
    if (c < 3) o.rgb[(int64_t)r * 3 + c] = v;
    else o.feat[(int64_t)r * kFeat + (c - 3)] = v;
  }
}

template <int MODE, bool TILES>
__global__ void composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                     const float* __restrict__ noise, int S, nefes_comp_grad_t g,
                                     float* __restrict__ d_raw, float* __restrict__ compact) {
  constexpr int C = Chan<MODE>::C;
  __shared__ Scan sc;
  __shared__ float s_ws[kMaxS], s_wt[kMaxS], s_dsig[kMaxS], s_dsigt[kMaxS], s_dbeta[kMaxS];
  __shared__ float s_grgb[3];
  const int r = blockIdx.x, t = threadIdx.x;
  const int rs = TILES ? 1 : C, cs = TILES ? 128 : 1;
  const int64_t roff = TILES ? ((int64_t)r * S / 128) * C * 128 + ((int64_t)r * S) % 128 : (int64_t)r * S * C;
  const float* rr = raw + roff;
  const float* zr = z + (int64_t)r * S;
  const float* row = rr + (int64_t)t * rs;
  const Samp q = sample_terms<MODE>(row, cs, zr, noise ? noise + (int64_t)r * S : nullptr, t, S, sc);
  const bool ok = t < S;
  if (t < 3) s_grgb[t] = g.rgb ? g.rgb[(int64_t)r * 3 + t] : 0.f;

  const float w = q.a * q.T;
  float w_static, w_out;
  if (MODE == NEFES_COMP_TRANSIENT) { w_static = q.a_s * q.T; w_out = w; }
  else if (MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) { w_static = q.a_s * q.Ts; w_out = w_static; }
  else { w_static = w; w_out = w; }
  const float w_t = (MODE == NEFES_COMP_TRANSIENT) ? q.a_t * q.T : 0.f;

  // cotangents of the per-ray scalars, with disp chained onto depth and the weight sum
  const float acc = block_total(ok ? w : 0.f, sc);
  float g_depth = 0.f, g_wsum = 0.f;                         // on depth and on sum(w_out)
  float g_acc = g.acc ? g.acc[r] : 0.f;
  if (MODE != NEFES_COMP_SIGMA) {
    const float depth = block_total(ok ? w_out * q.z : 0.f, sc);
    const float wsum = (MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) ? block_total(ok ? w_out : 0.f, sc) : acc;
    g_depth = g.depth ? g.depth[r] : 0.f;
    const float ratio = depth / wsum;
    if (g.disp != nullptr && ratio > 1e-10f) {
      const float gq = -g.disp[r] / (ratio * ratio);
      g_depth += gq / wsum;
      g_wsum += -gq * depth / (wsum * wsum);
    }
  }
  __syncthreads();                                           // s_grgb
  float dsig = 0.f, dsigt = 0.f, dbeta = 0.f;
  {
    const float gw_out = ok ? (g.weights ? g.weights[(int64_t)r * S + t] : 0.f) : 0.f;
    float G_static = 0.f, G_t = 0.f, G_w = 0.f;              // d/d w_static, d/d w_t, d/d w (combined)
    const float gb = (MODE == NEFES_COMP_TRANSIENT && g.beta) ? g.beta[r] : 0.f;
    if (ok && MODE != NEFES_COMP_SIGMA) {
      G_static = s_grgb[0] * row[0] + s_grgb[1] * row[cs] + s_grgb[2] * row[2 * cs];
      if (MODE == NEFES_COMP_TRANSIENT)
        G_t = s_grgb[0] * row[132 * cs] + s_grgb[1] * row[133 * cs] + s_grgb[2] * row[134 * cs] + gb * row[136 * cs];
    }
    if (ok) {
      const float g_out = g_depth * q.z + g_wsum + gw_out;   // on w_out
      if (MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) { G_static += g_out; G_w = g_acc; }
      else if (MODE == NEFES_COMP_TRANSIENT) { G_w = g_out + g_acc; }
      else { G_static += g_out + g_acc; }                    // static / sigma-only: one weight vector
    }
    if (MODE == NEFES_COMP_TRANSIENT) {
      const float dT = G_static * q.a_s + G_t * q.a_t + G_w * q.a;
      const float suf = excl_suffix_sum(ok ? dT * q.T : 0.f, S, sc);
      const float common = G_w * q.T * (1.f - q.a) - suf;
      dsig = q.delta * (G_static * q.T * (1.f - q.a_s) + common);
      dsigt = q.delta * (G_t * q.T * (1.f - q.a_t) + common);
      dbeta = gb * w_t;
    } else if (MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) {
      const float suf_s = excl_suffix_sum(ok ? G_static * q.a_s * q.Ts : 0.f, S, sc);
      const float suf_c = excl_suffix_sum(ok ? G_w * q.a * q.T : 0.f, S, sc);
      const float common = G_w * q.T * (1.f - q.a) - suf_c;
      dsig = q.delta * (G_static * q.Ts * (1.f - q.a_s) - suf_s + common);
      dsigt = q.delta * common;
    } else {
      const float suf = excl_suffix_sum(ok ? G_static * q.a * q.T : 0.f, S, sc);
      dsig = q.delta * (G_static * q.T * (1.f - q.a) - suf);
    }
    if ((MODE == NEFES_COMP_TRANSIENT || MODE == NEFES_COMP_TRANSIENT_STATIC_ONLY) && ok && g.tsig)
      dsigt += g.tsig[(int64_t)r * S + t];
  }
  if (compact != nullptr) {
    // compact form of d_raw [ray][5][S]: d_raw[s, c] is rank one in (sample, channel) for the 131 + 3 colour / feature
    // channels -- (static weight | transient weight) x (cotangent of the ray's rgb / feature) -- so only the two weight
    // vectors and the three per-sample scalar gradients leave this kernel; the field backward rebuilds the columns
    if (ok) {
      float* cr = compact + (int64_t)r * 5 * S + t;
      cr[0] = w_static; cr[S] = w_t; cr[2 * S] = dsig; cr[3 * S] = dsigt; cr[4 * S] = dbeta;
    }
    return;
  }
  if (ok) {
    s_ws[t] = w_static; s_wt[t] = w_t; s_dsig[t] = dsig; s_dsigt[t] = dsigt; s_dbeta[t] = dbeta;
  }
  __syncthreads();
  float* dr = d_raw + roff;
  if (MODE == NEFES_COMP_SIGMA) {
    if (ok) dr[t] = dsig;
    return;
  }
  if (TILES) {
    const int lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    for (int c = warp; c < C; c += nw) {
      float gsel = 0.f;
      if (c < 3) gsel = s_grgb[c];
      else if (c < kHeadCh) gsel = g.feat ? g.feat[(int64_t)r * kFeat + (c - 3)] : 0.f;
      else if (c >= 132 && c < 135) gsel = s_grgb[c - 132];
      float* col = dr + (int64_t)c * 128;
      for (int s2 = lane; s2 < S; s2 += 32) {
        float v;
        if (c < kHeadCh) v = gsel * s_ws[s2];
        else if (c == 131) v = s_dsig[s2];
        else if (c < 135) v = gsel * s_wt[s2];
        else if (c == 135) v = s_dsigt[s2];
        else v = s_dbeta[s2];
        col[s2] = v;
      }
    }
    return;
  }
  for (int c = t; c < C; c += blockDim.x) {
    float gsel = 0.f;
    if (c < 3) gsel = s_grgb[c];
    else if (c < kHeadCh) gsel = g.feat ? g.feat[(int64_t)r * kFeat + (c - 3)] : 0.f;
    else if (c >= 132 && c < 135) gsel = s_grgb[c - 132];
    for (int s = 0; s < S; ++s) {
      float v;
      if (c < kHeadCh) v = gsel * s_ws[s];
      else if (c == 131) v = s_dsig[s];
      else if (c < 135) v = gsel * s_wt[s];
      else if (c == 135) v = s_dsigt[s];
      else v = s_dbeta[s];
      dr[(int64_t)s * C + c] = v;
    }
  }
}

}  // namespace nefes

extern "C" {

static int comp_check(const char* who, const float* raw, const float* z, int N, int S, int mode) {
  NEFES_REQUIRE(N == 0 || (raw && z), NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(N >= 0 && S >= 1 && S <= nefes::kMaxS, NEFES_EINVAL, "%s: need 1 <= S <= 256 (S=%d)", who, S);
  NEFES_REQUIRE(mode >= 0 && mode <= 3, NEFES_EINVAL, "%s: bad mode %d", who, mode);
  return NEFES_OK;
}

}  // extern "C"

template <bool TILES>
static int comp_fwd_launch(const char* who, const float* raw, const float* z_vals, const float* noise, int N, int S, int mode,
                           float beta_min, const nefes_comp_out_t* out_host, void* stream) {
  if (int e = comp_check(who, raw, z_vals, N, S, mode)) return e;
  if (N == 0) return NEFES_OK;
  NEFES_REQUIRE(out_host && out_host->acc && out_host->weights, NEFES_EINVAL, "%s: acc and weights outputs are required", who);
  if (mode != NEFES_COMP_SIGMA)
    NEFES_REQUIRE(out_host->rgb && out_host->feat && out_host->disp && out_host->depth && out_host->beta,
                  NEFES_EINVAL, "%s: missing output pointer", who);
  NEFES_REQUIRE(!TILES || 128 % S == 0, NEFES_EINVAL, "%s: tile-major raw needs S to divide 128 (S=%d)", who, S);
  const int threads = (int)nefes::round_up(S > 160 ? S : (mode == NEFES_COMP_SIGMA ? S : (TILES ? 256 : 160)), 32);
  cudaStream_t st = (cudaStream_t)stream;
  nefes_comp_out_t o = *out_host;
  {
    const int C = mode == NEFES_COMP_SIGMA ? 1 : (mode == NEFES_COMP_STATIC ? 132 : 137);
    nefes::prof_begin(mode == NEFES_COMP_STATIC ? "composite_fwd_coarse" : (mode == NEFES_COMP_SIGMA ? "composite_fwd_sigma" : "composite_fwd_fine"), st,
                      (double)N * (4.0 * S * C + 4.0 * S * 3 + 4.0 * 140), 0.0);
  }
  switch (mode) {
    case NEFES_COMP_SIGMA: nefes::composite_fwd_kernel<NEFES_COMP_SIGMA, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, beta_min, o); break;
    case NEFES_COMP_STATIC: nefes::composite_fwd_kernel<NEFES_COMP_STATIC, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, beta_min, o); break;
    case NEFES_COMP_TRANSIENT: nefes::composite_fwd_kernel<NEFES_COMP_TRANSIENT, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, beta_min, o); break;
    default: nefes::composite_fwd_kernel<NEFES_COMP_TRANSIENT_STATIC_ONLY, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, beta_min, o); break;
  }
  nefes::prof_end(st);
  NEFES_CHECK_LAUNCH(who);
  return NEFES_OK;
}

template <bool TILES>
static int comp_bwd_launch(const char* who, const float* raw, const float* z_vals, const float* noise, int N, int S, int mode,
                           const nefes_comp_grad_t* g_host, float* d_raw, void* stream, float* compact = nullptr) {
  if (int e = comp_check(who, raw, z_vals, N, S, mode)) return e;
  if (N == 0) return NEFES_OK;
  NEFES_REQUIRE(g_host && (d_raw || compact), NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(!TILES || 128 % S == 0, NEFES_EINVAL, "%s: tile-major raw needs S to divide 128 (S=%d)", who, S);
  const int threads = (int)nefes::round_up(S > 160 ? S : (mode == NEFES_COMP_SIGMA ? S : (TILES ? 256 : 160)), 32);
  cudaStream_t st = (cudaStream_t)stream;
  nefes_comp_grad_t g = *g_host;
  {
    const int C = mode == NEFES_COMP_SIGMA ? 1 : (mode == NEFES_COMP_STATIC ? 132 : 137);
    nefes::prof_begin(mode == NEFES_COMP_STATIC ? "composite_bwd_coarse" : (mode == NEFES_COMP_SIGMA ? "composite_bwd_sigma" : "composite_bwd_fine"), st,
                      (double)N * ((compact ? 4.0 * S * 8 : 8.0 * S * C) + 4.0 * S * 3 + 4.0 * 140), 0.0);   // raw in, d_raw out
  }
  switch (mode) {
    case NEFES_COMP_SIGMA: nefes::composite_bwd_kernel<NEFES_COMP_SIGMA, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, g, d_raw, compact); break;
    case NEFES_COMP_STATIC: nefes::composite_bwd_kernel<NEFES_COMP_STATIC, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, g, d_raw, compact); break;
    case NEFES_COMP_TRANSIENT: nefes::composite_bwd_kernel<NEFES_COMP_TRANSIENT, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, g, d_raw, compact); break;
    default: nefes::composite_bwd_kernel<NEFES_COMP_TRANSIENT_STATIC_ONLY, TILES><<<N, threads, 0, st>>>(raw, z_vals, noise, S, g, d_raw, compact); break;
  }
  nefes::prof_end(st);
  NEFES_CHECK_LAUNCH(who);
  return NEFES_OK;
}

extern "C" {

int nefes_composite_fwd(const float* raw, const float* z_vals, const float* noise, int N, int S,
                        int mode, float beta_min, const nefes_comp_out_t* out_host, void* stream) {
  return comp_fwd_launch<false>("nefes_composite_fwd", raw, z_vals, noise, N, S, mode, beta_min, out_host, stream);
}
int nefes_composite_fwd_tiles(const float* raw_tiles, const float* z_vals, const float* noise, int N, int S,
                              int mode, float beta_min, const nefes_comp_out_t* out_host, void* stream) {
  return comp_fwd_launch<true>("nefes_composite_fwd_tiles", raw_tiles, z_vals, noise, N, S, mode, beta_min, out_host, stream);
}
int nefes_composite_bwd(const float* raw, const float* z_vals, const float* noise, int N, int S,
                        int mode, const nefes_comp_grad_t* g_host, float* d_raw, void* stream) {
  return comp_bwd_launch<false>("nefes_composite_bwd", raw, z_vals, noise, N, S, mode, g_host, d_raw, stream);
}
int nefes_composite_bwd_tiles(const float* raw_tiles, const float* z_vals, const float* noise, int N, int S,
                              int mode, const nefes_comp_grad_t* g_host, float* d_raw_tiles, void* stream) {
  return comp_bwd_launch<true>("nefes_composite_bwd_tiles", raw_tiles, z_vals, noise, N, S, mode, g_host, d_raw_tiles, stream);
}

int nefes_composite_bwd_compact(const float* raw_tiles, const float* z_vals, const float* noise, int N, int S,
                                int mode, const nefes_comp_grad_t* g_host, float* compact, void* stream) {
  NEFES_REQUIRE(compact != nullptr && mode != NEFES_COMP_SIGMA, NEFES_EINVAL, "nefes_composite_bwd_compact: bad arguments");
  return comp_bwd_launch<true>("nefes_composite_bwd_compact", raw_tiles, z_vals, noise, N, S, mode, g_host, nullptr, stream, compact);
}

}  // extern "C"
