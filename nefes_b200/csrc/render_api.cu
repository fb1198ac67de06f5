// Whole-path entry points: render_rays (script/models/rendering.py:68-180) forward and backward as ONE C call each.
// The call strings the stage kernels (K2 sample_coarse, K5 field, K6 compositing, K3 sample_fine, K5, K6) together on the
// caller's stream with every intermediate -- sample points, raw in tile-major layout, saved activations, compact
// cotangents -- living in two caller-provided workspaces, so nothing but the per-ray results crosses the boundary.
#include "common.cuh"
#include <stdlib.h>

namespace nefes {
int mlp_prepack_bf16(const float* P, int net, int mode, int64_t N, int S, void* saved, cudaStream_t st);   // mlp_tc.cu

// pts[i,s,:] = o_i + d_i * z[i,s] (rendering.py:114/:143: a multiply then an add, no FMA, so the points are the
// reference's bit for bit); optionally the contiguous copy of the view directions and, for the fine pass,
// z_std = population standard deviation of the importance samples (rendering.py:162).  One warp per ray.
__global__ void ray_points_kernel(const float* __restrict__ rays, int ld, const float* __restrict__ z, int N, int S,
                                  float* __restrict__ pts, float* __restrict__ dirs, const float* __restrict__ zs, int ns,
                                  float* __restrict__ z_std) {
  const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ray >= N) return;
  const float* r = rays + (int64_t)ray * ld;
  const float o0 = r[0], o1 = r[1], o2 = r[2], d0 = r[3], d1 = r[4], d2 = r[5];
  const float* zr = z + (int64_t)ray * S;
  float* p = pts + (int64_t)ray * S * 3;
  for (int e = lane; e < S * 3; e += 32) {
    const int s = e / 3, k = e - 3 * s;
    const float o = k == 0 ? o0 : (k == 1 ? o1 : o2), d = k == 0 ? d0 : (k == 1 ? d1 : d2);
    p[e] = __fadd_rn(o, __fmul_rn(d, zr[s]));
  }
  if (dirs != nullptr && lane < 3) dirs[(int64_t)ray * 3 + lane] = r[8 + lane];
  if (z_std != nullptr) {
    const float* q = zs + (int64_t)ray * ns;
    float sum = 0.f;
    for (int s = lane; s < ns; s += 32) sum += q[s];
    const float mean = warp_sum(sum) / (float)ns;
    float var = 0.f;
    for (int s = lane; s < ns; s += 32) { const float t = q[s] - mean; var += t * t; }
    var = warp_sum(var);
    if (lane == 0) z_std[ray] = sqrtf(var / (float)ns);
  }
}

// Cotangent of a ray_batch row from the cotangents of its sample points and view direction:
// d_o = sum_s d_pts, d_d = sum_s z_s d_pts (both passes), d_viewdirs = d_dirs (both passes); every other column 0.
__global__ void ray_points_bwd_kernel(const float* __restrict__ dpc, const float* __restrict__ zc, int Sc,
                                      const float* __restrict__ dpf, const float* __restrict__ zf, int Sf,
                                      const float* __restrict__ ddc, const float* __restrict__ ddf, int N, int ld,
                                      float* __restrict__ d_rays) {
  const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ray >= N) return;
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const float* dp = pass == 0 ? dpc : dpf;
    const float* z = pass == 0 ? zc : zf;
    const int S = pass == 0 ? Sc : Sf;
    if (dp == nullptr) continue;
    dp += (int64_t)ray * S * 3;
    z += (int64_t)ray * S;
    for (int s = lane; s < S; s += 32) {
      const float zz = z[s], g0 = dp[3 * s], g1 = dp[3 * s + 1], g2 = dp[3 * s + 2];
      a[0] += g0; a[1] += g1; a[2] += g2;
      a[3] += zz * g0; a[4] += zz * g1; a[5] += zz * g2;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) a[k] = warp_sum(a[k]);
  float* out = d_rays + (int64_t)ray * ld;
  for (int c = lane; c < ld; c += 32) {
    float v = 0.f;
    if (c < 6) {
#pragma unroll
      for (int k = 0; k < 6; ++k) if (c == k) v = a[k];
    } else if (c >= 8 && c < 11) {
      if (ddc != nullptr) v += ddc[(int64_t)ray * 3 + c - 8];
      if (ddf != nullptr) v += ddf[(int64_t)ray * 3 + c - 8];
    }
    out[c] = v;
  }
}

// sub-buffer alignment inside the workspaces (NEFES_WS_ALIGN overrides, for placement experiments)
static inline int64_t al(int64_t x) {
  static const int64_t a = [] { const char* e = getenv("NEFES_WS_ALIGN"); const int64_t v = e ? atoll(e) : 0; return v >= 256 ? v : 256; }();
  return round_up(x, a);
}

// Where everything lives.  `keep` survives from the forward to the backward call; `scratch` is per call.
struct RenderPlan {
  int Sc, Sf, mode_c, mode_f, comp_c, comp_f, Cc, Cf;
  bool fine, tiled_c, tiled_f;
  int64_t pts_c, pts_f, dirs, raw_c, raw_f, saved_c, saved_f, keep_bytes;
  int64_t fwd_scratch;
  int64_t cg_c, cg_f, draw, dpts_c, dpts_f, ddirs_c, ddirs_f, mlp_scratch, bwd_scratch;
};

static int raw_channels(int mode) { return mode == NEFES_MODE_SIGMA ? 1 : (mode == NEFES_MODE_STATIC ? 132 : 137); }

static int make_plan(const char* who, const nefes_render_cfg_t* cfg, int64_t N, RenderPlan* P) {
  NEFES_REQUIRE(cfg != nullptr, NEFES_EINVAL, "%s: null config", who);
  NEFES_REQUIRE(N >= 0 && N <= (1 << 24), NEFES_EINVAL, "%s: bad ray count %lld", who, (long long)N);
  NEFES_REQUIRE(cfg->n_samples >= 2 && cfg->n_samples <= 256 && cfg->n_importance >= 0 && cfg->n_importance <= 256 &&
                cfg->n_samples + cfg->n_importance <= 256, NEFES_EINVAL, "%s: bad sample counts %d + %d", who,
                cfg->n_samples, cfg->n_importance);
  NEFES_REQUIRE(cfg->prec == NEFES_PREC_FP32 || cfg->prec == NEFES_PREC_BF16 || cfg->prec == NEFES_PREC_TF32, NEFES_EINVAL, "%s: bad precision", who);
  RenderPlan& p = *P;
  p.fine = cfg->n_importance > 0;
  p.Sc = cfg->n_samples;
  p.Sf = cfg->n_samples + cfg->n_importance;
  // run_network_NeRFH_NFF's three cases (nerfh_nff.py:192-231) and raw2outputs' branches (:83-89, :92-150)
  const bool sigma_only = cfg->test_time && p.fine;                        // store_rgb = (N_importance == 0)
  p.mode_c = sigma_only ? NEFES_MODE_SIGMA : NEFES_MODE_STATIC;
  p.comp_c = sigma_only ? NEFES_COMP_SIGMA : NEFES_COMP_STATIC;
  p.mode_f = cfg->output_transient ? NEFES_MODE_FULL : NEFES_MODE_STATIC;
  p.comp_f = !cfg->output_transient ? NEFES_COMP_STATIC
             : ((cfg->test_time && !cfg->transient_at_test) ? NEFES_COMP_TRANSIENT_STATIC_ONLY : NEFES_COMP_TRANSIENT);
  p.Cc = raw_channels(p.mode_c);
  p.Cf = raw_channels(p.mode_f);
  p.tiled_c = cfg->prec == NEFES_PREC_BF16 && p.mode_c != NEFES_MODE_SIGMA && 128 % p.Sc == 0;
  p.tiled_f = cfg->prec == NEFES_PREC_BF16 && 128 % p.Sf == 0;
  const int64_t Nn = N > 0 ? N : 1, Mc = Nn * p.Sc, Mf = Nn * p.Sf;
  int64_t sv_c = 0, sf_c = 0, sb_c = 0, sv_f = 0, sf_f = 0, sb_f = 0;
  if (int e = nefes_mlp_workspace(cfg->net_coarse, p.mode_c, cfg->prec, Mc, Nn, &sv_c, &sf_c, &sb_c)) return e;
  if (p.fine)
    if (int e = nefes_mlp_workspace(cfg->net_fine, p.mode_f, cfg->prec, Mf, Nn, &sv_f, &sf_f, &sb_f)) return e;
  auto raw_bytes = [](bool tiled, int64_t M, int C) { return (tiled ? ceil_div(M, 128) * 128 : M) * C * 4; };
  int64_t o = 0;
  p.pts_c = o; o += al(Mc * 12);
  p.dirs = o; o += al(Nn * 12);
  p.raw_c = o; o += al(raw_bytes(p.tiled_c, Mc, p.Cc));
  p.saved_c = o; o += al(sv_c);
  if (p.fine) {
    p.pts_f = o; o += al(Mf * 12);
    p.raw_f = o; o += al(raw_bytes(p.tiled_f, Mf, p.Cf));
    p.saved_f = o; o += al(sv_f);
  } else {
    p.pts_f = p.raw_f = p.saved_f = 0;
  }
  p.keep_bytes = o;
  p.fwd_scratch = al(sf_c > sf_f ? sf_c : sf_f);
  o = 0;
  p.cg_c = o; o += al(Nn * 5 * p.Sc * 4);
  p.cg_f = o; o += al(Nn * 5 * p.Sf * 4);
  const int64_t dr_c = p.tiled_c ? 0 : raw_bytes(false, Mc, p.Cc), dr_f = (!p.fine || p.tiled_f) ? 0 : raw_bytes(false, Mf, p.Cf);
  p.draw = o; o += al(dr_c > dr_f ? dr_c : dr_f);
  p.dpts_c = o; o += al(Mc * 12);
  p.dpts_f = o; o += al(p.fine ? Mf * 12 : 0);
  p.ddirs_c = o; o += al(Nn * 12);
  p.ddirs_f = o; o += al(Nn * 12);
  p.mlp_scratch = o; o += al(sb_c > sb_f ? sb_c : sb_f);
  p.bwd_scratch = o;
  return NEFES_OK;
}

static bool any_grad(const nefes_comp_grad_t* g) {
  return g && (g->rgb || g->feat || g->disp || g->acc || g->weights || g->depth || g->beta || g->tsig);
}

}  // namespace nefes

extern "C" {

int nefes_render_rays_workspace(const nefes_render_cfg_t* cfg, int64_t N, int64_t* keep_bytes_host,
                                int64_t* scratch_fwd_bytes_host, int64_t* scratch_bwd_bytes_host) {
  NEFES_REQUIRE(keep_bytes_host && scratch_fwd_bytes_host && scratch_bwd_bytes_host, NEFES_EINVAL,
                "nefes_render_rays_workspace: null output");
  nefes::RenderPlan P;
  if (int e = nefes::make_plan("nefes_render_rays_workspace", cfg, N, &P)) return e;
  *keep_bytes_host = P.keep_bytes;
  *scratch_fwd_bytes_host = P.fwd_scratch;
  *scratch_bwd_bytes_host = P.bwd_scratch;
  return NEFES_OK;
}

int nefes_render_rays_prepack(const nefes_render_cfg_t* cfg, const nefes_render_in_t* in, int64_t N, void* keep, void* stream) {
  using namespace nefes;
  const char* who = "nefes_render_rays_prepack";
  RenderPlan P;
  if (int e = make_plan(who, cfg, N, &P)) return e;
  if (N == 0 || cfg->prec != NEFES_PREC_BF16) return NEFES_OK;     // the fp32 / tf32 paths read the parameters directly
  NEFES_REQUIRE(in && keep && in->params_coarse && (!P.fine || in->params_fine), NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(((uintptr_t)keep & 255) == 0, NEFES_EALIGN, "%s: workspaces must be 256-byte aligned", who);
  char* K = (char*)keep;
  if (int e = mlp_prepack_bf16(in->params_coarse, cfg->net_coarse, P.mode_c, N, P.Sc, K + P.saved_c, (cudaStream_t)stream)) return e;
  if (P.fine)
    if (int e = mlp_prepack_bf16(in->params_fine, cfg->net_fine, P.mode_f, N, P.Sf, K + P.saved_f, (cudaStream_t)stream)) return e;
  return NEFES_OK;
}

int nefes_render_rays_fwd(const nefes_render_cfg_t* cfg, const nefes_render_in_t* in, int64_t N,
                          const nefes_render_out_t* out, void* keep, void* scratch, void* stream) {
  using namespace nefes;
  const char* who = "nefes_render_rays_fwd";
  RenderPlan P;
  if (int e = make_plan(who, cfg, N, &P)) return e;
  if (N == 0) return NEFES_OK;                       // an empty ray batch: nothing to launch, nothing to check
  NEFES_REQUIRE(in && out && keep && scratch, NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(in->rays && in->ld_rays >= 11 && in->params_coarse && in->t_vals, NEFES_EINVAL,
                "%s: rays [N, ld >= 11], coarse parameters and t_vals are required", who);
  NEFES_REQUIRE(!P.fine || (in->params_fine && in->u), NEFES_EINVAL, "%s: the fine pass needs params_fine and u", who);
  NEFES_REQUIRE(out->z_coarse && out->coarse.acc && out->coarse.weights, NEFES_EINVAL, "%s: coarse outputs missing", who);
  NEFES_REQUIRE(!P.fine || (out->z_fine && out->fine.acc && out->fine.weights), NEFES_EINVAL, "%s: fine outputs missing", who);
  NEFES_REQUIRE(((uintptr_t)keep & 255) == 0 && ((uintptr_t)scratch & 255) == 0, NEFES_EALIGN,
                "%s: workspaces must be 256-byte aligned", who);
  cudaStream_t st = (cudaStream_t)stream;
  char* K = (char*)keep;
  float* pts_c = (float*)(K + P.pts_c);
  float* dirs = (float*)(K + P.dirs);
  float* raw_c = (float*)(K + P.raw_c);
  const unsigned grid = (unsigned)ceil_div(N, 8);

  // ---- coarse pass (rendering.py:88-127)
  if (int e = nefes_sample_coarse(in->rays + 6, in->rays + 7, in->ld_rays, in->t_vals, in->t_rand, (int)N, P.Sc,
                                  out->z_coarse, stream)) return e;
  ray_points_kernel<<<grid, 256, 0, st>>>(in->rays, in->ld_rays, out->z_coarse, (int)N, P.Sc, pts_c, dirs, nullptr, 0, nullptr);
  NEFES_CHECK_LAUNCH("ray_points");
  {
    // nothing flows back into the sigma-only coarse pass of a test-time render (the importance samples are detached)
    ForwardOnlyScope fo(cfg->forward_only != 0 || P.mode_c == NEFES_MODE_SIGMA);
    WeightsPackedScope wp(cfg->weights_packed != 0);
    if (int e = (P.tiled_c ? nefes_mlp_fwd_tiles : nefes_mlp_fwd)(in->params_coarse, cfg->net_coarse, P.mode_c, cfg->prec, pts_c,
                                                                  dirs, N, P.Sc, raw_c, K + P.saved_c, scratch, stream)) return e;
  }
  if (int e = (P.tiled_c ? nefes_composite_fwd_tiles : nefes_composite_fwd)(raw_c, out->z_coarse, in->noise_coarse, (int)N, P.Sc,
                                                                            P.comp_c, cfg->beta_min, &out->coarse, stream)) return e;
  if (!P.fine) return NEFES_OK;

  // ---- hierarchical sampling + fine pass (rendering.py:129-154)
  float* pts_f = (float*)(K + P.pts_f);
  float* raw_f = (float*)(K + P.raw_f);
  if (int e = nefes_sample_fine(out->z_coarse, out->coarse.weights, in->u, in->u_per_ray, (int)N, P.Sc, cfg->n_importance,
                                out->z_fine, out->z_samples, out->inds, stream)) return e;
  NEFES_REQUIRE(out->z_std == nullptr || out->z_samples != nullptr, NEFES_EINVAL, "%s: z_std needs z_samples", who);
  ray_points_kernel<<<grid, 256, 0, st>>>(in->rays, in->ld_rays, out->z_fine, (int)N, P.Sf, pts_f, nullptr, out->z_samples,
                                          cfg->n_importance, out->z_std);
  NEFES_CHECK_LAUNCH("ray_points");
  const int net_f = cfg->net_fine;
  {
    ForwardOnlyScope fo(cfg->forward_only != 0);
    WeightsPackedScope wp(cfg->weights_packed != 0);
    if (int e = (P.tiled_f ? nefes_mlp_fwd_tiles : nefes_mlp_fwd)(in->params_fine, net_f, P.mode_f, cfg->prec, pts_f, dirs, N, P.Sf,
                                                                  raw_f, K + P.saved_f, scratch, stream)) return e;
  }
  return (P.tiled_f ? nefes_composite_fwd_tiles : nefes_composite_fwd)(raw_f, out->z_fine, in->noise_fine, (int)N, P.Sf, P.comp_f,
                                                                       cfg->beta_min, &out->fine, stream);
}

int nefes_render_rays_bwd(const nefes_render_cfg_t* cfg, const nefes_render_in_t* in, int64_t N,
                          const nefes_render_out_t* out, const nefes_comp_grad_t* g_coarse, const nefes_comp_grad_t* g_fine,
                          const void* keep, void* scratch, float* d_params_coarse, float* d_params_fine, float* d_rays,
                          void* stream) {
  using namespace nefes;
  const char* who = "nefes_render_rays_bwd";
  RenderPlan P;
  if (int e = make_plan(who, cfg, N, &P)) return e;
  if (N == 0) return NEFES_OK;
  NEFES_REQUIRE(!cfg->forward_only, NEFES_EINVAL, "%s: the forward call ran with forward_only = 1 and kept no activations", who);
  NEFES_REQUIRE(P.mode_c != NEFES_MODE_SIGMA || !any_grad(g_coarse), NEFES_EUNSUPPORTED,
                "%s: the sigma-only coarse pass of a test-time render keeps no activations (no gradient path: rendering.py:136)", who);
  NEFES_REQUIRE(in && out && keep && scratch, NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(in->rays && in->params_coarse && out->z_coarse && (!P.fine || (in->params_fine && out->z_fine)), NEFES_EINVAL,
                "%s: the forward call's inputs and depths are required", who);
  NEFES_REQUIRE(((uintptr_t)keep & 255) == 0 && ((uintptr_t)scratch & 255) == 0, NEFES_EALIGN,
                "%s: workspaces must be 256-byte aligned", who);
  cudaStream_t st = (cudaStream_t)stream;
  const char* K = (const char*)keep;
  char* W = (char*)scratch;
  const float* dirs = (const float*)(K + P.dirs);
  const bool want_rays = d_rays != nullptr;
  const float *dp_c = nullptr, *dp_f = nullptr, *dd_c = nullptr, *dd_f = nullptr;

  // one pass: compositing backward (compact cotangent on the tile path) -> field backward
  auto pass = [&](bool tiled, int net, int mode, int comp, int S, const float* params, int64_t o_pts, int64_t o_raw,
                  int64_t o_saved, const float* z, const float* noise, const nefes_comp_grad_t* g, int64_t o_cg,
                  float* d_params, int64_t o_dpts, int64_t o_ddirs, const float** dp, const float** dd) -> int {
    const float* pts = (const float*)(K + o_pts);
    const float* raw = (const float*)(K + o_raw);
    float* d_pts = want_rays ? (float*)(W + o_dpts) : nullptr;
    float* d_dirs = (want_rays && mode != NEFES_MODE_SIGMA) ? (float*)(W + o_ddirs) : nullptr;
    if (tiled) {
      float* cg = (float*)(W + o_cg);
      if (int e = nefes_composite_bwd_compact(raw, z, noise, (int)N, S, comp, g, cg, stream)) return e;
      if (int e = nefes_mlp_bwd_compact(params, net, mode, cfg->prec, pts, dirs, N, S, raw, cg, g->rgb, g->feat, K + o_saved,
                                        W + P.mlp_scratch, d_params, d_pts, d_dirs, stream)) return e;
    } else {
      float* d_raw = (float*)(W + P.draw);
      if (int e = nefes_composite_bwd(raw, z, noise, (int)N, S, comp, g, d_raw, stream)) return e;
      if (int e = nefes_mlp_bwd(params, net, mode, cfg->prec, pts, mode == NEFES_MODE_SIGMA ? nullptr : dirs, N, S, raw, d_raw,
                                K + o_saved, W + P.mlp_scratch, d_params, d_pts, d_dirs, stream)) return e;
    }
    *dp = d_pts;
    *dd = d_dirs;
    return NEFES_OK;
  };

  if (P.fine && any_grad(g_fine) && (d_params_fine || want_rays)) {
    const int net_f = cfg->net_fine;
    if (int e = pass(P.tiled_f, net_f, P.mode_f, P.comp_f, P.Sf, in->params_fine, P.pts_f, P.raw_f, P.saved_f, out->z_fine,
                     in->noise_fine, g_fine, P.cg_f, d_params_fine, P.dpts_f, P.ddirs_f, &dp_f, &dd_f)) return e;
  }
  // the coarse net reaches the results only through its own composited outputs: the importance samples are detached
  // (rendering.py:136), so with test_time's sigma-only coarse pass nothing flows into it at all
  if (any_grad(g_coarse) && (d_params_coarse || want_rays)) {
    if (int e = pass(P.tiled_c, cfg->net_coarse, P.mode_c, P.comp_c, P.Sc, in->params_coarse, P.pts_c, P.raw_c, P.saved_c,
                     out->z_coarse, in->noise_coarse, g_coarse, P.cg_c, d_params_coarse, P.dpts_c, P.ddirs_c, &dp_c, &dd_c))
      return e;
  }
  if (want_rays) {
    ray_points_bwd_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(dp_c, out->z_coarse, P.Sc, dp_f, out->z_fine, P.Sf, dd_c,
                                                                    dd_f, (int)N, in->ld_rays, d_rays);
    NEFES_CHECK_LAUNCH("ray_points_bwd");
  }
  return NEFES_OK;
}

}  // extern "C"
