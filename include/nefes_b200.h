/*
 * nefes_b200 -- C ABI of the B200-native NeFeS render engine (libnefes_b200.so).
 *
 * The reference (ActiveVisionLab/NeFeS) has no FFI: its render path is Python callables
 * wired through a `render_kwargs` dict.  Each entry point below names the
 * reference callable (file:line under /root/reference/) whose arithmetic it replaces.  The Python
 * host side (the nefes_b200 Python package) keeps the reference's call surface and reaches these symbols
 * through ctypes; no torch types cross this boundary -- plain device pointers, sizes and a
 * cudaStream_t (passed as void*).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - all tensors are dense row-major fp32 unless stated; `ld*` arguments are row strides in
 *     elements;
 *   - all functions are asynchronous on `stream` and return 0 on success, a NEFES_E* code
 *     otherwise (nefes_last_error() gives the text).  There is no CPU fallback.
 */
#ifndef NEFES_B200_H_
#define NEFES_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEFES_VERSION 100          /* 0.1.0 */

enum {
  NEFES_OK = 0,
  NEFES_EINVAL = 1,                /* bad shape / mode / null pointer                       */
  NEFES_EALIGN = 2,                /* pointer not aligned as required                       */
  NEFES_ECUDA = 3,                 /* a CUDA runtime call or launch failed                  */
  NEFES_EUNSUPPORTED = 4           /* valid request this build does not implement           */
};

int nefes_version(void);
const char* nefes_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t nefes_launch_count(void);
/* Optional per-kernel timing for roofline reports: while enabled, the hot launches are bracketed by CUDA events on their
 * own stream and tagged with their algorithmic bytes / flops; the report is a JSON object
 * {tag: {launches, ms, alg_bytes, alg_flops}} (synchronises the device).  Off by default (no overhead). */
int nefes_prof_enable(int on);
int nefes_prof_report(char* buf, int cap);

/* ------------------------------------------------------------------------------------------
 * Field (MLP) description.  Architecture is the one every reference config uses:
 * D=8, W=128, skip at layer 4, xyz PE 63, dir PE 27, head 3+128 (script/models/nerfh_nff.py:
 * 421-505, options.py:30-31,99-100).  `net` selects which layers exist.
 * ------------------------------------------------------------------------------------------ */
enum { NEFES_NET_COARSE = 0, NEFES_NET_FINE = 1 };
/* forward modes == the three cases of run_network_NeRFH_NFF (nerfh_nff.py:192-231) */
enum {
  NEFES_MODE_SIGMA = 0,            /* NeRFH_NFF.forward(sigma_only=True)       -> raw [M,1]   */
  NEFES_MODE_STATIC = 1,           /* forward(output_transient=False)          -> raw [M,132] */
  NEFES_MODE_FULL = 2              /* forward(output_transient=True), fine net -> raw [M,137] */
};
/* arithmetic of the MLP */
enum {
  NEFES_PREC_FP32 = 0,             /* SIMT fp32 FMA -- the exact parity path                 */
  NEFES_PREC_BF16 = 1,             /* tcgen05 bf16 operands, fp32 TMEM accumulators, fused layer chains */
  NEFES_PREC_TF32 = 2              /* the fp32 path's layer-at-a-time structure (fp32 activations in HBM) with every GEMM on
                                      tcgen05 kind::tf32 (operands rounded to tf32, fp32 accumulate): the <= 1e-3 tensor path */
};
/* memory layout of the per-point network output `raw` (M points x C channels, fp32) */
enum {
  NEFES_RAW_ROWS = 0,              /* [M][C]: the reference's raw[N,S,C]                     */
  NEFES_RAW_TILES = 1              /* [ceil(M/128)][C][128]: engine-internal, see nefes_mlp_fwd_tiles */
};

#define NEFES_MAX_LAYERS 18
typedef struct {
  int n_layers;                    /* 12 (coarse) or 18 (fine)                              */
  int64_t n_params;                /* total floats in the flat buffer                       */
  int out_dim[NEFES_MAX_LAYERS];
  int in_dim[NEFES_MAX_LAYERS];
  int64_t w_off[NEFES_MAX_LAYERS]; /* offset of weight [out,in] (torch Linear layout)       */
  int64_t b_off[NEFES_MAX_LAYERS]; /* offset of bias [out]                                  */
  const char* name[NEFES_MAX_LAYERS]; /* state_dict prefix, e.g. "xyz_encoding_1.0"         */
} nefes_layout_t;
/* Flat fp32 parameter buffer layout (host struct filled in). Replaces the per-module
 * nn.Linear storage of nerfh_nff.py:469-505; names are the reference's state_dict keys. */
int nefes_param_layout(int net, nefes_layout_t* out_host);

/* ------------------------------------------------------------------------------------------
 * K1  get_rays / get_rays_batch            script/models/ray_utils.py:5-16, 46-59
 * c2w [B,3,4] -> rays_o, rays_d [B,H,W,3].  bwd: d_c2w [B,3,4] (zero-filled then accumulated).
 * ------------------------------------------------------------------------------------------ */
int nefes_get_rays_fwd(const float* c2w, int B, int H, int W, float focal,
                       float* rays_o, float* rays_d, void* stream);
int nefes_get_rays_bwd(const float* d_rays_o, const float* d_rays_d, int B, int H, int W,
                       float focal, float* d_c2w, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  stratified coarse depths              script/models/rendering.py:96-112
 * near/far: one value per ray read at near[i*ld_nf], far[i*ld_nf] (ray_batch columns 6,7).
 * t_vals [S] = torch.linspace(0,1,S) supplied by the host (never recomputed: SURVEY 7.3).
 * t_rand [N,S] or NULL (perturb == 0).
 * ------------------------------------------------------------------------------------------ */
int nefes_sample_coarse(const float* near, const float* far, int ld_nf, const float* t_vals,
                        const float* t_rand, int N, int S, float* z_vals, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  sample_pdf                            script/models/rendering.py:23-66
 * bins [N,nb], weights [N,nb-1]; u [N,ns] (per-ray) or [ns] when u_per_ray == 0 (det=True).
 * Outputs: samples [N,ns]; inds int32 [N,ns] = searchsorted(cdf,u,right=True) (may be NULL);
 * cdf_out [N,nb] (may be NULL).  nb <= 256, ns <= 256.
 * nefes_sample_pdf_from_cdf skips the pdf->cdf step (stage-level index parity).
 * ------------------------------------------------------------------------------------------ */
int nefes_sample_pdf(const float* bins, const float* weights, const float* u, int u_per_ray,
                     int N, int nb, int ns, float* samples, int32_t* inds, float* cdf_out,
                     void* stream);
int nefes_sample_pdf_from_cdf(const float* bins, const float* cdf, const float* u, int u_per_ray,
                              int N, int nb, int ns, float* samples, int32_t* inds, void* stream);
/* Fused hierarchical step of render_rays (rendering.py:132-141): mids of z_coarse [N,S],
 * weights_coarse[:,1:-1], sample_pdf, detach, sort(cat(z_coarse, z_samples)) -> z_fine [N,S+ns].
 * z_samples [N,ns] and inds [N,ns] may be NULL. */
int nefes_sample_fine(const float* z_coarse, const float* weights_coarse, const float* u,
                      int u_per_ray, int N, int S, int ns, float* z_fine, float* z_samples,
                      int32_t* inds, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4a frequency positional encoding         script/models/nerfh_nff.py:241-270
 * x [M,3] -> out [M, 3+6*n_freqs] written with row stride ld_out.
 * bwd: d_x [M,3] = d_out chained through sin/cos (overwrites d_x).
 * ------------------------------------------------------------------------------------------ */
int nefes_encode_pe_fwd(const float* x, int64_t M, int n_freqs, float* out, int ld_out, void* stream);
int nefes_encode_pe_bwd(const float* x, const float* d_out, int ld_out, int64_t M, int n_freqs,
                        float* d_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4b encoder front-end B: multiresolution HashGrid + spherical harmonics (tiny-cuda-nn's
 * algorithm, reached by the reference through tcnn.Encoding: script/models/nerfh_tcnn.py:65-75,
 * 97-103, 156, 209-210).  x [M,3] in [0,1]; table [n_entries, 2] fp32 (levels concatenated);
 * out [M, 2*n_levels], feature index = level*2 + f.  bwd: d_table is ACCUMULATED (fp32 atomics;
 * may be NULL), d_x [M,3] overwritten (may be NULL).  SH: d [M,3] in [0,1] (unit vector = 2d-1),
 * out [M,16] (degree 4).  Parity for this row is unpinned (oracle/hashgrid_oracle.py header).
 * ------------------------------------------------------------------------------------------ */
#define NEFES_HASH_MAX_LEVELS 32
typedef struct { float scale; uint32_t res; uint32_t size; uint32_t offset; uint32_t dense; } nefes_hash_level_t;
typedef struct { int n_levels; int64_t n_entries; nefes_hash_level_t level[NEFES_HASH_MAX_LEVELS]; } nefes_hash_layout_t;
int nefes_hash_layout(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale,
                      nefes_hash_layout_t* out_host);
int nefes_encode_hash_fwd(const float* x, const float* table, int64_t M, const nefes_hash_layout_t* layout_host,
                          float* out, void* stream);
int nefes_encode_hash_bwd(const float* x, const float* d_out, const float* table, int64_t M,
                          const nefes_hash_layout_t* layout_host, float* d_table, float* d_x, void* stream);
int nefes_encode_sh_fwd(const float* d, int64_t M, float* out, void* stream);
int nefes_encode_sh_bwd(const float* d, const float* d_out, int64_t M, float* d_d, void* stream);

/* Generic fp32 linear layer on the SIMT GEMM (torch Linear layout W [N,K]); used by front-end B's
 * small bias-free MLPs (tcnn FullyFusedMLP call sites nerfh_tcnn.py:79-89, 111-121).
 *   fwd:   C[M,N] = act(A[M,K] W^T + bias)      act: 0 none, 1 relu, 3 sigmoid; bias may be NULL
 *   dgrad: dA[M,K] = dC[M,N] W, multiplied by (mask[M,K] > 0) when mask != NULL (ReLU backward)
 *   wgrad: dW[N,K] += dC^T A                     (accumulated, fp32 atomics)                          */
/* Which GEMM nefes_linear_* (and the FusionNet convolutions of nefes_fusion_*) run on: 0 = SIMT fp32 (default), 1 = tcgen05
 * tf32.  Process-wide; returns the previous mode.  The NEFES_PREC_TF32 entry points switch it around their own calls. */
int nefes_gemm_mode(int tf32);
int nefes_linear_fwd(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
                     int64_t M, int N, int K, int act, void* stream);
int nefes_linear_dgrad(const float* dC, int64_t ldc, const float* W, float* dA, int64_t lda, int64_t M, int N, int K,
                       const float* mask, int64_t ldm, void* stream);
int nefes_linear_wgrad(const float* dC, int64_t ldc, const float* A, int64_t lda, float* dW, int64_t M, int N, int K,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  the NeFeS field: PE + MLP             script/models/nerfh_nff.py:168-231 (query) and
 *                                           :525-576 (NeRFH_NFF.forward)
 * pts [M,3] with M = N*S, dirs [N,3] (unit view directions, one per ray, broadcast over the S
 * samples of the ray exactly as nerfh_nff.py:206-207 expands them; ignored in MODE_SIGMA).
 * raw [M,C], C = 1 / 132 / 137.
 * `saved` keeps the activations backward needs (NULL => inference only); `scratch` is
 * temporary.  Sizes from nefes_mlp_workspace.
 * ------------------------------------------------------------------------------------------ */
int nefes_mlp_workspace(int net, int mode, int prec, int64_t M, int64_t N,
                        int64_t* saved_bytes_host, int64_t* scratch_fwd_bytes_host,
                        int64_t* scratch_bwd_bytes_host);
int nefes_mlp_fwd(const float* params, int net, int mode, int prec, const float* pts,
                  const float* dirs, int64_t N, int S, float* raw, void* saved, void* scratch,
                  void* stream);
/* Same, with raw in the engine's tile-major layout: blocks of 128 consecutive points, each [C][128] fp32
 * (raw_tiles[(m/128)*C*128 + c*128 + m%128]); the buffer holds ceil(M/128)*C*128 floats.  This is the layout the
 * fused chain writes with coalesced stores and nefes_composite_*_tiles reads; it never crosses the reference-facing
 * Python surface (run_network_NeRFH_NFF / raw2outputs_NeRFH_NFF speak [N,S,C]).  bf16 precision only. */
int nefes_mlp_fwd_tiles(const float* params, int net, int mode, int prec, const float* pts,
                        const float* dirs, int64_t N, int S, float* raw_tiles, void* saved, void* scratch,
                        void* stream);
int nefes_mlp_bwd_tiles(const float* params, int net, int mode, int prec, const float* pts,
                        const float* dirs, int64_t N, int S, const float* raw_tiles, const float* d_raw_tiles,
                        const void* saved, void* scratch, float* d_params, float* d_pts, float* d_dirs,
                        void* stream);
/* Backward from the COMPACT cotangent of raw that nefes_composite_bwd_compact writes ([N][5][S] fp32: static weight,
 * transient weight, d sigma, d transient sigma, d beta per sample) plus the per-ray cotangents of the composited rgb
 * [N,3] and feature [N,128] (either may be NULL = zero): d_raw[s, c] = weight[s] * g_ray[c] for the 134 colour / feature
 * channels, so the 137-channel fp32 block never exists in HBM.  Tile-major raw, bf16 path, colour modes. */
int nefes_mlp_bwd_compact(const float* params, int net, int mode, int prec, const float* pts, const float* dirs,
                          int64_t N, int S, const float* raw_tiles, const float* compact, const float* g_rgb,
                          const float* g_feat, const void* saved, void* scratch, float* d_params, float* d_pts,
                          float* d_dirs, void* stream);
/* d_params (flat, same layout as params) is ACCUMULATED into (caller zero-fills) or NULL
 * (frozen weights: refinement); d_pts [M,3] / d_dirs [N,3] are overwritten, or NULL. */
int nefes_mlp_bwd(const float* params, int net, int mode, int prec, const float* pts,
                  const float* dirs, int64_t N, int S, const float* raw, const float* d_raw,
                  const void* saved, void* scratch, float* d_params, float* d_pts, float* d_dirs,
                  void* stream);

/* The two halves of nefes_mlp_bwd on their own: data gradient only (frozen field: pose refinement; d_pts / d_dirs as
 * above, at least one non-NULL) and weight gradient only (d_params ACCUMULATED).  Same saved / scratch workspaces.
 * Weights are re-packed into the tensor path's bf16 operand images inside every forward call (the images live in
 * `saved`, which backward reads), so there is no separate repack call for a caller to forget after an optimiser step. */
int nefes_mlp_dgrad(const float* params, int net, int mode, int prec, const float* pts, const float* dirs, int64_t N, int S,
                    const float* raw, const float* d_raw, const void* saved, void* scratch, float* d_pts, float* d_dirs,
                    void* stream);
int nefes_mlp_wgrad(const float* params, int net, int mode, int prec, const float* pts, const float* dirs, int64_t N, int S,
                    const float* raw, const float* d_raw, const void* saved, void* scratch, float* d_params, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6  raw2outputs_NeRFH_NFF                 script/models/nerfh_nff.py:25-166
 * ------------------------------------------------------------------------------------------ */
enum {
  NEFES_COMP_SIGMA = 0,            /* coarse & test_time: raw [N,S,1] -> acc, weights only (:83-89) */
  NEFES_COMP_STATIC = 1,           /* output_transient=False: raw [N,S,132]          (:153-165) */
  NEFES_COMP_TRANSIENT = 2,        /* raw [N,S,137], static+transient composite      (:119-150) */
  NEFES_COMP_TRANSIENT_STATIC_ONLY = 3 /* test_time && !transient_at_test             (:92-117) */
};
typedef struct {
  float* rgb;                      /* [N,3]   */
  float* feat;                     /* [N,128] */
  float* disp;                     /* [N]     */
  float* acc;                      /* [N]     */
  float* weights;                  /* [N,S]   */
  float* depth;                    /* [N]     */
  float* beta;                     /* [N]     */
  float* tsig;                     /* [N,S] transient_sigmas = raw[...,135] (transient modes) or NULL */
} nefes_comp_out_t;
/* noise [N,S] = randn * raw_noise_std or NULL. */
int nefes_composite_fwd(const float* raw, const float* z_vals, const float* noise, int N, int S,
                        int mode, float beta_min, const nefes_comp_out_t* out_host, void* stream);
/* Cotangents (any may be NULL = zero): same shapes as the outputs, plus d_tsig [N,S] for the
 * transient_sigmas view.  d_raw [N,S,C] is overwritten.  No gradient flows to z_vals (they are
 * detached constants on this path: rendering.py:136). */
typedef struct {
  const float* rgb; const float* feat; const float* disp; const float* acc;
  const float* weights; const float* depth; const float* beta; const float* tsig;
} nefes_comp_grad_t;
int nefes_composite_bwd(const float* raw, const float* z_vals, const float* noise, int N, int S,
                        int mode, const nefes_comp_grad_t* g_host, float* d_raw, void* stream);
/* The same two operators on tile-major raw / d_raw blocks (NEFES_RAW_TILES, see nefes_mlp_fwd_tiles); S must divide
 * 128 so that a ray never straddles a tile (64 coarse / 128 fine samples on every reference config). */
int nefes_composite_fwd_tiles(const float* raw_tiles, const float* z_vals, const float* noise, int N, int S,
                              int mode, float beta_min, const nefes_comp_out_t* out_host, void* stream);
int nefes_composite_bwd_tiles(const float* raw_tiles, const float* z_vals, const float* noise, int N, int S,
                              int mode, const nefes_comp_grad_t* g_host, float* d_raw_tiles, void* stream);
/* ... writing the compact cotangent [N][5][S] consumed by nefes_mlp_bwd_compact instead of d_raw */
int nefes_composite_bwd_compact(const float* raw_tiles, const float* z_vals, const float* noise, int N, int S,
                                int mode, const nefes_comp_grad_t* g_host, float* compact, void* stream);

/* ------------------------------------------------------------------------------------------
 * The whole path as one call: render_rays    script/models/rendering.py:68-180
 * (stratified depths -> coarse field -> compositing -> sample_pdf + merge -> fine field -> compositing), forward and
 * backward.  Intermediates -- sample points, raw (tile-major on the bf16 path), saved activations, compact cotangents --
 * live in two caller-provided workspaces sized by nefes_render_rays_workspace: `keep` must survive from the forward to
 * the backward call, `scratch` is per call.  Both 256-byte aligned.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int n_samples;                   /* N_samples: coarse depths per ray                                   */
  int n_importance;                /* N_importance: > 0 runs the fine pass on n_samples + n_importance   */
  int prec;                        /* NEFES_PREC_*                                                        */
  int test_time;                   /* coarse pass sigma-only, no rgb0/feat0 (rendering.py:122-127)        */
  int output_transient;            /* args.NeRFW: fine pass in NEFES_MODE_FULL / transient compositing    */
  int transient_at_test;           /* args.transient_at_test                                              */
  int net_coarse, net_fine;        /* NEFES_NET_* of network_fn / network_fine                            */
  float beta_min;                  /* network_fine.beta_min                                               */
  int forward_only;                /* 1: no backward call will follow (inference): activations are not saved */
  int weights_packed;              /* 1: frozen parameters (pose refinement): the bf16 operand images in `keep` were written by
                                      nefes_render_rays_prepack (or an earlier forward call) with this cfg, N and keep, and the
                                      parameters have not changed since -- the forward skips its re-pack.  0: re-pack (default) */
} nefes_render_cfg_t;
typedef struct {
  const float* rays;               /* ray_batch [N, ld_rays]: o 0:3, d 3:6, near 6, far 7, viewdirs 8:11  */
  int ld_rays;                     /* >= 11                                                               */
  const float* params_coarse;      /* flat parameter buffers (nefes_param_layout)                         */
  const float* params_fine;        /* NULL when n_importance == 0                                         */
  const float* t_vals;             /* [n_samples] = torch.linspace(0,1,n_samples), host-made (see K2)     */
  const float* t_rand;             /* [N, n_samples] or NULL (perturb == 0)                               */
  const float* u;                  /* [N, n_importance], or [n_importance] with u_per_ray == 0 (det)      */
  int u_per_ray;
  const float* noise_coarse;       /* [N, n_samples] randn * raw_noise_std or NULL                        */
  const float* noise_fine;         /* [N, n_samples + n_importance] or NULL (static fine compositing only) */
} nefes_render_in_t;
typedef struct {
  nefes_comp_out_t coarse;         /* rgb0, feat0, disp0, acc0, weights [N,S], depth, beta (test_time: acc, weights only) */
  nefes_comp_out_t fine;           /* rgb_map, feat_map, disp_map, acc_map, weights, depth, beta, transient_sigmas        */
  float* z_coarse;                 /* [N, n_samples]                                                      */
  float* z_fine;                   /* [N, n_samples + n_importance] sorted union                          */
  float* z_samples;                /* [N, n_importance] (may be NULL)                                     */
  int32_t* inds;                   /* [N, n_importance] searchsorted indices (may be NULL)                */
  float* z_std;                    /* [N] population std of z_samples (may be NULL; needs z_samples)      */
} nefes_render_out_t;
int nefes_render_rays_workspace(const nefes_render_cfg_t* cfg_host, int64_t N, int64_t* keep_bytes_host,
                                int64_t* scratch_fwd_bytes_host, int64_t* scratch_bwd_bytes_host);
/* Packs the parameters of both fields into the bf16 operand images inside `keep` (what every forward call does first unless
 * cfg.weights_packed is set).  A caller whose parameters are frozen -- the refinement loop, dm/DFM_pose_refine.py:290-348, renders
 * the same two fields 50 times per query -- calls this once per query and runs its iterations with weights_packed = 1.
 * No-op for the fp32 / tf32 precisions. */
int nefes_render_rays_prepack(const nefes_render_cfg_t* cfg_host, const nefes_render_in_t* in_host, int64_t N, void* keep,
                              void* stream);
int nefes_render_rays_fwd(const nefes_render_cfg_t* cfg_host, const nefes_render_in_t* in_host, int64_t N,
                          const nefes_render_out_t* out_host, void* keep, void* scratch, void* stream);
/* g_coarse / g_fine: cotangents of the two composited output sets (NULL or all-NULL = none).  d_params_* are
 * ACCUMULATED into (caller zero-fills) or NULL (frozen field); d_rays [N, ld_rays] is overwritten (columns 0:6 and 8:11,
 * zeros elsewhere) or NULL.  `in` / `out` are the forward call's (same buffers, still holding its results). */
int nefes_render_rays_bwd(const nefes_render_cfg_t* cfg_host, const nefes_render_in_t* in_host, int64_t N,
                          const nefes_render_out_t* out_host, const nefes_comp_grad_t* g_coarse_host,
                          const nefes_comp_grad_t* g_fine_host, const void* keep, void* scratch, float* d_params_coarse,
                          float* d_params_fine, float* d_rays, void* stream);

/* ------------------------------------------------------------------------------------------
 * Refinement iteration glue: the pose chain in front of get_rays, the feature-metric loss behind the
 * render and the optimiser step, so that one iteration of DFM_pose_refine.py:380-440 is engine launches only.
 *   pose6 = [r(3), t(3)]: c2w = [Exp(r) @ R0 | t + t0] with init_c2w = [R0 | t0] [3,4]
 *                         (script/models/poses.py:25-50 with lietorch=False; utils/lie_group_helper.py:60-81)
 *   chain6 (HOST pointer to 6 floats, or NULL = {0, 1, 0, 0, 0, 1}) = {se3, pose_scale, move_x, move_y, move_z, pose_scale2}:
 *                         se3 != 0: the translation goes through the SE(3) exponential, c2w = [Exp(r) R0 | V(r) t + t0]
 *                         (poses.py:31-32, 44: SE3.exp([t, r]).matrix() -- the lietorch=True branch DFM_pose_refine.py:374 uses;
 *                         closed form evaluated in fp64); then the translation column x becomes ((x * pose_scale) + move) *
 *                         pose_scale2 (dm/direct_pose_model.py:210-232, fix_coord_supp) before the rays are generated
 *   nefes_pose_rays_fwd  writes c2w_out [3,4] (may be NULL) and the packed ray_batch [H*W, ld >= 11] render() builds
 *                        (rendering.py:197-243: o, d, near, far, d/|d|, zeros) with get_rays' arithmetic (ray_utils.py:5-16)
 *   nefes_pose_rays_bwd  cotangent of ray_batch -> d_c2w [3,4], ACCUMULATED (caller zero-fills once)
 *   nefes_cosine_loss_*  loss = 1 - mean_c cosine_similarity(feat[:, c], target[c, :]) (DFM_pose_refine.py:236-255,
 *                        per_pixel=False); mask [N] (may be NULL): only pixels with mask > 0 enter (masked_feature_loss,
 *                        DFM_pose_refine.py:257-288); feat [N,C] row-major, target [C,N]; stats [3,C] ACCUMULATED by _fwd (caller
 *                        zero-fills once); _bwd writes loss (1 float, may be NULL), loss_hist[(int)*step] when both are
 *                        given, and d_feat [N,C] (may be NULL)
 *   nefes_pose_adam_step d_c2w -> (d_r, d_t) through the exponential, torch.optim.Adam on the two groups (lr_r, lr_t);
 *                        state13 = exp_avg[6], exp_avg_sq[6], step; clears d_c2w and zero[0:n_zero] (the loss statistics)
 * ------------------------------------------------------------------------------------------ */
int nefes_pose_rays_fwd(const float* pose6, const float* init_c2w, int H, int W, float focal, float near, float far,
                        float* c2w_out, float* ray_batch, int ld, const float* chain6, void* stream);
int nefes_pose_rays_bwd(const float* d_ray_batch, const float* ray_batch, int ld, int H, int W, float focal, float* d_c2w,
                        void* stream);
int nefes_cosine_loss_fwd(const float* feat, const float* target, const float* mask, int N, int C, float* stats, void* stream);
int nefes_cosine_loss_bwd(const float* feat, const float* target, const float* mask, const float* stats, int N, int C, float* loss,
                          float* loss_hist, const float* step, int hist_cap, float* d_feat, void* stream);
/* bicubic up-sampling + border crop of a pixel-major map (dm/DFM_APR_refine.py:114-124: torch.nn.Upsample(size=(H, W),
 * mode='bicubic') followed by [:, :, crop:-crop, crop:-crop]): x [h*w, C] -> out [(H - 2 crop) * (W - 2 crop), C];
 * _bwd: d_out -> d_x [h*w, C] (overwritten), tmp = h * (W - 2 crop) * C floats of scratch */
int nefes_upsample_crop_fwd(const float* x, int h, int w, int C, int H, int W, int crop, float* out, void* stream);
int nefes_upsample_crop_bwd(const float* d_out, int h, int w, int C, int H, int W, int crop, float* tmp, float* d_x, void* stream);
int nefes_pose_adam_step(float* pose6, const float* init_c2w, float* d_c2w, float* zero, int n_zero, float* state13,
                         float lr_r, float lr_t, float beta1, float beta2, float eps, const float* chain6, void* stream);

/* ------------------------------------------------------------------------------------------
 * Caller-side helpers on the "next" rows of SURVEY 8f that the training step needs resident.
 * Fused Adam on the flat buffers (torch.optim.Adam, betas (0.9, 0.999), eps 1e-8, no weight
 * decay: nerfh_nff.py:682); grad is scaled by grad_scale first (1/world_size after all-reduce).
 * ------------------------------------------------------------------------------------------ */
int nefes_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                    int64_t n, float lr, float beta1, float beta2, float eps, int step,
                    float grad_scale, void* stream);
/* The same update with the per-step state on the device -- state2[0] = step count (float, incremented by the call),
 * state2[1] = learning rate -- so a CUDA graph that captured one training step replays with the right bias
 * correction and a schedule the host can change by writing state2[1]. */
int nefes_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        float* state2, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* NeRF-W colour loss of the stage-1 step (script/models/losses.py:96-132, NerfWLoss with the transient head):
 *   loss = coef * (0.5 mean((rgb_coarse-t)^2) + mean((rgb_fine-t)^2 / (2 beta^2)) + 3 + mean(log beta)
 *                  + lambda_u mean(transient_sigmas))
 * rgb_* / target [N,3], beta [N], transient_sigmas [N,S]; scratch8: 8 floats of device scratch; loss: 1 float.
 * bwd: d_loss is the upstream cotangent of the scalar (device pointer); the four gradients are overwritten. */
int nefes_nerfw_loss_fwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* transient_sigmas,
                         const float* target, int64_t N, int S, float coef, float lambda_u, float* scratch8, float* loss,
                         void* stream);
int nefes_nerfw_loss_bwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* target,
                         const float* d_loss, int64_t N, int S, float coef, float lambda_u, float* d_rgb_coarse,
                         float* d_rgb_fine, float* d_beta, float* d_transient_sigmas, void* stream);

/* Feature loss of the stage-2/3 training step -- script/models/losses.py:134-173 (ColorFeatureFusionNerfWLoss.f_loss:
 * nn.L1Loss / nn.MSELoss, reduction 'mean').  loss = mean_f(feat_a - target) [+ mean_f(feat_b - target) when feat_b != NULL],
 * f = |.| (mode 0) or (.)^2 (mode 1); n_elems = N * 128 (a multiple of 4; pointers 16-byte aligned).  scratch2: 2 floats. */
int nefes_feat_loss_fwd(const float* feat_a, const float* feat_b, const float* target, int64_t n_elems, int mode, float* scratch2,
                        float* loss, void* stream);
int nefes_feat_loss_bwd(const float* feat_a, const float* feat_b, const float* target, const float* d_loss, int64_t n_elems,
                        int mode, float* d_feat_a, float* d_feat_b, void* stream);

/* ------------------------------------------------------------------------------------------
 * Post-render appearance + fusion stage, pixel-major tensors ([P, C], P = B*H*W image-major).
 *   FusionNet  script/models/nerfh_nff.py:356-418, :578-603: rgb [P,3] (normalised inside with the ImageNet mean / std) and
 *              feat [P,128] -> conv 131->64 (3x3) ReLU -> 64->64 ReLU -> 64->64 ReLU -> 64->128 (5x5) -> BatchNorm2d(128)
 *              -> out [P,128].  Weights in torch layout [Cout, Cin, kh, kw]; training != 0: batch statistics and the running
 *              estimates are updated (momentum); training == 0: running statistics; no_bn: the net ends at conv4;
 *              residual (fusion_residule): out += feat.  workspace: nefes_fusion_workspace(P) bytes, shared by _fwd and _bwd
 *              (it keeps X0, the three hidden activations and the pre-BN output).  _bwd ACCUMULATES into the non-NULL
 *              members of `grads` and overwrites d_rgb [P,3] / d_feat [P,128] (either may be NULL).
 *   affine colour transform  nerfh_nff.py:511-522, :605-626: exposure_params = the exposure network's flat tiny-cuda-nn
 *              buffer (weights [32x16 | 32x32 | 32x32 | 16x32] row-major, no biases; its 10 inputs are padded to 16 with
 *              ONES); hist [B,10] (truncated to integers, as hist.long()); rgb [B * n_per_image, 3] image-major;
 *              out = sigmoid(K_b rgb + b_b).  ab12 [B,12] and hidden96 [B,96] are written by _fwd and read by _bwd, which
 *              overwrites d_ab12 [B,12] and d_rgb (may be NULL) and ACCUMULATES into d_exposure_params (may be NULL).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* weight[4];          /* net.0 / net.2 / net.4 / net.6 .weight: [64,131,3,3] [64,64,3,3] [64,64,3,3] [128,64,5,5] */
  const float* bias[4];            /* ... .bias                                                                             */
  const float* bn_weight;          /* net.7.weight [128] (NULL when no_bn)                                                   */
  const float* bn_bias;            /* net.7.bias                                                                             */
  const float* bn_running_mean;    /* net.7.running_mean (updated in place when training)                                    */
  const float* bn_running_var;     /* net.7.running_var                                                                      */
} nefes_fusion_params_t;
typedef struct {
  float* weight[4];                /* accumulated into; any member may be NULL                                               */
  float* bias[4];
  float* bn_weight;
  float* bn_bias;
} nefes_fusion_grads_t;
int64_t nefes_fusion_workspace(int64_t n_pixels);
int nefes_fusion_fwd(const nefes_fusion_params_t* params_host, const float* rgb, const float* feat, int B, int H, int W, int training,
                     int no_bn, int residual, float momentum, float eps, void* workspace, float* out, void* stream);
int nefes_fusion_bwd(const nefes_fusion_params_t* params_host, const nefes_fusion_grads_t* grads_host, const float* d_out, int B, int H,
                     int W, int training, int no_bn, int residual, void* workspace, float* d_rgb, float* d_feat, void* stream);
int nefes_affine_color_fwd(const float* exposure_params, const float* hist, const float* rgb, int B, int64_t n_per_image, float* ab12,
                           float* hidden96, float* out, void* stream);
int nefes_affine_color_bwd(const float* exposure_params, const float* hist, const float* rgb, const float* out, const float* d_out,
                           const float* ab12, const float* hidden96, int B, int64_t n_per_image, float* d_ab12, float* d_rgb,
                           float* d_exposure_params, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* NEFES_B200_H_ */
