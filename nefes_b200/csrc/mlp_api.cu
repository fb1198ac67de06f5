// extern "C" dispatch for the field MLP (K5) and the fused Adam step.
#include "common.cuh"

namespace nefes {
int mlp_fwd_fp32(const float*, int, int, const float*, const float*, int64_t, int, float*, void*, void*, cudaStream_t);
int mlp_bwd_fp32(const float*, int, int, const float*, const float*, int64_t, int, const float*, const float*,
                 const void*, void*, float*, float*, float*, cudaStream_t);
int mlp_workspace_fp32(int, int64_t, int64_t, int64_t*, int64_t*, int64_t*);
int gemm_mode_set(int tf32);          // mlp_fp32.cu: which GEMM the fp32-structured path runs on; returns the previous mode
struct GemmModeScope {                // NEFES_PREC_TF32 = the fp32 path with its GEMMs on tcgen05 kind::tf32
  int prev; bool on;
  explicit GemmModeScope(int prec) : prev(0), on(prec == NEFES_PREC_TF32) { if (on) prev = gemm_mode_set(1); }
  ~GemmModeScope() { if (on) gemm_mode_set(prev); }
};
int mlp_fwd_bf16(const float*, int, int, const float*, const float*, int64_t, int, float*, void*, void*, int, cudaStream_t);
int mlp_bwd_bf16(const float*, int, int, const float*, const float*, int64_t, int, const float*, const float*,
                 const void*, void*, float*, float*, float*, int, cudaStream_t, const float*, const float*, const float*);
int mlp_workspace_bf16(int, int, int64_t, int64_t, int64_t*, int64_t*, int64_t*);

// torch.optim.Adam semantics (amsgrad=False, weight_decay=0, maximize=False)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt, float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// The same step with the state that changes from step to step read from DEVICE memory, so that a captured CUDA graph of
// a whole training step replays correctly: state[0] = step count (incremented by adam_tick_kernel right before this
// launch), state[1] = learning rate.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ state, float b1, float b2,
                                float eps, float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = state[0], lr = state[1];
  const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}
__global__ void adam_tick_kernel(float* state) { state[0] += 1.f; }

static int check_mlp(const char* who, int net, int mode, int prec, int64_t N, int S) {
  NEFES_REQUIRE(net == NEFES_NET_COARSE || net == NEFES_NET_FINE, NEFES_EINVAL, "%s: bad net %d", who, net);
  NEFES_REQUIRE(mode >= NEFES_MODE_SIGMA && mode <= NEFES_MODE_FULL, NEFES_EINVAL, "%s: bad mode %d", who, mode);
  NEFES_REQUIRE(mode != NEFES_MODE_FULL || net == NEFES_NET_FINE, NEFES_EINVAL,
                "%s: MODE_FULL needs the fine net (the coarse net has no transient heads)", who);
  NEFES_REQUIRE(prec == NEFES_PREC_FP32 || prec == NEFES_PREC_BF16 || prec == NEFES_PREC_TF32, NEFES_EINVAL, "%s: bad precision %d", who, prec);
  NEFES_REQUIRE(N >= 0 && S >= 1 && N * (int64_t)S < (int64_t)1 << 31, NEFES_EINVAL,
                "%s: bad shape N=%lld S=%d", who, (long long)N, S);
  return NEFES_OK;
}
}  // namespace nefes

extern "C" {

int nefes_mlp_workspace(int net, int mode, int prec, int64_t M, int64_t N, int64_t* saved_bytes_host,
                        int64_t* scratch_fwd_bytes_host, int64_t* scratch_bwd_bytes_host) {
  NEFES_REQUIRE(saved_bytes_host && scratch_fwd_bytes_host && scratch_bwd_bytes_host, NEFES_EINVAL,
                "nefes_mlp_workspace: null output");
  if (int e = nefes::check_mlp("nefes_mlp_workspace", net, mode, prec, N > 0 ? N : 1, 1)) return e;
  if (prec == NEFES_PREC_BF16)
    return nefes::mlp_workspace_bf16(net, mode, M, N, saved_bytes_host, scratch_fwd_bytes_host, scratch_bwd_bytes_host);
  return nefes::mlp_workspace_fp32(mode, M, N, saved_bytes_host, scratch_fwd_bytes_host, scratch_bwd_bytes_host);
}

static int mlp_fwd_any(const char* who, int layout, const float* params, int net, int mode, int prec, const float* pts,
                       const float* dirs, int64_t N, int S, float* raw, void* saved, void* scratch, void* stream) {
  if (int e = nefes::check_mlp(who, net, mode, prec, N, S)) return e;
  NEFES_REQUIRE(params && pts && raw && saved, NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(mode == NEFES_MODE_SIGMA || (dirs && scratch), NEFES_EINVAL, "%s: dirs/scratch required", who);
  NEFES_REQUIRE(((uintptr_t)saved & 15) == 0 && ((uintptr_t)scratch & 15) == 0 && ((uintptr_t)raw & 15) == 0, NEFES_EALIGN,
                "%s: raw and workspaces must be 16-byte aligned", who);
  NEFES_REQUIRE(layout == NEFES_RAW_ROWS || prec == NEFES_PREC_BF16, NEFES_EINVAL,
                "%s: the tile-major raw layout belongs to the bf16 tensor path", who);
  if (N == 0) return NEFES_OK;
  if (prec == NEFES_PREC_BF16)
    return nefes::mlp_fwd_bf16(params, net, mode, pts, dirs, N, S, raw, saved, scratch, layout, (cudaStream_t)stream);
  nefes::GemmModeScope gemm(prec);
  return nefes::mlp_fwd_fp32(params, net, mode, pts, dirs, N, S, raw, saved, scratch, (cudaStream_t)stream);
}

static int mlp_bwd_any(const char* who, int layout, const float* params, int net, int mode, int prec, const float* pts,
                       const float* dirs, int64_t N, int S, const float* raw, const float* d_raw, const void* saved,
                       void* scratch, float* d_params, float* d_pts, float* d_dirs, void* stream,
                       const float* compact = nullptr, const float* g_rgb = nullptr, const float* g_feat = nullptr) {
  if (int e = nefes::check_mlp(who, net, mode, prec, N, S)) return e;
  NEFES_REQUIRE(params && pts && raw && (d_raw || compact) && saved && scratch, NEFES_EINVAL, "%s: null pointer", who);
  NEFES_REQUIRE(mode == NEFES_MODE_SIGMA || dirs, NEFES_EINVAL, "%s: dirs required", who);
  NEFES_REQUIRE(((uintptr_t)saved & 15) == 0 && ((uintptr_t)scratch & 15) == 0, NEFES_EALIGN,
                "%s: workspaces must be 16-byte aligned", who);
  NEFES_REQUIRE(layout == NEFES_RAW_ROWS || prec == NEFES_PREC_BF16, NEFES_EINVAL,
                "%s: the tile-major raw layout belongs to the bf16 tensor path", who);
  if (N == 0) return NEFES_OK;
  if (prec == NEFES_PREC_BF16)
    return nefes::mlp_bwd_bf16(params, net, mode, pts, dirs, N, S, raw, d_raw, saved, scratch, d_params, d_pts,
                               d_dirs, layout, (cudaStream_t)stream, compact, g_rgb, g_feat);
  nefes::GemmModeScope gemm(prec);
  return nefes::mlp_bwd_fp32(params, net, mode, pts, dirs, N, S, raw, d_raw, saved, scratch, d_params, d_pts,
                             d_dirs, (cudaStream_t)stream);
}

int nefes_mlp_fwd(const float* params, int net, int mode, int prec, const float* pts, const float* dirs,
                  int64_t N, int S, float* raw, void* saved, void* scratch, void* stream) {
  return mlp_fwd_any("nefes_mlp_fwd", NEFES_RAW_ROWS, params, net, mode, prec, pts, dirs, N, S, raw, saved, scratch, stream);
}
int nefes_mlp_fwd_tiles(const float* params, int net, int mode, int prec, const float* pts, const float* dirs,
                        int64_t N, int S, float* raw_tiles, void* saved, void* scratch, void* stream) {
  return mlp_fwd_any("nefes_mlp_fwd_tiles", NEFES_RAW_TILES, params, net, mode, prec, pts, dirs, N, S, raw_tiles, saved,
                     scratch, stream);
}
int nefes_mlp_bwd(const float* params, int net, int mode, int prec, const float* pts, const float* dirs,
                  int64_t N, int S, const float* raw, const float* d_raw, const void* saved, void* scratch,
                  float* d_params, float* d_pts, float* d_dirs, void* stream) {
  return mlp_bwd_any("nefes_mlp_bwd", NEFES_RAW_ROWS, params, net, mode, prec, pts, dirs, N, S, raw, d_raw, saved, scratch,
                     d_params, d_pts, d_dirs, stream);
}
int nefes_mlp_bwd_tiles(const float* params, int net, int mode, int prec, const float* pts, const float* dirs,
                        int64_t N, int S, const float* raw_tiles, const float* d_raw_tiles, const void* saved,
                        void* scratch, float* d_params, float* d_pts, float* d_dirs, void* stream) {
  return mlp_bwd_any("nefes_mlp_bwd_tiles", NEFES_RAW_TILES, params, net, mode, prec, pts, dirs, N, S, raw_tiles,
                     d_raw_tiles, saved, scratch, d_params, d_pts, d_dirs, stream);
}

// the two halves of nefes_mlp_bwd under the names of SURVEY 8b's export list
int nefes_mlp_dgrad(const float* params, int net, int mode, int prec, const float* pts, const float* dirs, int64_t N, int S,
                    const float* raw, const float* d_raw, const void* saved, void* scratch, float* d_pts, float* d_dirs,
                    void* stream) {
  NEFES_REQUIRE(d_pts != nullptr || d_dirs != nullptr, NEFES_EINVAL, "nefes_mlp_dgrad: no output requested");
  return mlp_bwd_any("nefes_mlp_dgrad", NEFES_RAW_ROWS, params, net, mode, prec, pts, dirs, N, S, raw, d_raw, saved, scratch,
                     nullptr, d_pts, d_dirs, stream);
}
int nefes_mlp_wgrad(const float* params, int net, int mode, int prec, const float* pts, const float* dirs, int64_t N, int S,
                    const float* raw, const float* d_raw, const void* saved, void* scratch, float* d_params, void* stream) {
  NEFES_REQUIRE(d_params != nullptr, NEFES_EINVAL, "nefes_mlp_wgrad: null d_params");
  return mlp_bwd_any("nefes_mlp_wgrad", NEFES_RAW_ROWS, params, net, mode, prec, pts, dirs, N, S, raw, d_raw, saved, scratch,
                     d_params, nullptr, nullptr, stream);
}

int nefes_mlp_bwd_compact(const float* params, int net, int mode, int prec, const float* pts, const float* dirs,
                          int64_t N, int S, const float* raw_tiles, const float* compact, const float* g_rgb,
                          const float* g_feat, const void* saved, void* scratch, float* d_params, float* d_pts,
                          float* d_dirs, void* stream) {
  NEFES_REQUIRE(compact != nullptr && mode != NEFES_MODE_SIGMA && prec == NEFES_PREC_BF16, NEFES_EINVAL,
                "nefes_mlp_bwd_compact: needs the compact cotangent, a colour mode and the bf16 path");
  return mlp_bwd_any("nefes_mlp_bwd_compact", NEFES_RAW_TILES, params, net, mode, prec, pts, dirs, N, S, raw_tiles, nullptr,
                     saved, scratch, d_params, d_pts, d_dirs, stream, compact, g_rgb, g_feat);
}

int nefes_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  NEFES_REQUIRE(params && grads && exp_avg && exp_avg_sq, NEFES_EINVAL, "nefes_adam_step: null pointer");
  NEFES_REQUIRE(n >= 0 && step >= 1, NEFES_EINVAL, "nefes_adam_step: bad n/step");
  if (n == 0) return NEFES_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  nefes::adam_kernel<<<(unsigned)nefes::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
  NEFES_CHECK_LAUNCH("adam");
  return NEFES_OK;
}

int nefes_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        float* state2, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  NEFES_REQUIRE(params && grads && exp_avg && exp_avg_sq && state2, NEFES_EINVAL, "nefes_adam_step_dev: null pointer");
  NEFES_REQUIRE(n >= 0, NEFES_EINVAL, "nefes_adam_step_dev: bad n");
  if (n == 0) return NEFES_OK;
  nefes::adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state2);
  nefes::adam_dev_kernel<<<(unsigned)nefes::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, state2, beta1, beta2, eps, grad_scale);
  NEFES_CHECK_LAUNCH("adam_dev");
  return NEFES_OK;
}

}  // extern "C"
