/* A host with no Python and no torch: render_rays forward + backward through the C ABI of libnefes_b200.so.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/render_c_abi.c \
 *       -L nefes_b200/lib -lnefes_b200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/nefes_b200/lib -o render_c_abi
 *
 * Random-init style weights (small uniform numbers), 512 camera-like rays, 64 + 64 samples, bf16 tensor path, training
 * configuration (perturbed depths, NeRF-W transient heads).  Prints a few composited values and gradient norms; exits
 * non-zero on any error.  This is the call sequence nefes_b200/ops.py (_RenderRays) makes through ctypes. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime.h>

#include "nefes_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define NF(x) do { int e_ = (x); if (e_ != NEFES_OK) { fprintf(stderr, "%s: code %d: %s\n", #x, e_, nefes_last_error()); return 3; } } while (0)

static float urand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (float)(*s >> 8) / 16777216.0f; }

static float* dev_floats(size_t n, const float* host) {
  float* p = NULL;
  if (cudaMalloc((void**)&p, (n ? n : 1) * sizeof(float)) != cudaSuccess) return NULL;
  if (host) cudaMemcpy(p, host, n * sizeof(float), cudaMemcpyHostToDevice);
  else cudaMemset(p, 0, (n ? n : 1) * sizeof(float));
  return p;
}

int main(void) {
  const int N = 512, S = 64, NI = 64, SF = S + NI, LD = 21;
  unsigned seed = 1u;
  printf("libnefes_b200 version %d\n", nefes_version());

  /* parameters: the flat fp32 buffers of nefes_param_layout, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear */
  nefes_layout_t lay[2];
  float* params[2];
  float* d_params[2];
  for (int net = 0; net < 2; ++net) {
    NF(nefes_param_layout(net, &lay[net]));
    float* h = (float*)malloc(sizeof(float) * (size_t)lay[net].n_params);
    for (int l = 0; l < lay[net].n_layers; ++l) {
      const float b = 1.0f / sqrtf((float)lay[net].in_dim[l]);
      for (int64_t i = 0; i < (int64_t)lay[net].out_dim[l] * lay[net].in_dim[l]; ++i) h[lay[net].w_off[l] + i] = (2.f * urand(&seed) - 1.f) * b;
      for (int i = 0; i < lay[net].out_dim[l]; ++i) h[lay[net].b_off[l] + i] = (2.f * urand(&seed) - 1.f) * b;
    }
    params[net] = dev_floats((size_t)lay[net].n_params, h);
    d_params[net] = dev_floats((size_t)lay[net].n_params, NULL);
    free(h);
  }

  /* ray batch rows: o, d, near, far, d/|d|, zeros (rendering.py:197-243) + the random draws of one training step */
  float* h_rays = (float*)calloc((size_t)N * LD, sizeof(float));
  float* h_t = (float*)malloc(sizeof(float) * (size_t)N * S);
  float* h_u = (float*)malloc(sizeof(float) * (size_t)N * NI);
  float h_tv[64];
  for (int i = 0; i < S; ++i) h_tv[i] = (float)i / (float)(S - 1);         /* torch.linspace(0, 1, 64) */
  for (int i = 0; i < N; ++i) {
    float* r = h_rays + (size_t)i * LD;
    float d[3] = {(float)(i % 32 - 16) / 20.f, -(float)(i / 32 - 8) / 20.f, -1.f};
    const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int c = 0; c < 3; ++c) { r[c] = 0.1f * c; r[3 + c] = d[c]; r[8 + c] = d[c] / n; }
    r[6] = 0.f; r[7] = 4.f;
  }
  for (int i = 0; i < N * S; ++i) h_t[i] = urand(&seed);
  for (int i = 0; i < N * NI; ++i) h_u[i] = urand(&seed);

  nefes_render_cfg_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.n_samples = S; cfg.n_importance = NI; cfg.prec = NEFES_PREC_BF16; cfg.test_time = 0; cfg.output_transient = 1;
  cfg.transient_at_test = 1; cfg.net_coarse = NEFES_NET_COARSE; cfg.net_fine = NEFES_NET_FINE; cfg.beta_min = 0.1f;
  int64_t keep_b = 0, fwd_b = 0, bwd_b = 0;
  NF(nefes_render_rays_workspace(&cfg, N, &keep_b, &fwd_b, &bwd_b));
  void *keep = NULL, *scratch = NULL;
  CK(cudaMalloc(&keep, (size_t)keep_b));
  CK(cudaMalloc(&scratch, (size_t)(fwd_b > bwd_b ? fwd_b : bwd_b)));
  printf("workspaces: keep %.1f MB, scratch %.1f MB\n", keep_b / 1e6, (fwd_b > bwd_b ? fwd_b : bwd_b) / 1e6);

  nefes_render_in_t in;
  memset(&in, 0, sizeof in);
  in.rays = dev_floats((size_t)N * LD, h_rays); in.ld_rays = LD;
  in.params_coarse = params[0]; in.params_fine = params[1];
  in.t_vals = dev_floats(S, h_tv); in.t_rand = dev_floats((size_t)N * S, h_t);
  in.u = dev_floats((size_t)N * NI, h_u); in.u_per_ray = 1;

  nefes_render_out_t out;
  memset(&out, 0, sizeof out);
  out.coarse.rgb = dev_floats(N * 3, NULL); out.coarse.feat = dev_floats(N * 128, NULL); out.coarse.disp = dev_floats(N, NULL);
  out.coarse.acc = dev_floats(N, NULL); out.coarse.weights = dev_floats((size_t)N * S, NULL); out.coarse.depth = dev_floats(N, NULL);
  out.coarse.beta = dev_floats(N, NULL);
  out.fine.rgb = dev_floats(N * 3, NULL); out.fine.feat = dev_floats(N * 128, NULL); out.fine.disp = dev_floats(N, NULL);
  out.fine.acc = dev_floats(N, NULL); out.fine.weights = dev_floats((size_t)N * SF, NULL); out.fine.depth = dev_floats(N, NULL);
  out.fine.beta = dev_floats(N, NULL); out.fine.tsig = dev_floats((size_t)N * SF, NULL);
  out.z_coarse = dev_floats((size_t)N * S, NULL); out.z_fine = dev_floats((size_t)N * SF, NULL);
  out.z_samples = dev_floats((size_t)N * NI, NULL); out.z_std = dev_floats(N, NULL);
  CK(cudaMalloc((void**)&out.inds, sizeof(int32_t) * (size_t)N * NI));

  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  NF(nefes_render_rays_fwd(&cfg, &in, N, &out, keep, scratch, st));

  /* loss = sum(rgb) + sum(feat) + sum(rgb0): cotangents of ones */
  float* ones = (float*)malloc(sizeof(float) * (size_t)N * 128);
  for (int i = 0; i < N * 128; ++i) ones[i] = 1.f;
  float* d_ones = dev_floats((size_t)N * 128, ones);
  nefes_comp_grad_t g_coarse, g_fine;
  memset(&g_coarse, 0, sizeof g_coarse);
  memset(&g_fine, 0, sizeof g_fine);
  g_fine.rgb = d_ones; g_fine.feat = d_ones; g_coarse.rgb = d_ones;
  float* d_rays = dev_floats((size_t)N * LD, NULL);
  NF(nefes_render_rays_bwd(&cfg, &in, N, &out, &g_coarse, &g_fine, keep, scratch, d_params[0], d_params[1], d_rays, st));
  CK(cudaStreamSynchronize(st));

  float rgb[6], acc[2];
  CK(cudaMemcpy(rgb, out.fine.rgb, sizeof rgb, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(acc, out.fine.acc, sizeof acc, cudaMemcpyDeviceToHost));
  printf("ray 0: rgb %.5f %.5f %.5f acc %.5f | ray 1: rgb %.5f %.5f %.5f acc %.5f\n", rgb[0], rgb[1], rgb[2], acc[0], rgb[3], rgb[4], rgb[5], acc[1]);
  int bad = 0;
  for (int net = 0; net < 2; ++net) {
    float* h = (float*)malloc(sizeof(float) * (size_t)lay[net].n_params);
    CK(cudaMemcpy(h, d_params[net], sizeof(float) * (size_t)lay[net].n_params, cudaMemcpyDeviceToHost));
    double s2 = 0;
    for (int64_t i = 0; i < lay[net].n_params; ++i) { s2 += (double)h[i] * h[i]; if (!isfinite(h[i])) bad = 1; }
    printf("|d params %s| = %.6g\n", net ? "fine" : "coarse", sqrt(s2));
    if (!(s2 > 0)) bad = 1;
    free(h);
  }
  for (int k = 0; k < 6; ++k) if (!isfinite(rgb[k])) bad = 1;
  printf("kernels launched by the library: %lld\n", (long long)nefes_launch_count());
  if (bad) { fprintf(stderr, "non-finite or empty results\n"); return 4; }
  printf("ok\n");
  return 0;
}
