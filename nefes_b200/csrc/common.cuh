// Shared host/device helpers for libnefes_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nefes_b200.h"

namespace nefes {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// optional per-kernel timing (nefes_prof_enable): CUDA events on the launching stream around a launch, together with
// the launch's ALGORITHMIC bytes and flops, so a roofline fraction can be reported per kernel (bench.py)
void prof_begin(const char* tag, cudaStream_t st, double alg_bytes, double alg_flops);
void prof_end(cudaStream_t st);

// Forward-only field queries (a render under no_grad, the sigma-only coarse pass of a test-time render): while set on the
// calling thread, the bf16 forward chain does not write the saved activation copies (nobody will run the backward).
void set_forward_only(bool on);
bool forward_only();
// While set, the bf16 forward does not re-pack the weights: the operand images in the caller's `saved` workspace are the
// ones an earlier call wrote from the same (unchanged) parameters (nefes_render_cfg_t.weights_packed).
void set_weights_packed(bool on);
bool weights_packed();
struct WeightsPackedScope {
  bool prev;
  explicit WeightsPackedScope(bool on) : prev(weights_packed()) { set_weights_packed(on); }
  ~WeightsPackedScope() { set_weights_packed(prev); }
};
struct ForwardOnlyScope {
  bool prev;
  explicit ForwardOnlyScope(bool on) : prev(forward_only()) { set_forward_only(on); }
  ~ForwardOnlyScope() { set_forward_only(prev); }
};

#define NEFES_REQUIRE(cond, code, ...)                 \
  do {                                                 \
    if (!(cond)) {                                     \
      ::nefes::set_error(__VA_ARGS__);                 \
      return (code);                                   \
    }                                                  \
  } while (0)

// check the launch that was just made
#define NEFES_CHECK_LAUNCH(what)                                                  \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      ::nefes::set_error("%s: %s", (what), cudaGetErrorString(e__));              \
      return NEFES_ECUDA;                                                         \
    }                                                                             \
    ::nefes::count_launch();                                                      \
  } while (0)

#define NEFES_CUDA(call)                                                          \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::nefes::set_error("%s: %s", #call, cudaGetErrorString(e__));               \
      return NEFES_ECUDA;                                                         \
    }                                                                             \
  } while (0)

#ifdef __CUDACC__
#define NEFES_HD __host__ __device__
#else
#define NEFES_HD
#endif
NEFES_HD static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
NEFES_HD static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- architecture constants (script/models/nerfh_nff.py:421-505, options.py) ------------
constexpr int kW = 128;         // hidden width
constexpr int kXyzCh = 63;      // 3 + 3*2*10
constexpr int kDirCh = 27;      // 3 + 3*2*4
constexpr int kXyzFreqs = 10;
constexpr int kDirFreqs = 4;
constexpr int kHeadCh = 131;    // 3 rgb + 128 feature
constexpr int kFeat = 128;

// layer indices in the flat parameter buffer (see nefes_param_layout)
enum Layer {
  L_T0 = 0, L_T1, L_T2, L_T3, L_T4, L_T5, L_T6, L_T7,   // xyz_encoding_1..8
  L_FINAL = 8,      // xyz_encoding_final  128 <- 128
  L_SIGMA = 9,      // static_sigma.0        1 <- 128
  L_DIR = 10,       // dir_encoding.0       64 <- 155
  L_TENC0 = 11,     // transient_encoding.0 64 <- 155   (fine only)
  L_RGB = 12,       // static_rgb.0        131 <- 64
  L_TENC1 = 13,     // transient_encoding.2 64 <- 64    (fine only)
  L_TENC2 = 14,     // transient_encoding.4 64 <- 64    (fine only)
  L_TRGB = 15,      // transient_rgb.0       3 <- 64    (fine only)
  L_TSIG = 16,      // transient_sigma.0     1 <- 64    (fine only)
  L_TBETA = 17,     // transient_beta.0      1 <- 64    (fine only)
};

struct Layout {
  nefes_layout_t c;                       // public mirror
  int64_t w[NEFES_MAX_LAYERS];            // indexed by Layer (-1 if absent)
  int64_t b[NEFES_MAX_LAYERS];
};
const Layout& layout_for(int net);

#ifdef __CUDACC__
__device__ __forceinline__ float softplus_f(float x) {          // nn.Softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace nefes
