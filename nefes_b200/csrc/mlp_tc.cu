// K5, NEFES_PREC_BF16 path (tcgen05 / TMEM).  Placeholder until the tensor-core kernels land:
// the entry points fail loudly rather than fall back.
#include "common.cuh"

namespace nefes {
int mlp_workspace_bf16(int, int, int64_t, int64_t, int64_t*, int64_t*, int64_t*) {
  set_error("NEFES_PREC_BF16 is not built yet");
  return NEFES_EUNSUPPORTED;
}
int mlp_fwd_bf16(const float*, int, int, const float*, const float*, int64_t, int, float*, void*, void*, cudaStream_t) {
  set_error("NEFES_PREC_BF16 is not built yet");
  return NEFES_EUNSUPPORTED;
}
int mlp_bwd_bf16(const float*, int, int, const float*, const float*, int64_t, int, const float*, const float*,
                 const void*, void*, float*, float*, float*, cudaStream_t) {
  set_error("NEFES_PREC_BF16 is not built yet");
  return NEFES_EUNSUPPORTED;
}
}  // namespace nefes
