// K4b encoder front-end B: multiresolution HashGrid (tiny-cuda-nn algorithm; the reference reaches it only
// through tcnn.Encoding at script/models/nerfh_tcnn.py:65-75) and degree-4 spherical harmonics (:97-103).
// Gather-bound: 16 levels x 8 corners x 2 features per point; the table (46.5 MB fp32 at T = 2^19) is
// L2-resident on B200.  One thread per (point, level): the 16 threads of a point are adjacent lanes, so the
// [M, 32] output row is one coalesced 128-byte store and the input-gradient reduction is a half-warp shuffle.
// Backward scatters with vector fp32 atomics (red.global.add.v2.f32).
#include "common.cuh"

namespace nefes {

__device__ __forceinline__ uint32_t grid_index(const nefes_hash_level_t& lv, uint32_t gx, uint32_t gy, uint32_t gz) {
  uint32_t idx;
  if (lv.dense) idx = gx + gy * lv.res + gz * lv.res * lv.res;
  else idx = (gx * 1u) ^ (gy * 2654435761u) ^ (gz * 805459861u);
  return idx % lv.size;
}

template <bool BWD>
__global__ void hash_kernel(const float* __restrict__ x, const float* __restrict__ table, const float* __restrict__ d_out,
                            int64_t M, nefes_hash_layout_t L, float* __restrict__ out, float* __restrict__ d_table,
                            float* __restrict__ d_x) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = idx / L.n_levels;
  const int l = (int)(idx % L.n_levels);
  const bool ok = m < M;
  const nefes_hash_level_t lv = L.level[l];
  float px = 0.f, py = 0.f, pz = 0.f;
  if (ok) { px = x[m * 3]; py = x[m * 3 + 1]; pz = x[m * 3 + 2]; }
  px = fmaf(lv.scale, px, 0.5f); py = fmaf(lv.scale, py, 0.5f); pz = fmaf(lv.scale, pz, 0.5f);
  const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
  const float wx = px - fx, wy = py - fy, wz = pz - fz;
  const float2* tab = reinterpret_cast<const float2*>(table) + lv.offset;
  float2 g = make_float2(0.f, 0.f);
  if (BWD && ok) g = *reinterpret_cast<const float2*>(d_out + m * (2 * L.n_levels) + 2 * l);
  float2 acc = make_float2(0.f, 0.f);
  float dx = 0.f, dy = 0.f, dz = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
    const float ax = bx ? wx : 1.f - wx, ay = by ? wy : 1.f - wy, az = bz ? wz : 1.f - wz;
    const uint32_t e = grid_index(lv, gx + bx, gy + by, gz + bz);
    if (!ok) continue;
    const float w = ax * ay * az;
    if (!BWD) {
      const float2 v = __ldg(tab + e);
      acc.x = fmaf(w, v.x, acc.x);
      acc.y = fmaf(w, v.y, acc.y);
    } else {
      if (d_table != nullptr) atomicAdd(reinterpret_cast<float2*>(d_table) + lv.offset + e, make_float2(w * g.x, w * g.y));
      if (d_x != nullptr) {
        const float2 v = __ldg(tab + e);
        const float dot = v.x * g.x + v.y * g.y;
        dx += (bx ? 1.f : -1.f) * ay * az * dot;
        dy += (by ? 1.f : -1.f) * ax * az * dot;
        dz += (bz ? 1.f : -1.f) * ax * ay * dot;
      }
    }
  }
  if (!BWD) {
    if (ok) *reinterpret_cast<float2*>(out + m * (2 * L.n_levels) + 2 * l) = acc;
  } else if (d_x != nullptr) {
    dx *= lv.scale; dy *= lv.scale; dz *= lv.scale;
    // sum over the levels of this point: n_levels (a power of two <= 32) adjacent lanes
    for (int o = L.n_levels >> 1; o > 0; o >>= 1) {
      dx += __shfl_xor_sync(0xffffffffu, dx, o);
      dy += __shfl_xor_sync(0xffffffffu, dy, o);
      dz += __shfl_xor_sync(0xffffffffu, dz, o);
    }
    if (ok && l == 0) { d_x[m * 3] = dx; d_x[m * 3 + 1] = dy; d_x[m * 3 + 2] = dz; }
  }
}

__device__ __forceinline__ void sh4(float x, float y, float z, float* o) {
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y; o[2] = 0.48860251190291987f * z; o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy; o[5] = -1.0925484305920792f * yz; o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz; o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.f * x2 + y2); o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.f - 5.f * z2); o[12] = 0.3731763325901154f * z * (5.f * z2 - 3.f);
  o[13] = 0.45704579946446572f * x * (1.f - 5.f * z2); o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.f * y2);
}

__global__ void sh_fwd_kernel(const float* __restrict__ d, int64_t M, float* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float o[16];
  sh4(d[m * 3] * 2.f - 1.f, d[m * 3 + 1] * 2.f - 1.f, d[m * 3 + 2] * 2.f - 1.f, o);
  float4* dst = reinterpret_cast<float4*>(out + m * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

__global__ void sh_bwd_kernel(const float* __restrict__ d, const float* __restrict__ g_out, int64_t M, float* __restrict__ d_d) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float x = d[m * 3] * 2.f - 1.f, y = d[m * 3 + 1] * 2.f - 1.f, z = d[m * 3 + 2] * 2.f - 1.f;
  const float* g = g_out + m * 16;
  const float x2 = x * x, y2 = y * y, z2 = z * z;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  gy += -0.48860251190291987f * g[1]; gz += 0.48860251190291987f * g[2]; gx += -0.48860251190291987f * g[3];
  gx += 1.0925484305920792f * y * g[4]; gy += 1.0925484305920792f * x * g[4];
  gy += -1.0925484305920792f * z * g[5]; gz += -1.0925484305920792f * y * g[5];
  gz += 2.f * 0.94617469575755997f * z * g[6];
  gx += -1.0925484305920792f * z * g[7]; gz += -1.0925484305920792f * x * g[7];
  gx += 2.f * 0.54627421529603959f * x * g[8]; gy += -2.f * 0.54627421529603959f * y * g[8];
  gx += 0.59004358992664352f * y * (-6.f * x) * g[9]; gy += 0.59004358992664352f * (-3.f * x2 + 3.f * y2) * g[9];
  gx += 2.8906114426405538f * y * z * g[10]; gy += 2.8906114426405538f * x * z * g[10]; gz += 2.8906114426405538f * x * y * g[10];
  gy += 0.45704579946446572f * (1.f - 5.f * z2) * g[11]; gz += 0.45704579946446572f * y * (-10.f * z) * g[11];
  gz += 0.3731763325901154f * (15.f * z2 - 3.f) * g[12];
  gx += 0.45704579946446572f * (1.f - 5.f * z2) * g[13]; gz += 0.45704579946446572f * x * (-10.f * z) * g[13];
  gx += 1.4453057213202769f * z * 2.f * x * g[14]; gy += -1.4453057213202769f * z * 2.f * y * g[14]; gz += 1.4453057213202769f * (x2 - y2) * g[14];
  gx += 0.59004358992664352f * (-3.f * x2 + 3.f * y2) * g[15]; gy += 0.59004358992664352f * x * 6.f * y * g[15];
  d_d[m * 3] = 2.f * gx; d_d[m * 3 + 1] = 2.f * gy; d_d[m * 3 + 2] = 2.f * gz;     // v = 2 d - 1
}

}  // namespace nefes

extern "C" {

int nefes_hash_layout(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale,
                      nefes_hash_layout_t* out_host) {
  NEFES_REQUIRE(out_host, NEFES_EINVAL, "nefes_hash_layout: null output");
  NEFES_REQUIRE(n_levels >= 1 && n_levels <= NEFES_HASH_MAX_LEVELS && (n_levels & (n_levels - 1)) == 0, NEFES_EINVAL,
                "nefes_hash_layout: n_levels must be a power of two <= 32 (got %d)", n_levels);
  NEFES_REQUIRE(log2_hashmap_size >= 8 && log2_hashmap_size <= 28 && base_resolution >= 2 && per_level_scale >= 1.f,
                NEFES_EINVAL, "nefes_hash_layout: bad configuration");
  const float log2_pls = log2f(per_level_scale);
  int64_t off = 0;
  out_host->n_levels = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    nefes_hash_level_t& lv = out_host->level[l];
    lv.scale = exp2f((float)l * log2_pls) * (float)base_resolution - 1.0f;
    lv.res = (uint32_t)ceilf(lv.scale) + 1u;
    uint64_t n = (uint64_t)lv.res * lv.res * lv.res;
    if (n > 0x7fffffffull) n = 0x7fffffffull;
    n = (n + 7) / 8 * 8;
    const uint64_t cap = 1ull << log2_hashmap_size;
    lv.dense = ((uint64_t)lv.res * lv.res * lv.res <= (n < cap ? n : cap)) ? 1u : 0u;
    if (n > cap) n = cap;
    lv.size = (uint32_t)n;
    lv.offset = (uint32_t)off;
    off += (int64_t)n;
  }
  out_host->n_entries = off;
  return NEFES_OK;
}

int nefes_encode_hash_fwd(const float* x, const float* table, int64_t M, const nefes_hash_layout_t* layout_host,
                          float* out, void* stream) {
  NEFES_REQUIRE(layout_host, NEFES_EINVAL, "nefes_encode_hash_fwd: null layout");
  if (M == 0) return NEFES_OK;
  NEFES_REQUIRE(x && table && out && M > 0, NEFES_EINVAL, "nefes_encode_hash_fwd: null pointer");
  const int64_t n = M * layout_host->n_levels;
  nefes::hash_kernel<false><<<(unsigned)nefes::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      x, table, nullptr, M, *layout_host, out, nullptr, nullptr);
  NEFES_CHECK_LAUNCH("hash_fwd");
  return NEFES_OK;
}

int nefes_encode_hash_bwd(const float* x, const float* d_out, const float* table, int64_t M,
                          const nefes_hash_layout_t* layout_host, float* d_table, float* d_x, void* stream) {
  NEFES_REQUIRE(layout_host, NEFES_EINVAL, "nefes_encode_hash_bwd: null layout");
  if (M == 0) return NEFES_OK;
  NEFES_REQUIRE(x && d_out && table && (d_table || d_x), NEFES_EINVAL, "nefes_encode_hash_bwd: null pointer");
  const int64_t n = M * layout_host->n_levels;
  nefes::hash_kernel<true><<<(unsigned)nefes::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      x, table, d_out, M, *layout_host, nullptr, d_table, d_x);
  NEFES_CHECK_LAUNCH("hash_bwd");
  return NEFES_OK;
}

int nefes_encode_sh_fwd(const float* d, int64_t M, float* out, void* stream) {
  if (M == 0) return NEFES_OK;
  NEFES_REQUIRE(d && out && M > 0, NEFES_EINVAL, "nefes_encode_sh_fwd: null pointer");
  nefes::sh_fwd_kernel<<<(unsigned)nefes::ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(d, M, out);
  NEFES_CHECK_LAUNCH("sh_fwd");
  return NEFES_OK;
}

int nefes_encode_sh_bwd(const float* d, const float* d_out, int64_t M, float* d_d, void* stream) {
  if (M == 0) return NEFES_OK;
  NEFES_REQUIRE(d && d_out && d_d && M > 0, NEFES_EINVAL, "nefes_encode_sh_bwd: null pointer");
  nefes::sh_bwd_kernel<<<(unsigned)nefes::ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(d, d_out, M, d_d);
  NEFES_CHECK_LAUNCH("sh_bwd");
  return NEFES_OK;
}

}  // extern "C"
