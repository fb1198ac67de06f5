// Library-level entry points: version, error text, launch counter, flat parameter layout.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace nefes {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static thread_local bool t_forward_only = false;
void set_forward_only(bool on) { t_forward_only = on; }
bool forward_only() { return t_forward_only; }
static thread_local bool t_weights_packed = false;
void set_weights_packed(bool on) { t_weights_packed = on; }
bool weights_packed() { return t_weights_packed; }

namespace {
struct ProfRec { const char* tag; cudaEvent_t e0, e1; double bytes, flops; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}  // namespace
void prof_begin(const char* tag, cudaStream_t st, double alg_bytes, double alg_flops) {
  if (!g_prof_on) return;
  ProfRec r = {tag, nullptr, nullptr, alg_bytes, alg_flops};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().e1, st);
}

namespace {
struct Row { Layer id; const char* name; int out, in; bool fine_only; };
// Order of the flat buffer.  Names are the reference's state_dict prefixes
// (script/models/nerfh_nff.py:469-505).  Layers that read the same input
// are adjacent so a kernel may treat them as one matrix: (final, sigma) <- h8,
// (dir, tenc0) <- [final | dirPE], (t_rgb, t_sigma, t_beta) <- t3 in raw-column order 132..136.
const Row kRows[NEFES_MAX_LAYERS] = {
    {L_T0, "xyz_encoding_1.0", 128, 63, false},   {L_T1, "xyz_encoding_2.0", 128, 128, false},
    {L_T2, "xyz_encoding_3.0", 128, 128, false},  {L_T3, "xyz_encoding_4.0", 128, 128, false},
    {L_T4, "xyz_encoding_5.0", 128, 191, false},  {L_T5, "xyz_encoding_6.0", 128, 128, false},
    {L_T6, "xyz_encoding_7.0", 128, 128, false},  {L_T7, "xyz_encoding_8.0", 128, 128, false},
    {L_FINAL, "xyz_encoding_final", 128, 128, false}, {L_SIGMA, "static_sigma.0", 1, 128, false},
    {L_DIR, "dir_encoding.0", 64, 155, false},    {L_TENC0, "transient_encoding.0", 64, 155, true},
    {L_RGB, "static_rgb.0", 131, 64, false},      {L_TENC1, "transient_encoding.2", 64, 64, true},
    {L_TENC2, "transient_encoding.4", 64, 64, true}, {L_TRGB, "transient_rgb.0", 3, 64, true},
    {L_TSIG, "transient_sigma.0", 1, 64, true},   {L_TBETA, "transient_beta.0", 1, 64, true},
};

Layout build(int net) {
  Layout L;
  memset(&L, 0, sizeof(L));
  for (int i = 0; i < NEFES_MAX_LAYERS; ++i) L.w[i] = L.b[i] = -1;
  int n = 0;
  int64_t off = 0;
  for (const Row& r : kRows) {
    if (r.fine_only && net != NEFES_NET_FINE) continue;
    L.c.out_dim[n] = r.out;
    L.c.in_dim[n] = r.in;
    L.c.name[n] = r.name;
    L.c.w_off[n] = off;
    L.w[r.id] = off;
    off += (int64_t)r.out * r.in;
    ++n;
  }
  n = 0;
  for (const Row& r : kRows) {
    if (r.fine_only && net != NEFES_NET_FINE) continue;
    L.c.b_off[n] = off;
    L.b[r.id] = off;
    off += r.out;
    ++n;
  }
  L.c.n_layers = n;
  L.c.n_params = off;
  return L;
}
}  // namespace

const Layout& layout_for(int net) {
  static const Layout coarse = build(NEFES_NET_COARSE);
  static const Layout fine = build(NEFES_NET_FINE);
  return net == NEFES_NET_FINE ? fine : coarse;
}

}  // namespace nefes

extern "C" {

int nefes_version(void) { return NEFES_VERSION; }
const char* nefes_last_error(void) { return nefes::g_err; }
int64_t nefes_launch_count(void) { return nefes::g_launches.load(); }

int nefes_prof_enable(int on) {
  for (auto& r : nefes::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  nefes::g_prof.clear();
  nefes::g_prof_on = on != 0;
  return NEFES_OK;
}

int nefes_prof_report(char* buf, int cap) {
  NEFES_REQUIRE(buf != nullptr && cap > 2, NEFES_EINVAL, "nefes_prof_report: bad buffer");
  cudaDeviceSynchronize();
  struct Agg { int n = 0; double ms = 0, bytes = 0, flops = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : nefes::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    Agg& a = agg[r.tag];
    a.n += 1; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
  }
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %d, \"ms\": %.6f, \"alg_bytes\": %.0f, \"alg_flops\": %.0f}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes, kv.second.flops);
    out += line;
    first = false;
  }
  out += "}";
  NEFES_REQUIRE((int)out.size() + 1 <= cap, NEFES_EINVAL, "nefes_prof_report: buffer too small (%d needed)", (int)out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return NEFES_OK;
}

int nefes_param_layout(int net, nefes_layout_t* out_host) {
  NEFES_REQUIRE(out_host != nullptr, NEFES_EINVAL, "nefes_param_layout: null output");
  NEFES_REQUIRE(net == NEFES_NET_COARSE || net == NEFES_NET_FINE, NEFES_EINVAL,
                "nefes_param_layout: net must be 0 (coarse) or 1 (fine), got %d", net);
  *out_host = nefes::layout_for(net).c;
  return NEFES_OK;
}

}  // extern "C"
