// HiFT vocoder: handle (weight packing), per-batch plan (ragged packing, tile tables), forward
// orchestration and inspection taps.  Upstream chatterbox-tts==0.1.6 hifigan.py
// HiFTGenerator.inference / decode + s3gen.py trim_fade, reached in the reference through
// tts_backends/chatterbox_impl.py:189.
#include "vt_hift.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

namespace vt {

int launch_conv_tc(const ConvArgs& a, const ConvLayer& L, int act_elem, const void* tc_tiles, int n_tc_tiles,
                   int tile_rows, cudaStream_t st);
int pack_conv_tc(ConvLayer& L, const std::vector<float>& w_kcico, int act_elem, std::vector<void*>& allocs);
bool conv_tc_supported(const ConvLayer& L);
int conv_tc_tile_rows(const ConvLayer& L);
bool convT_tc_supported(const ConvLayer& L);
int launch_convT_tc(const ConvArgs& a, const ConvLayer& L, int act_elem, const void* tiles, int n_tiles, cudaStream_t st);
bool pair_tc_supported(const ConvLayer& c1, const ConvLayer& c2);
int pair_tc_tile_rows(int k);
bool pair3_tc_supported(const ConvLayer* const c1[3], const ConvLayer* const c2[3], int act_elem);
int launch_pair3_tc(const ConvArgs& a, const ConvLayer* const c1[3], const ConvLayer* const c2[3], const float* const alpha1[3],
                    const float* const alpha2[3], const float* const x_in[3], int act_elem, cudaStream_t st);
bool pair64_tc_supported(const ConvLayer& c1, const ConvLayer& c2);
int pair64_tc_tile_rows(int k);
int pack_pair64(ConvLayer& L, std::vector<void*>& allocs);
int launch_pair64_tc(const ConvArgs& a, const ConvLayer& c1, const ConvLayer& c2, const float* alpha1, const float* alpha2,
                     int act_elem, cudaStream_t st);
int launch_pair_tc(const ConvArgs& a, const ConvLayer& c1, const ConvLayer& c2, const float* alpha1, const float* alpha2,
                   int act_elem, cudaStream_t st);

namespace {

const int kRbKernels[3] = {3, 7, 11};
const int kRbDil[3] = {1, 3, 5};

struct HostTensor {
  const float* data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

}  // namespace

struct Plan {
  int B = 0;
  std::vector<int> T;
  long long total_T = 0;
  int T_max = 0;
  long long rows[3] = {0, 0, 0};            // packed rows per level (incl. gaps, excl. slack)
  long long rowsM = 0;                      // packed rows of the gapped mel-rate level (conv_pre output)
  // device tables
  int* d_T = nullptr;
  int* d_mel_off = nullptr;
  long long* d_off[3] = {nullptr, nullptr, nullptr};
  long long* d_offM = nullptr;
  ConvTile* d_tiles = nullptr;
  // tile-table segments (offset, count) inside d_tiles
  struct Seg { int off = 0, n = 0; };
  Seg mel, pre, up[3], sd[3], lvl[3];
  Seg tc[3][2];                             // tensor-core tiles per level, [0]: 128 rows, [1]: 256 rows
  Seg tcu[3];                               // tensor-core tiles of the transposed convs (128 input steps)
  Seg pair[3][3];                           // fused ResBlock-pair tiles per level and kernel size 3 / 7 / 11
  Seg pair64[3];                            // ... of the tap-paired kernel (C = 64 level)
  Seg g_mel, g_melu, g_sd[3];               // 256-step tiles of the K-blocked kernel: gapped mel -> gapped mel,
                                            // gapped mel -> ungapped mel, level-2 STFT rows -> level l
  std::vector<long long> h_off[3], h_offM;
  std::vector<int> h_mel_off;
  void* d_block = nullptr;
  size_t d_block_bytes = 0;
  void* h_block = nullptr;                  // pinned staging of d_block (the upload is asynchronous)
  size_t h_block_bytes = 0;
  cudaEvent_t done = nullptr;               // recorded after the last forward that used this plan
  unsigned long long last_use = 0;
};

struct Workspace {
  float *f0a, *f0b, *f0, *s, *spec, *xpre, *post;
  double* phase_base;
  float *U[3], *S[3], *X[3], *XR[3], *Y[3];
  float *S2[3], *XR2[3];                    // ping-pong partners of S / XR for the fused pairs (no in-place update)
  float *XR3[3], *XR4[3];                   // with XR2: the inputs of the three ResBlocks' last pairs (mean-fused launch)
  void *A[3][4];                            // activation copies A0, A1, A2, Q per level
  void *Yact[3];                            // leaky_relu(stage output) operand copies (next ups / conv_post)
  void *xpre_act;                           // leaky_relu(conv_pre) operand copy, gapped mel-rate layout
  void *mel_hi, *mel_lo;                    // fp16 split of the mel, gapped mel-rate layout, kMelOp channels
  void *fx[2][2];                           // F0 trunk activations, fp16 split [ping-pong][hi, lo], gapped mel-rate layout
  void *spec_op;                            // STFT operand rows (kSpecOp elements) for the tensor-core source_downs
  long long cap_rows[3], cap_rowsM;
  size_t bytes;
};

}  // namespace vt

using namespace vt;

struct vt_hift {
  HiftCfg cfg;
  int act_elem = ELEM_F32;
  bool use_tc = false;
  std::vector<void*> allocs;
  ConvLayer conv_pre, ups[3], sdown[3], src_c1[3][3], src_c2[3][3], rb_c1[9][3], rb_c2[9][3], conv_post, f0c[5];
  float *src_a1[3][3], *src_a2[3][3], *rb_a1[9][3], *rb_a2[9][3];
  float *f0_w = nullptr, *f0_b = nullptr, *lin_w = nullptr, *lin_b = nullptr, *trim_fade = nullptr;
  Plan plan;                                // the plan of the current / last forward
  std::vector<Plan> plan_cache;             // recently used plans (a long job runs a handful of bucket shapes over and over)
  unsigned long long plan_tick = 0;
  Workspace ws{};
  bool fuse[3] = {false, false, false};     // level runs its ResBlock pairs on the fused kernel (vt_pair_tc.cu)
  bool pair64_last = false;                 // VT_PAIR64=all: also the last pair of each ResBlock
  bool pair64[3] = {false, false, false};   // ... C = 64 level: kernel size index kk runs on the tap-paired kernel (vt_pair64_tc.cu)
  bool have_forward = false;
  // profiling (vt_hift_set_profiling): events around the whole forward and around each stage's
  // run of resblock convolutions (the dominant kernel class)
  bool profiling = false;
  cudaEvent_t ev_fwd[2] = {nullptr, nullptr};
  cudaEvent_t ev_rb[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  double rb_flops = 0;
  int rb_launches = 0;
  // named timeline marks of the last forward (profiling only): events are created once and reused
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  size_t n_marks = 0;
};

namespace vt {
namespace {

size_t elem_size(int e) { return e == ELEM_F32 ? 4 : 2; }

int dev_upload(vt_hift* h, const void* src, size_t bytes, void** out) {
  void* p = nullptr;
  VT_CUDA_OK(cudaMalloc(&p, bytes ? bytes : 4));
  h->allocs.push_back(p);
  if (bytes) VT_CUDA_OK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
  *out = p;
  return VT_OK;
}

// Per-output-channel power-of-two scaling of the tensor-core weight images.  fp16 operands lose precision below
// 6.1e-5 (subnormals) and a trained checkpoint's weight-normed rows can sit orders of magnitude apart (w = g v / ||v||
// with a per-row gain g).  A LAYER is scaled only if one of its rows needs it - max|row| below 2^-8 (its typical elements
// then approach the subnormal range) or above 2^10: then every row c holds s_c * w with s_c = 2^-e the power of two that
// brings max|row| into [0.5, 1) - exact in every format - the accumulators come out as s_c * (W a) and the epilogues undo
// it in the FMA that adds the bias (ConvArgs::wscale = 1 / s_c).  Layers inside the comfortable range keep s = 1 and their
// kernels' unscaled instances.  `w` is [k][cin_pad][ncol] and is scaled IN PLACE (callers pass a copy: the fp32 CUDA-core
// path keeps the plain weights).
int scale_weight_rows(vt_hift* h, ConvLayer& L, std::vector<float>& w, int k, int cin_pad, int ncol) {
  std::vector<float> inv(ncol, 1.0f), rowmax(ncol, 0.0f);
  bool need = false;
  for (int co = 0; co < ncol; ++co) {
    float m = 0.0f;
    for (int j = 0; j < k; ++j)
      for (int ci = 0; ci < cin_pad; ++ci) m = std::max(m, std::fabs(w[((size_t)j * cin_pad + ci) * ncol + co]));
    rowmax[co] = m;
    if (m > 0.0f && std::isfinite(m) && (m < 0.00390625f || m > 1024.0f)) need = true;
  }
  static const char* force = getenv("VT_WSCALE");               // VT_WSCALE=1: scale every layer (tests), 0: never
  if (force && force[0] == '1') need = true;
  if (force && force[0] == '0') need = false;
  L.scaled = need;
  for (int co = 0; co < ncol && need; ++co) {
    const float m = rowmax[co];
    if (!(m > 0.0f) || !std::isfinite(m)) continue;
    int e = 0;
    std::frexp(m, &e);                       // m = f * 2^e, f in [0.5, 1)
    e = std::max(-100, std::min(e, 100));
    const float s = std::ldexp(1.0f, -e);
    inv[co] = std::ldexp(1.0f, e);
    for (int j = 0; j < k; ++j)
      for (int ci = 0; ci < cin_pad; ++ci) w[((size_t)j * cin_pad + ci) * ncol + co] *= s;
  }
  return dev_upload(h, inv.data(), inv.size() * 4, (void**)&L.wscale);
}

// conv weight [cout][cin][k] -> [k][cin_pad][cout_pad]
// K-blocked tensor-core packing of a layer (vt_gemm_tc.cu).  `w` is [k][cin_pad][cout].
//   GEMM_CONV  : one K block per (tap, 64-channel block) of an operand buffer with `op_ld` channels
//   GEMM_SPLIT : the same three times - (x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi) - always fp16
//   GEMM_IM2COL: strided conv over operand rows of `op_ld` elements: K blocks walk the k*op_ld contiguous run
enum GemmMode { GEMM_NONE = 0, GEMM_CONV = 1, GEMM_SPLIT = 2, GEMM_IM2COL = 3 };

int pack_gemm(vt_hift* h, ConvLayer& L, const std::vector<float>& w, int cin_real, int cin_pad, int mode, int op_ld) {
  const int N = L.cout, k = L.k;
  std::vector<KBlock> kbs;
  std::vector<int> part;
  std::vector<float> wkb;
  auto add_block = [&](int src, int shift, int ch_off, int prt, auto&& weight_of /* (e, co) -> float */) {
    kbs.push_back(KBlock{src, shift, ch_off, 0});
    part.push_back(prt);
    const size_t base = wkb.size();
    wkb.resize(base + (size_t)64 * N);
    for (int e = 0; e < 64; ++e)
      for (int co = 0; co < N; ++co) wkb[base + (size_t)e * N + co] = weight_of(e, co);
  };
  if (mode == GEMM_IM2COL) {
    const int K = k * op_ld;
    for (int cb = 0; cb * 64 < K; ++cb)
      add_block(0, -L.pad, cb * 64, 0, [&](int e, int co) {
        const int q = cb * 64 + e, j = q / op_ld, ci = q - j * op_ld;
        return (j < k && ci < cin_real) ? w[((size_t)j * cin_pad + ci) * N + co] : 0.0f;
      });
  } else {
    const int terms = mode == GEMM_SPLIT ? 3 : 1;
    for (int t = 0; t < terms; ++t)
      for (int j = 0; j < k; ++j)
        for (int cb = 0; cb * 64 < op_ld; ++cb)
          add_block(t == 2 ? 1 : 0, j * L.dil - L.pad, cb * 64, t == 1 ? 1 : 0, [&](int e, int co) {
            const int ci = cb * 64 + e;
            return ci < cin_real ? w[((size_t)j * cin_pad + ci) * N + co] : 0.0f;
          });
  }
  const int elem = mode == GEMM_SPLIT ? (int)ELEM_F16 : h->act_elem;
  return pack_gemm_tc(L, kbs, part, wkb, N % 128 == 0 ? 128 : 64, elem, h->allocs);
}

int pack_conv(vt_hift* h, ConvLayer& L, const std::map<std::string, HostTensor>& tab, const std::string& name,
              int cin, int cout, int k, int dil, int stride, int pad, int cin_pad, int cout_pad,
              int gemm_mode = GEMM_NONE, int op_ld = 0) {
  auto wi = tab.find(name + ".weight"), bi = tab.find(name + ".bias");
  VT_REQUIRE(wi != tab.end() && bi != tab.end(), "missing tensor %s.weight/.bias", name.c_str());
  const HostTensor& W = wi->second;
  VT_REQUIRE(W.shape.size() == 3 && W.shape[0] == cout && W.shape[1] == cin && W.shape[2] == k,
             "%s.weight has the wrong shape (want [%d,%d,%d])", name.c_str(), cout, cin, k);
  VT_REQUIRE(bi->second.numel() == cout, "%s.bias has the wrong size", name.c_str());
  std::vector<float> w((size_t)k * cin_pad * cout_pad, 0.0f), b(cout_pad, 0.0f);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int j = 0; j < k; ++j)
        w[((size_t)j * cin_pad + ci) * cout_pad + co] = W.data[((size_t)co * cin + ci) * k + j];
  for (int co = 0; co < cout; ++co) b[co] = bi->second.data[co];
  L.name = name;
  L.cin = cin_pad; L.cout = cout_pad; L.k = k; L.dil = dil; L.stride = stride; L.pad = pad;
  L.out_mul = 1; L.phase_c = cout_pad;
  L.flops_per_step = 2.0 * cin * cout * k;
  int rc = dev_upload(h, w.data(), w.size() * 4, (void**)&L.w);
  if (rc) return rc;
  rc = dev_upload(h, b.data(), b.size() * 4, (void**)&L.bias);
  if (rc) return rc;
  if (!h->use_tc) return VT_OK;
  rc = scale_weight_rows(h, L, w, k, cin_pad, cout_pad);       // w is a local copy: the upload above kept the plain weights
  if (rc) return rc;
  if (gemm_mode != GEMM_NONE) return pack_gemm(h, L, w, cin, cin_pad, gemm_mode, op_ld);
  if (conv_tc_supported(L)) {
    rc = pack_conv_tc(L, w, h->act_elem, h->allocs);
    if (rc) return rc;
    return pack_pair64(L, h->allocs);
  }
  return VT_OK;
}

// ConvTranspose1d weight [cin][cout][k], stride s, padding p -> 3-tap conv with cout' = s*cout:
//   out[s*q + r] = sum_{d in -1..1} x[q + d] . W[:, :, r + p - s*d]
int pack_convT(vt_hift* h, ConvLayer& L, const std::map<std::string, HostTensor>& tab, const std::string& name,
               int cin, int cout, int k, int s, int p) {
  auto wi = tab.find(name + ".weight"), bi = tab.find(name + ".bias");
  VT_REQUIRE(wi != tab.end() && bi != tab.end(), "missing tensor %s.weight/.bias", name.c_str());
  const HostTensor& W = wi->second;
  VT_REQUIRE(W.shape.size() == 3 && W.shape[0] == cin && W.shape[1] == cout && W.shape[2] == k,
             "%s.weight has the wrong shape (want [%d,%d,%d])", name.c_str(), cin, cout, k);
  const int cN = s * cout;
  std::vector<float> w((size_t)3 * cin * cN, 0.0f), b(cN, 0.0f);
  for (int d = -1; d <= 1; ++d)
    for (int r = 0; r < s; ++r) {
      const int j = r + p - s * d;
      if (j < 0 || j >= k) continue;
      for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co)
          w[((size_t)(d + 1) * cin + ci) * cN + r * cout + co] = W.data[((size_t)ci * cout + co) * k + j];
    }
  // every tap j in [0, k) must be reachable with d in {-1, 0, 1}
  for (int r = 0; r < s; ++r)
    for (int d = -3; d <= 3; ++d) {
      const int j = r + p - s * d;
      VT_REQUIRE(!(j >= 0 && j < k) || (d >= -1 && d <= 1), "%s: transposed conv does not fit 3 taps", name.c_str());
    }
  for (int r = 0; r < s; ++r)
    for (int co = 0; co < cout; ++co) b[r * cout + co] = bi->second.data[co];
  L.name = name;
  L.cin = cin; L.cout = cN; L.k = 3; L.dil = 1; L.stride = 1; L.pad = 1;
  L.out_mul = s; L.phase_c = cout;
  L.flops_per_step = 2.0 * cin * cout * k;   // per INPUT step (= per s output steps)
  int rc = dev_upload(h, w.data(), w.size() * 4, (void**)&L.w);
  if (rc) return rc;
  rc = dev_upload(h, b.data(), b.size() * 4, (void**)&L.bias);
  if (rc) return rc;
  if (!h->use_tc) return VT_OK;
  rc = scale_weight_rows(h, L, w, 3, cin, cN);
  if (rc) return rc;
  if (conv_tc_supported(L)) return pack_conv_tc(L, w, h->act_elem, h->allocs);
  return VT_OK;
}

int upload_vec(vt_hift* h, const std::map<std::string, HostTensor>& tab, const std::string& name, int64_t n, float** out) {
  auto it = tab.find(name);
  VT_REQUIRE(it != tab.end(), "missing tensor %s", name.c_str());
  VT_REQUIRE(it->second.numel() == n, "%s has %lld elements, want %lld", name.c_str(), (long long)it->second.numel(), (long long)n);
  return dev_upload(h, it->second.data, (size_t)n * 4, (void**)out);
}

void add_tiles(std::vector<ConvTile>& v, Plan::Seg& seg, int B, const long long* in_row0, const int* in_len,
               const long long* out_row0, const int* out_len, int tile) {
  seg.off = (int)v.size();
  for (int b = 0; b < B; ++b)
    for (int q0 = 0; q0 < out_len[b]; q0 += tile) {
      ConvTile t;
      t.in_row0 = in_row0[b]; t.out_row0 = out_row0[b];
      t.in_len = in_len[b]; t.out_len = out_len[b];
      t.q0 = q0; t.n = std::min(tile, out_len[b] - q0);
      v.push_back(t);
    }
  seg.n = (int)v.size() - seg.off;
}

void free_plan(Plan& P) {
  if (P.d_block) cudaFree(P.d_block);
  if (P.h_block) cudaFreeHost(P.h_block);
  if (P.done) cudaEventDestroy(P.done);
  P = Plan();
}

constexpr size_t kPlanCache = 7;            // + the current one: 8 batch shapes stay resident

int build_plan(vt_hift* h, const int32_t* T, int B, cudaStream_t st) {
  auto matches = [&](const Plan& q) { return q.B == B && (int)q.T.size() == B && std::equal(q.T.begin(), q.T.end(), T); };
  if (matches(h->plan)) { h->plan.last_use = ++h->plan_tick; return VT_OK; }
  for (auto& c : h->plan_cache)
    if (matches(c)) { std::swap(h->plan, c); h->plan.last_use = ++h->plan_tick; return VT_OK; }
  // miss: the current plan retires into the cache; the least recently used one gives its blocks to the new plan
  if (h->plan.B > 0) h->plan_cache.push_back(std::move(h->plan));
  h->plan = Plan();
  if (h->plan_cache.size() > kPlanCache) {
    size_t lru = 0;
    for (size_t i = 1; i < h->plan_cache.size(); ++i)
      if (h->plan_cache[i].last_use < h->plan_cache[lru].last_use) lru = i;
    Plan old = std::move(h->plan_cache[lru]);
    h->plan_cache.erase(h->plan_cache.begin() + lru);
    if (old.done) VT_CUDA_OK(cudaEventSynchronize(old.done));      // kernels and the upload that used its blocks are finished
    h->plan.d_block = old.d_block; h->plan.d_block_bytes = old.d_block_bytes;
    h->plan.h_block = old.h_block; h->plan.h_block_bytes = old.h_block_bytes;
    h->plan.done = old.done;
  }
  Plan& P = h->plan;
  P.last_use = ++h->plan_tick;
  P.B = B;
  P.T.assign(T, T + B);
  P.total_T = 0;
  P.T_max = 0;
  P.h_mel_off.resize(B);
  std::vector<long long> melrow(B);
  const HiftCfg& cfg = h->cfg;
  const int NL = cfg.n_levels, LL = NL - 1;          // LL: the last (sample-rate / hop) level, one reflection-padded row longer
  std::vector<int> lenM(B), len[3];
  for (int l = 0; l < NL; ++l) { len[l].resize(B); P.h_off[l].resize(B); }
  for (int b = 0; b < B; ++b) {
    P.h_mel_off[b] = (int)P.total_T;
    melrow[b] = P.total_T;
    lenM[b] = T[b];
    P.total_T += T[b];
    P.T_max = std::max(P.T_max, (int)T[b]);
  }
  {
    P.h_offM.resize(B);
    long long o = kGap;
    for (int b = 0; b < B; ++b) {
      P.h_offM[b] = o;
      o += T[b] + kGap;
    }
    P.rowsM = o;
  }
  for (int l = 0; l < NL; ++l) {
    long long o = kGap;
    for (int b = 0; b < B; ++b) {
      len[l][b] = cfg.level_mul[l] * T[b] + (l == LL ? 1 : 0);
      P.h_off[l][b] = o;
      o += len[l][b] + kGap;
    }
    P.rows[l] = o;
  }
  std::vector<ConvTile> tiles;
  add_tiles(tiles, P.mel, B, melrow.data(), lenM.data(), melrow.data(), lenM.data(), kTileQ);
  // ups[i]: conv space = input steps; input is mel-level (i=0) or level i-1
  add_tiles(tiles, P.pre, B, melrow.data(), lenM.data(), P.h_offM.data(), lenM.data(), kTileQ);
  add_tiles(tiles, P.up[0], B, P.h_offM.data(), lenM.data(), P.h_off[0].data(), lenM.data(), kTileQ);
  add_tiles(tiles, P.tcu[0], B, P.h_offM.data(), lenM.data(), P.h_off[0].data(), lenM.data(), 128);
  for (int l = 1; l < NL; ++l) {
    add_tiles(tiles, P.tcu[l], B, P.h_off[l - 1].data(), len[l - 1].data(), P.h_off[l].data(), len[l - 1].data(), 128);
    add_tiles(tiles, P.up[l], B, P.h_off[l - 1].data(), len[l - 1].data(), P.h_off[l].data(), len[l - 1].data(), kTileQ);
  }
  add_tiles(tiles, P.g_mel, B, P.h_offM.data(), lenM.data(), P.h_offM.data(), lenM.data(), 256);
  add_tiles(tiles, P.g_melu, B, P.h_offM.data(), lenM.data(), melrow.data(), lenM.data(), 256);
  for (int l = 0; l < NL; ++l) {
    add_tiles(tiles, P.g_sd[l], B, P.h_off[LL].data(), len[LL].data(), P.h_off[l].data(), len[l].data(), 256);
    add_tiles(tiles, P.sd[l], B, P.h_off[LL].data(), len[LL].data(), P.h_off[l].data(), len[l].data(), kTileQ);
    add_tiles(tiles, P.lvl[l], B, P.h_off[l].data(), len[l].data(), P.h_off[l].data(), len[l].data(), kTileQ);
    add_tiles(tiles, P.tc[l][0], B, P.h_off[l].data(), len[l].data(), P.h_off[l].data(), len[l].data(), 128);
    add_tiles(tiles, P.tc[l][1], B, P.h_off[l].data(), len[l].data(), P.h_off[l].data(), len[l].data(), 256);
    for (int kk = 0; kk < 3; ++kk) {
      add_tiles(tiles, P.pair[l][kk], B, P.h_off[l].data(), len[l].data(), P.h_off[l].data(), len[l].data(),
                pair_tc_tile_rows(kRbKernels[kk]));
      if ((kBase >> (l + 1)) == 64)
        add_tiles(tiles, P.pair64[kk], B, P.h_off[l].data(), len[l].data(), P.h_off[l].data(), len[l].data(),
                  pair64_tc_tile_rows(kRbKernels[kk]));
    }
  }
  // one device block: T | mel_off | off[3] | tiles
  const size_t nI = align_up((size_t)B * 4, 256), nL = align_up((size_t)B * 8, 256);
  const size_t bytes = 2 * nI + 4 * nL + align_up(tiles.size() * sizeof(ConvTile), 256);
  // blocks grow geometrically (a recycled plan rarely needs a new allocation: cudaFree / cudaMalloc synchronise the device)
  const size_t cap = align_up(bytes + bytes / 2, (size_t)1 << 20);
  if (bytes > P.d_block_bytes) {
    if (P.d_block) cudaFree(P.d_block);
    P.d_block = nullptr;
    P.d_block_bytes = 0;
    VT_CUDA_OK(cudaMalloc(&P.d_block, cap));
    P.d_block_bytes = cap;
  }
  if (bytes > P.h_block_bytes) {
    if (P.h_block) cudaFreeHost(P.h_block);
    P.h_block = nullptr;
    P.h_block_bytes = 0;
    VT_CUDA_OK(cudaMallocHost(&P.h_block, cap));
    P.h_block_bytes = cap;
  }
  if (!P.done) VT_CUDA_OK(cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming));
  char* hp = (char*)P.h_block;
  std::memset(hp, 0, bytes);
  char* dp = (char*)P.d_block;
  std::memcpy(hp, T, (size_t)B * 4); P.d_T = (int*)dp;
  std::memcpy(hp + nI, P.h_mel_off.data(), (size_t)B * 4); P.d_mel_off = (int*)(dp + nI);
  for (int l = 0; l < NL; ++l) {
    std::memcpy(hp + 2 * nI + l * nL, P.h_off[l].data(), (size_t)B * 8);
    P.d_off[l] = (long long*)(dp + 2 * nI + l * nL);
  }
  std::memcpy(hp + 2 * nI + 3 * nL, P.h_offM.data(), (size_t)B * 8);
  P.d_offM = (long long*)(dp + 2 * nI + 3 * nL);
  std::memcpy(hp + 2 * nI + 4 * nL, tiles.data(), tiles.size() * sizeof(ConvTile));
  P.d_tiles = (ConvTile*)(dp + 2 * nI + 4 * nL);
  // asynchronous upload from the plan's own pinned staging block: a new batch shape costs host time only, no stream
  // synchronisation (a long job alternates between bucket shapes; `done` guards the blocks when a plan is recycled)
  VT_CUDA_OK(cudaMemcpyAsync(P.d_block, P.h_block, bytes, cudaMemcpyHostToDevice, st));
  return VT_OK;
}

Workspace carve_ws(const vt_hift* h, int B, long long total_T, void* base) {
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (void*)((char*)base + off) : nullptr;
    off += align_up((int64_t)bytes, 1024);
    return p;
  };
  const size_t es = elem_size(h->act_elem);
  const HiftCfg& cfg = h->cfg;
  const int NL = cfg.n_levels;
  const long long rowsM = total_T + 64;
  w.f0a = (float*)take((size_t)rowsM * kF0Ch * 4);
  w.f0b = (float*)take((size_t)rowsM * kF0Ch * 4);
  w.f0 = (float*)take((size_t)rowsM * 4);
  w.phase_base = (double*)take((size_t)kHarm * rowsM * 16);   // phase prefix | increment, per (harmonic, frame)
  w.s = (float*)take((size_t)total_T * cfg.spf * 4 + 64);
  w.cap_rowsM = total_T + (long long)B * kGap + kGap + 512;
  w.xpre = (float*)take((size_t)w.cap_rowsM * kBase * 4);
  w.xpre_act = take((size_t)w.cap_rowsM * kBase * es);
  w.mel_hi = take((size_t)w.cap_rowsM * kMelOp * 2);
  w.mel_lo = take((size_t)w.cap_rowsM * kMelOp * 2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) w.fx[i][j] = take((size_t)w.cap_rowsM * kF0Ch * 2);
  for (int l = 0; l < NL; ++l) {
    const long long cap = (long long)cfg.level_mul[l] * total_T + (long long)B * (kGap + 1) + kGap + 512;
    w.cap_rows[l] = cap;
    const int C = kBase >> (l + 1);
    w.U[l] = (float*)take((size_t)cap * C * 4);
    w.S[l] = (float*)take((size_t)cap * C * 4);
    w.X[l] = (float*)take((size_t)cap * C * 4);
    w.XR[l] = (float*)take((size_t)cap * C * 4);
    w.Y[l] = (float*)take((size_t)cap * C * 4);
    w.S2[l] = (float*)take((size_t)cap * C * 4);
    w.XR2[l] = (float*)take((size_t)cap * C * 4);
    w.XR3[l] = (float*)take((size_t)cap * C * 4);
    w.XR4[l] = (float*)take((size_t)cap * C * 4);
    for (int i = 0; i < 4; ++i) w.A[l][i] = take((size_t)cap * C * es);
    w.Yact[l] = take((size_t)cap * C * es);
  }
  w.spec = (float*)take((size_t)w.cap_rows[NL - 1] * kSpecCh * 4);
  w.post = (float*)take((size_t)w.cap_rows[NL - 1] * kSpecCh * 4);
  w.spec_op = take((size_t)w.cap_rows[NL - 1] * kSpecOp * 2);
  w.bytes = off;
  return w;
}

// Zero the gap rows of every buffer the tensor-core kernels read with a halo: ONE launch over a table.
struct GapBuf { void* p; int row_bytes; int level; };      // level 0..2, 3 = gapped mel rate
constexpr int kMaxGapBufs = 40;
struct GapTable {
  GapBuf b[kMaxGapBufs];
  const long long* off[4];
  int mul[4], plus[4];
};
__global__ void k_zero_gaps(const GapTable t, const int* T) {
  // gap g (0..B) of buffer blockIdx.y: rows [start_g, start_g + kGap)
  const GapBuf gb = t.b[blockIdx.y];
  const int g = blockIdx.x, l = gb.level;
  long long start = 0;
  if (g > 0) start = t.off[l][g - 1] + (long long)t.mul[l] * T[g - 1] + t.plus[l];
  const long long bytes = (long long)kGap * gb.row_bytes;
  uint4* p = reinterpret_cast<uint4*>((char*)gb.p + start * gb.row_bytes);
  for (long long i = threadIdx.x; i < bytes / 16; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
}

}  // namespace

// Profiling timeline: record a named event after the work enqueued so far.
static int mark(vt_hift* h, const char* name, cudaStream_t st) {
  if (!h->profiling) return VT_OK;
  if (h->n_marks == h->marks.size()) {
    cudaEvent_t e;
    VT_CUDA_OK(cudaEventCreate(&e));
    h->marks.emplace_back(name, e);
  }
  h->marks[h->n_marks].first = name;
  VT_CUDA_OK(cudaEventRecord(h->marks[h->n_marks].second, st));
  ++h->n_marks;
  return VT_OK;
}

static ConvArgs base_args(const ConvLayer& L, const Plan& P, const Plan::Seg& seg) {
  ConvArgs a{};
  a.w = L.w; a.bias = L.bias; a.wscale = L.wscale;
  a.cin = L.cin; a.cout = L.cout; a.k = L.k; a.dil = L.dil; a.stride = L.stride; a.pad = L.pad;
  a.in_ld = L.cin;
  a.pro_act = ACT_NONE; a.pro_slope = 0.f;
  a.out_scale = 1.0f;
  a.out_mul = L.out_mul; a.out_shift = 0; a.phase_c = L.phase_c; a.dup_row2 = 0;
  a.tiles = P.d_tiles + seg.off;
  a.n_tiles = seg.n;
  return a;
}

static int run_conv(vt_hift* h, const ConvArgs& a, const ConvLayer& L, int level, cudaStream_t st) {
  if (h->use_tc && L.w_tc && level >= 0 && convT_tc_supported(L)) {
    // C = 256: transposed formulation on 256-step tiles (weights streamed once per 256 steps)
    const Plan::Seg& seg = h->plan.tc[level][1];
    return launch_convT_tc(a, L, h->act_elem, h->plan.d_tiles + seg.off, seg.n, st);
  }
  if (h->use_tc && L.w_tc && level >= 0) {
    const int rows = conv_tc_tile_rows(L);
    const Plan::Seg& seg = h->plan.tc[level][rows == 256 ? 1 : 0];
    return launch_conv_tc(a, L, h->act_elem, h->plan.d_tiles + seg.off, seg.n, rows, st);
  }
  return launch_conv_ref(a, h->act_elem, st);
}

}  // namespace vt

namespace vt {
template <typename T> __device__ float to_f(T v);
template <> __device__ float to_f<float>(float v) { return v; }
template <> __device__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__global__ void k_tap(const T* src, long long row0, int ld, int ch, long long rows, float* dst) {
  const long long n = rows * ch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ch;
    const int c = (int)(i - r * ch);
    dst[i] = to_f<T>(src[(row0 + r) * ld + c]);
  }
}
}  // namespace

extern "C" {

int vt_hift_create(const vt_tensor* tensors, int n_tensors, int operand_dtype, vt_hift** out_handle) {
  return vt_hift_create_ex(tensors, n_tensors, operand_dtype, nullptr, out_handle);
}

int vt_hift_samples_per_frame(const vt_hift* h) { return h ? h->cfg.spf : VT_ERR_INVALID; }
int vt_hift_sampling_rate(const vt_hift* h) { return h ? h->cfg.sr : VT_ERR_INVALID; }

int vt_hift_create_ex(const vt_tensor* tensors, int n_tensors, int operand_dtype, const vt_hift_config* user_cfg,
                      vt_hift** out_handle) {
  VT_REQUIRE(tensors && n_tensors > 0 && out_handle, "vt_hift_create: NULL argument");
  HiftCfg cfg;                                   // defaults: Chatterbox S3Gen
  if (user_cfg) {
    VT_REQUIRE(user_cfg->n_upsamples == 2 || user_cfg->n_upsamples == 3, "vt_hift_create: 2 or 3 upsampling stages are supported (got %d)",
               user_cfg->n_upsamples);
    VT_REQUIRE(user_cfg->sampling_rate >= 8000 && user_cfg->sampling_rate <= 96000, "vt_hift_create: bad sampling rate %d", user_cfg->sampling_rate);
    cfg.n_levels = user_cfg->n_upsamples;
    cfg.sr = user_cfg->sampling_rate;
    cfg.trim_fade = user_cfg->trim_fade ? 1 : 0;
    int mul = 1;
    for (int i = 0; i < cfg.n_levels; ++i) {
      const int u = user_cfg->upsample_rates[i], k = user_cfg->upsample_kernel_sizes[i], sk = user_cfg->source_resblock_kernel_sizes[i];
      VT_REQUIRE(u >= 1 && u <= 16 && k >= u && ((k - u) & 1) == 0, "vt_hift_create: unsupported upsampling stage %d (rate %d, kernel %d)", i, u, k);
      VT_REQUIRE(sk == 3 || sk == 7 || sk == 11, "vt_hift_create: source ResBlock kernel size must be 3, 7 or 11 (got %d)", sk);
      cfg.up_rate[i] = u; cfg.up_kernel[i] = k; cfg.src_rb_kernel[i] = sk;
      mul *= u;
      cfg.level_mul[i] = mul;
    }
    // source_downs (upstream HiFTGenerator.__init__): stride = product of the LATER upsampling rates; 1 -> k = 1,
    // else k = 2 u, padding u / 2
    for (int i = 0; i < cfg.n_levels; ++i) {
      int u = 1;
      for (int j = i + 1; j < cfg.n_levels; ++j) u *= cfg.up_rate[j];
      cfg.sd_s[i] = u; cfg.sd_k[i] = u == 1 ? 1 : 2 * u; cfg.sd_p[i] = u == 1 ? 0 : u / 2;
    }
    cfg.spf = mul * kHop;
    cfg.trim_n = cfg.sr / 50;
  }
  const int NL = cfg.n_levels;
  VT_REQUIRE(operand_dtype == VT_OPERAND_FP16 || operand_dtype == VT_OPERAND_BF16 || operand_dtype == VT_OPERAND_FP32,
             "vt_hift_create: unknown operand dtype %d", operand_dtype);
  std::map<std::string, HostTensor> tab;
  for (int i = 0; i < n_tensors; ++i) {
    VT_REQUIRE(tensors[i].name && tensors[i].data && tensors[i].ndim >= 1 && tensors[i].ndim <= 4,
               "vt_hift_create: bad tensor entry %d", i);
    HostTensor t;
    t.data = tensors[i].data;
    t.shape.assign(tensors[i].shape, tensors[i].shape + tensors[i].ndim);
    tab[tensors[i].name] = t;
  }
  vt_hift* h = new vt_hift();
  h->cfg = cfg;
  h->act_elem = operand_dtype == VT_OPERAND_FP32 ? ELEM_F32 : (operand_dtype == VT_OPERAND_FP16 ? ELEM_F16 : ELEM_BF16);
  h->use_tc = operand_dtype != VT_OPERAND_FP32;
  int rc = VT_OK;
  auto fail = [&](int code) { vt_hift_destroy(h); return code; };
#define TRY(expr) do { rc = (expr); if (rc) return fail(rc); } while (0)
  TRY(pack_conv(h, h->conv_pre, tab, "conv_pre", kMel, kBase, 7, 1, 1, 3, kMel, kBase, GEMM_SPLIT, kMelOp));
  for (int i = 0; i < NL; ++i) {
    const int cin = kBase >> i, cout = kBase >> (i + 1);
    TRY(pack_convT(h, h->ups[i], tab, "ups." + std::to_string(i), cin, cout, cfg.up_kernel[i], cfg.up_rate[i],
                   (cfg.up_kernel[i] - cfg.up_rate[i]) / 2));
    TRY(pack_conv(h, h->sdown[i], tab, "source_downs." + std::to_string(i), kNfft + 2, cout, cfg.sd_k[i], 1, cfg.sd_s[i],
                  cfg.sd_p[i], kSpecCh, cout, GEMM_IM2COL, kSpecOp));
    for (int j = 0; j < 3; ++j) {
      const std::string p = "source_resblocks." + std::to_string(i);
      const int k = cfg.src_rb_kernel[i];
      TRY(pack_conv(h, h->src_c1[i][j], tab, p + ".convs1." + std::to_string(j), cout, cout, k, kRbDil[j], 1,
                    (k * kRbDil[j] - kRbDil[j]) / 2, cout, cout));
      TRY(pack_conv(h, h->src_c2[i][j], tab, p + ".convs2." + std::to_string(j), cout, cout, k, 1, 1, (k - 1) / 2, cout, cout));
      TRY(upload_vec(h, tab, p + ".activations1." + std::to_string(j) + ".alpha", cout, &h->src_a1[i][j]));
      TRY(upload_vec(h, tab, p + ".activations2." + std::to_string(j) + ".alpha", cout, &h->src_a2[i][j]));
    }
    for (int kk = 0; kk < 3; ++kk) {
      const int r = i * 3 + kk, k = kRbKernels[kk];
      const std::string p = "resblocks." + std::to_string(r);
      for (int j = 0; j < 3; ++j) {
        TRY(pack_conv(h, h->rb_c1[r][j], tab, p + ".convs1." + std::to_string(j), cout, cout, k, kRbDil[j], 1,
                      (k * kRbDil[j] - kRbDil[j]) / 2, cout, cout));
        TRY(pack_conv(h, h->rb_c2[r][j], tab, p + ".convs2." + std::to_string(j), cout, cout, k, 1, 1, (k - 1) / 2, cout, cout));
        TRY(upload_vec(h, tab, p + ".activations1." + std::to_string(j) + ".alpha", cout, &h->rb_a1[r][j]));
        TRY(upload_vec(h, tab, p + ".activations2." + std::to_string(j) + ".alpha", cout, &h->rb_a2[r][j]));
      }
    }
  }
  for (int i = 0; i < NL && h->use_tc; ++i) {
    const char* nf = getenv("VT_NO_FUSE");
    bool ok = !(nf && nf[0] == '1');
    for (int j = 0; j < 3 && ok; ++j) {
      ok = ok && pair_tc_supported(h->src_c1[i][j], h->src_c2[i][j]);
      for (int kk = 0; kk < 3; ++kk) ok = ok && pair_tc_supported(h->rb_c1[i * 3 + kk][j], h->rb_c2[i * 3 + kk][j]);
    }
    // C = 128 can also run unfused on the transposed single-conv kernel (VT_FUSE_L1=0).  Measured: 8.6 ms against
    // 7.8 ms fused for the level - unfused, a pair moves 2 KB of HBM per step and the k = 3 / 7 convs fall under the
    // ridge again.
    const char* fl1 = getenv("VT_FUSE_L1");
    if (ok && (kBase >> (i + 1)) == 128 && fl1 && fl1[0] == '0' && convT_tc_supported(h->rb_c1[i * 3][0])) ok = false;
    h->fuse[i] = ok;
    if (ok && (kBase >> (i + 1)) == 64) {
      // Tap-paired kernel per kernel size.  Measured per launch (ncu, us; plain pairs, tap-paired / activation-major):
      // k = 11: 860-890 / 1010, k = 7: 760-785 / 770-777, k = 3: 717-721 / 593-594 - the paired MMAs halve the
      // tensor-pipe time, but with it gone the tile period is bound by the CUDA-core work (two Snakes and two epilogues
      // per step), so only the MMA-bound k = 11 pairs gain.  VT_PAIR64=0: never, VT_PAIR64=7: k = 7 and 11, VT_PAIR64=all: every kernel size and
      // also the last pair of each ResBlock.
      const char* e64 = getenv("VT_PAIR64");
      h->pair64_last = e64 && e64[0] == 'a';
      for (int kk = 0; kk < 3; ++kk) {
        bool p64 = e64 ? (e64[0] == 'a' || (e64[0] == '7' && kRbKernels[kk] >= 7)) : kRbKernels[kk] == 11;
        for (int j = 0; j < 3; ++j) {
          p64 = p64 && pair64_tc_supported(h->rb_c1[i * 3 + kk][j], h->rb_c2[i * 3 + kk][j]);
          if (cfg.src_rb_kernel[i] == kRbKernels[kk]) p64 = p64 && pair64_tc_supported(h->src_c1[i][j], h->src_c2[i][j]);
        }
        h->pair64[kk] = p64;
      }
    }
  }
  TRY(pack_conv(h, h->conv_post, tab, "conv_post", kBase >> NL, kNfft + 2, 7, 1, 1, 3, kBase >> NL, kSpecCh));
  for (int i = 0; i < 5; ++i)
    TRY(pack_conv(h, h->f0c[i], tab, "f0_predictor.condnet." + std::to_string(2 * i), i == 0 ? kMel : kF0Ch, kF0Ch, 3, 1, 1, 1,
                  i == 0 ? kMel : kF0Ch, kF0Ch, GEMM_SPLIT, i == 0 ? kMelOp : kF0Ch));
  TRY(upload_vec(h, tab, "f0_predictor.classifier.weight", kF0Ch, &h->f0_w));
  TRY(upload_vec(h, tab, "f0_predictor.classifier.bias", 1, &h->f0_b));
  TRY(upload_vec(h, tab, "m_source.l_linear.weight", kHarm, &h->lin_w));
  TRY(upload_vec(h, tab, "m_source.l_linear.bias", 1, &h->lin_b));
  {
    // s3gen.py: trim_fade = zeros(2n); trim_fade[n:] = (cos(linspace(pi, 0, n)) + 1) / 2, n = S3GEN_SR // 50 = 480 (fp32)
    const int tn = cfg.trim_n;
    std::vector<float> tf(2 * tn, 0.0f);
    const float step = (0.0f - 3.14159265358979323846f) / (float)(tn - 1);
    for (int i = 0; i < tn; ++i) {
      // torch.linspace (fp32): first half from the start, second half from the end
      const float x = i < tn / 2 ? 3.14159265358979323846f + step * (float)i : 0.0f - step * (float)(tn - 1 - i);
      tf[tn + i] = (cosf(x) + 1.0f) / 2.0f;
    }
    TRY(dev_upload(h, tf.data(), tf.size() * 4, (void**)&h->trim_fade));
  }
#undef TRY
  *out_handle = h;
  return VT_OK;
}

void vt_hift_destroy(vt_hift* h) {
  if (!h) return;
  for (void* p : h->allocs) cudaFree(p);
  free_plan(h->plan);
  for (auto& c : h->plan_cache) free_plan(c);
  for (int i = 0; i < 2; ++i) if (h->ev_fwd[i]) cudaEventDestroy(h->ev_fwd[i]);
  for (int l = 0; l < 3; ++l)
    for (int i = 0; i < 2; ++i) if (h->ev_rb[l][i]) cudaEventDestroy(h->ev_rb[l][i]);
  for (auto& m : h->marks) cudaEventDestroy(m.second);
  delete h;
}

int64_t vt_hift_workspace_bytes(const vt_hift* h, int B, int64_t total_T, int64_t T_max) {
  if (!h || B < 0 || total_T < 0 || T_max < 0) return VT_ERR_INVALID;
  return (int64_t)carve_ws(h, B, total_T, nullptr).bytes;
}

int vt_hift_forward(vt_hift* h, const float* mel, const int32_t* T, int B, const float* f0_in, const float* phase_vec,
                    const float* noise, uint64_t seed, float* wav, void* workspace, int64_t workspace_bytes,
                    void* stream_v) {
  VT_REQUIRE(h != nullptr, "vt_hift_forward: NULL handle");
  VT_REQUIRE(B >= 0 && B <= 65535, "vt_hift_forward: B must be in [0, 65535]");
  launch_counter() = 0;
  if (B == 0) return VT_OK;
  VT_REQUIRE(mel && T && wav && workspace, "vt_hift_forward: NULL argument");
  VT_REQUIRE((reinterpret_cast<uintptr_t>(mel) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
             "vt_hift_forward: mel must be 16-byte and workspace 256-byte aligned");
  long long total_T = 0;
  for (int b = 0; b < B; ++b) {
    VT_REQUIRE(T[b] >= 1, "vt_hift_forward: every sequence needs at least one mel frame (T[%d]=%d)", b, T[b]);
    total_T += T[b];
  }
  const HiftCfg& cfg = h->cfg;
  const int NL = cfg.n_levels, LL = NL - 1;
  VT_REQUIRE(total_T * (long long)(cfg.level_mul[LL] + 1) + B * 40LL < 2000000000LL, "vt_hift_forward: batch too large for 32-bit tile tables");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = build_plan(h, T, B, st);
  if (rc) return rc;
  const Plan& P = h->plan;
  Workspace w = carve_ws(h, B, total_T, workspace);
  VT_REQUIRE((int64_t)w.bytes <= workspace_bytes, "vt_hift_forward: workspace too small (%lld < %lld)",
             (long long)workspace_bytes, (long long)w.bytes);
  h->ws = w;
  h->have_forward = true;
  const int ae = h->act_elem;
  const bool prof = h->profiling;
  if (prof) {
    VT_CUDA_OK(cudaEventRecord(h->ev_fwd[0], st));
    h->rb_flops = 0;
    h->rb_launches = 0;
    h->n_marks = 0;
    mark(h, "start", st);
  }

  if (h->use_tc) {
    GapTable gt{};
    int nb = 0;
    auto add = [&](void* p, int row_bytes, int level) { gt.b[nb++] = GapBuf{p, row_bytes, level}; };
    for (int l = 0; l < NL; ++l) {
      const int C = kBase >> (l + 1);
      gt.off[l] = P.d_off[l]; gt.mul[l] = cfg.level_mul[l]; gt.plus[l] = l == LL ? 1 : 0;
      add(w.Yact[l], C * (int)elem_size(ae), l);
      if (h->fuse[l]) {
        // fused pairs read the fp32 streams with their halo
        float* fb[7] = {w.S[l], w.S2[l], w.X[l], w.XR[l], w.XR2[l], w.XR3[l], w.XR4[l]};
        for (int i = 0; i < 7; ++i) add(fb[i], C * 4, l);
      } else {
        for (int i = 0; i < 4; ++i) add(w.A[l][i], C * (int)elem_size(ae), l);
      }
    }
    gt.off[3] = P.d_offM; gt.mul[3] = 1; gt.plus[3] = 0;
    add(w.xpre_act, kBase * (int)elem_size(ae), 3);
    // operands of the K-blocked layers: mel split, F0 trunk ping-pong, STFT rows
    add(w.mel_hi, kMelOp * 2, 3);
    add(w.mel_lo, kMelOp * 2, 3);
    if (!f0_in)
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) add(w.fx[i][j], kF0Ch * 2, 3);
    add(w.spec_op, kSpecOp * 2, LL);
    VT_REQUIRE(nb <= kMaxGapBufs, "gap table overflow");
    k_zero_gaps<<<dim3(B + 1, nb), 256, 0, st>>>(gt, P.d_T);
    VT_LAUNCHED();
    rc = launch_pack_mel(mel, P.d_mel_off, P.d_T, P.d_offM, B, total_T, w.mel_hi, w.mel_lo, st);
    if (rc) return rc;
  }

  mark(h, "zero_gaps", st);
  // ---- F0 (ConvRNNF0Predictor) unless injected
  const float* f0 = f0_in;
  if (!f0) {
    float* bufs[2] = {w.f0a, w.f0b};
    for (int i = 0; i < 5 && h->use_tc; ++i) {
      // tensor-core trunk on two-term fp16 operands (x = hi + lo): ~fp32 accuracy for the value that
      // feeds the phase integral; the last layer writes fp32 ELU rows (ungapped) for the classifier head
      ConvArgs a = base_args(h->f0c[i], P, i < 4 ? P.g_mel : P.g_melu);
      a.in_act = i == 0 ? w.mel_hi : w.fx[(i - 1) & 1][0];
      a.in_act2 = i == 0 ? w.mel_lo : w.fx[(i - 1) & 1][1];
      a.in_ld = i == 0 ? kMelOp : kF0Ch;
      if (i < 4) {
        a.act[0] = {w.fx[i & 1][0], nullptr, ACT_ELU, 0.f};
        a.act[1] = {w.fx[i & 1][1], nullptr, ACT_ELU, 0.f};
      } else {
        a.out = bufs[0];
        a.act[0] = {nullptr, nullptr, ACT_ELU, 0.f};
      }
      rc = launch_gemm_tc(a, h->f0c[i], ae, st);
      if (rc) return rc;
    }
    for (int i = 0; i < 5 && !h->use_tc; ++i) {
      ConvArgs a = base_args(h->f0c[i], P, P.mel);
      a.in = i == 0 ? mel : bufs[(i - 1) & 1];
      // the F0 trunk is fp32 end to end: its ELU output is the next conv's fp32 input
      a.out = nullptr;
      a.act[0] = {bufs[i & 1], nullptr, ACT_ELU, 0.f};
      rc = launch_conv_ref(a, ELEM_F32, st);
      if (rc) return rc;
    }
    rc = launch_f0_head(bufs[0], h->f0_w, h->f0_b, w.f0, total_T, st);
    if (rc) return rc;
    f0 = w.f0;
  } else {
    VT_CUDA_OK(cudaMemcpyAsync(w.f0, f0_in, (size_t)total_T * 4, cudaMemcpyDeviceToDevice, st));  // keeps the "f0" tap valid
  }
  mark(h, "f0_predictor", st);
  // ---- source: SineGen -> tanh(Linear) -> STFT
  rc = launch_sine_source(f0, P.d_mel_off, P.d_T, B, total_T, phase_vec, noise, seed, h->lin_w, h->lin_b,
                          w.phase_base, w.s, cfg.spf, cfg.sr, st);
  if (rc) return rc;
  rc = launch_stft(w.s, P.d_mel_off, P.d_T, P.d_off[LL], B, total_T, h->use_tc ? nullptr : w.spec, h->use_tc ? w.spec_op : nullptr, ae,
                   cfg.spf, st);
  if (rc) return rc;
  mark(h, "source_stft", st);
  // ---- conv_pre
  {
    ConvArgs a = base_args(h->conv_pre, P, h->use_tc ? P.g_mel : P.pre);
    a.out = w.xpre;
    if (h->use_tc) {
      a.in_act = w.mel_hi; a.in_act2 = w.mel_lo; a.in_ld = kMelOp;
      a.act[0] = {w.xpre_act, nullptr, ACT_LRELU, 0.1f};
      a.act_from_out = 1;
      rc = launch_gemm_tc(a, h->conv_pre, ae, st);
    } else {
      a.in = mel;
      rc = launch_conv_ref(a, ae, st);
    }
    if (rc) return rc;
  }
  mark(h, "conv_pre", st);
  for (int i = 0; i < NL; ++i) {
    const std::string sfx = std::to_string(i);
    // ups[i]( leaky_relu(x, 0.1) ), reflection pad (1, 0) on the last stage
    {
      ConvArgs a = base_args(h->ups[i], P, P.up[i]);
      a.out = w.U[i];
      if (i == LL) { a.out_shift = 1; a.dup_row2 = 1; }
      if (h->use_tc && h->ups[i].w_tc) {
        a.in_act = i == 0 ? w.xpre_act : w.Yact[i - 1];   // leaky_relu already applied by the producer
        rc = launch_conv_tc(a, h->ups[i], ae, P.d_tiles + P.tcu[i].off, P.tcu[i].n, 128, st);
      } else {
        a.in = i == 0 ? w.xpre : w.Y[i - 1];
        a.pro_act = ACT_LRELU; a.pro_slope = 0.1f;
        rc = launch_conv_ref(a, ae, st);
      }
      if (rc) return rc;
    }
    mark(h, ("ups" + sfx).c_str(), st);
    // source_downs[i](s_stft) -> S stream + Snake copy for the first source-resblock conv
    {
      ConvArgs a = base_args(h->sdown[i], P, h->use_tc ? P.g_sd[i] : P.sd[i]);
      a.out = w.S[i];
      if (!h->fuse[i]) a.act[0] = {w.A[i][1], h->src_a1[i][0], ACT_SNAKE, 0.f};
      if (h->use_tc) {
        a.in_act = w.spec_op; a.in_ld = kSpecOp;
        rc = launch_gemm_tc(a, h->sdown[i], ae, st);
      } else {
        a.in = w.spec;
        rc = launch_conv_ref(a, ae, st);
      }
      if (rc) return rc;
    }
    mark(h, ("source_down" + sfx).c_str(), st);
    // source_resblocks[i]; its last conv also adds the upsampled stream: x = ups + si
    if (prof) {
      VT_CUDA_OK(cudaEventRecord(h->ev_rb[i][0], st));
      const double steps = (double)cfg.level_mul[i] * (double)total_T + (i == LL ? B : 0);
      const double c = (double)(kBase >> (i + 1));
      h->rb_flops += steps * 2.0 * c * c * 6.0 * (cfg.src_rb_kernel[i] + kRbKernels[0] + kRbKernels[1] + kRbKernels[2]);
      h->rb_launches += h->fuse[i] ? 12 : 24;
    }
    if (h->fuse[i]) {
      // fused pairs: fp32 stream in, fp32 stream out, no operand copies in HBM
      float* sbuf[2] = {w.S[i], w.S2[i]};
      const int ksrc = cfg.src_rb_kernel[i] == 3 ? 0 : (cfg.src_rb_kernel[i] == 7 ? 1 : 2);
      const bool c64 = (kBase >> (i + 1)) == 64;          // the tap-paired kernel is the C = 64 kernel
      for (int j = 0; j < 3; ++j) {
        // the tap-paired kernel takes the plain pairs; the last pair of a ResBlock (second residual / running mean /
        // operand copy: more streams in the fin epilogue) stays on the activation-major kernel - measured faster there
        const bool p64 = c64 && h->pair64[ksrc] && (j < 2 || h->pair64_last);
        ConvArgs c = base_args(h->src_c2[i][j], P, p64 ? P.pair64[ksrc] : P.pair[i][ksrc]);
        c.res1 = sbuf[j & 1];
        if (j < 2) c.out = sbuf[(j + 1) & 1];
        else { c.res2 = w.U[i]; c.out = w.X[i]; }
        rc = p64 ? launch_pair64_tc(c, h->src_c1[i][j], h->src_c2[i][j], h->src_a1[i][j], h->src_a2[i][j], ae, st)
                                   : launch_pair_tc(c, h->src_c1[i][j], h->src_c2[i][j], h->src_a1[i][j], h->src_a2[i][j], ae, st);
        if (rc) return rc;
      }
      mark(h, ("source_resblock" + sfx).c_str(), st);
      // the last pairs of the three ResBlocks as ONE mean-fused launch when their weights allow it (no scaled rows)
      const ConvLayer* l1[3] = {&h->rb_c1[i * 3][2], &h->rb_c1[i * 3 + 1][2], &h->rb_c1[i * 3 + 2][2]};
      const ConvLayer* l2[3] = {&h->rb_c2[i * 3][2], &h->rb_c2[i * 3 + 1][2], &h->rb_c2[i * 3 + 2][2]};
      const bool tri = pair3_tc_supported(l1, l2, ae);
      float* ylast[3] = {w.XR2[i], w.XR3[i], w.XR4[i]};
      for (int r = 0; r < 3; ++r) {
        const int R = i * 3 + r;
        float* xin[3] = {w.X[i], w.XR[i], tri ? ylast[r] : w.XR2[i]};
        for (int j = 0; j < (tri ? 2 : 3); ++j) {
          const bool p64 = c64 && h->pair64[r] && (j < 2 || h->pair64_last);
          ConvArgs c = base_args(h->rb_c2[R][j], P, p64 ? P.pair64[r] : P.pair[i][r]);
          c.res1 = xin[j];
          if (j < 2) c.out = xin[j + 1];
          else {
            c.out = w.Y[i];
            c.out_accum = r > 0;
            c.out_scale = 1.0f / 3.0f;
            if (r == 2) {
              c.act[0] = {w.Yact[i], nullptr, ACT_LRELU, i < LL ? 0.1f : 0.01f};
              c.act_from_out = 1;
            }
          }
          rc = p64 ? launch_pair64_tc(c, h->rb_c1[R][j], h->rb_c2[R][j], h->rb_a1[R][j], h->rb_a2[R][j], ae, st)
                                     : launch_pair_tc(c, h->rb_c1[R][j], h->rb_c2[R][j], h->rb_a1[R][j], h->rb_a2[R][j], ae, st);
          if (rc) return rc;
        }
      }
      if (tri) {
        if (prof) h->rb_launches -= 2;            // three last pairs in one launch
        ConvArgs c = base_args(h->rb_c2[i * 3 + 2][2], P, P.pair[i][2]);     // tiles of the largest kernel size
        c.out = w.Y[i];
        c.out_scale = 1.0f / 3.0f;
        c.act[0] = {w.Yact[i], nullptr, ACT_LRELU, i < LL ? 0.1f : 0.01f};
        c.act_from_out = 1;
        const float* a1[3] = {h->rb_a1[i * 3][2], h->rb_a1[i * 3 + 1][2], h->rb_a1[i * 3 + 2][2]};
        const float* a2[3] = {h->rb_a2[i * 3][2], h->rb_a2[i * 3 + 1][2], h->rb_a2[i * 3 + 2][2]};
        // Sub order.  At C = 128 the tile's single D2 buffer is drained by the fin pass while the NEXT tile's first sub
        // already runs conv1 and mid; its conv2 then waits for the fin pass to end.  The largest kernel first gives the fin
        // pass (three residual streams) the longest conv1 to hide behind (VT_PAIR3_ORDER=0: kernel sizes ascending).
        static const bool asc = getenv("VT_PAIR3_ORDER") && getenv("VT_PAIR3_ORDER")[0] == '0';
        if ((kBase >> (i + 1)) == 128 && !asc) {
          std::swap(l1[0], l1[2]); std::swap(l2[0], l2[2]); std::swap(a1[0], a1[2]); std::swap(a2[0], a2[2]);
          std::swap(ylast[0], ylast[2]);
        }
        rc = launch_pair3_tc(c, l1, l2, a1, a2, ylast, ae, st);
        if (rc) return rc;
      }
    }
    for (int j = 0; j < 3 && !h->fuse[i]; ++j) {
      ConvArgs a = base_args(h->src_c1[i][j], P, P.lvl[i]);
      a.in_act = w.A[i][1];
      a.act[0] = {w.A[i][3], h->src_a2[i][j], ACT_SNAKE, 0.f};
      rc = run_conv(h, a, h->src_c1[i][j], i, st);
      if (rc) return rc;
      ConvArgs c = base_args(h->src_c2[i][j], P, P.lvl[i]);
      c.in_act = w.A[i][3];
      c.res1 = w.S[i];
      if (j < 2) {
        c.out = w.S[i];
        c.act[0] = {w.A[i][1], h->src_a1[i][j + 1], ACT_SNAKE, 0.f};
      } else {
        c.res2 = w.U[i];
        c.out = w.X[i];
        for (int r = 0; r < 3; ++r) c.act[r] = {w.A[i][r], h->rb_a1[i * 3 + r][0], ACT_SNAKE, 0.f};
      }
      rc = run_conv(h, c, h->src_c2[i][j], i, st);
      if (rc) return rc;
    }
    if (!h->fuse[i]) mark(h, ("source_resblock" + sfx).c_str(), st);
    // three multi-receptive-field resblocks, averaged
    for (int r = 0; r < 3 && !h->fuse[i]; ++r) {
      const int R = i * 3 + r;
      for (int j = 0; j < 3; ++j) {
        ConvArgs a = base_args(h->rb_c1[R][j], P, P.lvl[i]);
        a.in_act = w.A[i][r];
        a.act[0] = {w.A[i][3], h->rb_a2[R][j], ACT_SNAKE, 0.f};
        rc = run_conv(h, a, h->rb_c1[R][j], i, st);
        if (rc) return rc;
        ConvArgs c = base_args(h->rb_c2[R][j], P, P.lvl[i]);
        c.in_act = w.A[i][3];
        c.res1 = j == 0 ? w.X[i] : w.XR[i];
        if (j < 2) {
          c.out = w.XR[i];
          c.act[0] = {w.A[i][r], h->rb_a1[R][j + 1], ACT_SNAKE, 0.f};
        } else {
          c.out = w.Y[i];
          c.out_accum = r > 0;
          c.out_scale = 1.0f / 3.0f;
          if (h->use_tc && r == 2) {   // the mean is complete: emit the next layer's leaky_relu operand copy
            c.act[0] = {w.Yact[i], nullptr, ACT_LRELU, i < LL ? 0.1f : 0.01f};
            c.act_from_out = 1;
          }
        }
        rc = run_conv(h, c, h->rb_c2[R][j], i, st);
        if (rc) return rc;
      }
    }
    if (prof) VT_CUDA_OK(cudaEventRecord(h->ev_rb[i][1], st));
    mark(h, ("resblocks" + sfx).c_str(), st);
  }
  // ---- conv_post( leaky_relu(x) ) with the default slope 0.01, then the spectral head
  {
    ConvArgs a = base_args(h->conv_post, P, P.lvl[LL]);
    a.out = w.post;
    if (h->use_tc && h->conv_post.w_tc) {
      a.in_act = w.Yact[LL];
      const int rows = conv_tc_tile_rows(h->conv_post);
      const Plan::Seg& seg = P.tc[LL][rows == 256 ? 1 : 0];
      rc = launch_conv_tc(a, h->conv_post, ae, P.d_tiles + seg.off, seg.n, rows, st);
    } else {
      a.in = w.Y[LL];
      a.pro_act = ACT_LRELU; a.pro_slope = 0.01f;
      rc = launch_conv_ref(a, ae, st);
    }
    if (rc) return rc;
  }
  mark(h, "conv_post", st);
  rc = launch_istft_head(w.post, P.d_mel_off, P.d_T, P.d_off[LL], B, P.T_max, h->trim_fade, cfg.trim_fade ? 2 * cfg.trim_n : 0, cfg.spf, h->use_tc, wav, st);
  if (rc) return rc;
  mark(h, "istft_head", st);
  if (prof) VT_CUDA_OK(cudaEventRecord(h->ev_fwd[1], st));
  VT_CUDA_OK(cudaEventRecord(h->plan.done, st));
  return VT_OK;
}

int vt_hift_set_profiling(vt_hift* h, int enable) {
  VT_REQUIRE(h != nullptr, "vt_hift_set_profiling: NULL handle");
  if (enable && !h->ev_fwd[0]) {
    for (int i = 0; i < 2; ++i) VT_CUDA_OK(cudaEventCreate(&h->ev_fwd[i]));
    for (int l = 0; l < 3; ++l)
      for (int i = 0; i < 2; ++i) VT_CUDA_OK(cudaEventCreate(&h->ev_rb[l][i]));
  }
  h->profiling = enable != 0;
  return VT_OK;
}

int vt_hift_read_profile(vt_hift* h, double* total_ms, double* resblock_ms, double* resblock_flops,
                         int* resblock_launches) {
  VT_REQUIRE(h != nullptr && h->profiling && h->have_forward, "vt_hift_read_profile: profiling is off or no forward ran");
  VT_CUDA_OK(cudaEventSynchronize(h->ev_fwd[1]));
  float ms = 0.f;
  VT_CUDA_OK(cudaEventElapsedTime(&ms, h->ev_fwd[0], h->ev_fwd[1]));
  if (total_ms) *total_ms = ms;
  double rb = 0;
  for (int l = 0; l < h->cfg.n_levels; ++l) {
    VT_CUDA_OK(cudaEventElapsedTime(&ms, h->ev_rb[l][0], h->ev_rb[l][1]));
    rb += ms;
  }
  if (resblock_ms) *resblock_ms = rb;
  if (resblock_flops) *resblock_flops = h->rb_flops;
  if (resblock_launches) *resblock_launches = h->rb_launches;
  return VT_OK;
}


int vt_hift_read_timeline(vt_hift* h, char* out, int capacity) {
  VT_REQUIRE(h != nullptr && out && capacity > 0, "vt_hift_read_timeline: NULL argument");
  VT_REQUIRE(h->profiling && h->have_forward && h->n_marks > 0, "vt_hift_read_timeline: profiling is off or no forward ran");
  VT_CUDA_OK(cudaEventSynchronize(h->marks[h->n_marks - 1].second));
  std::string txt;
  for (size_t i = 1; i < h->n_marks; ++i) {
    float ms = 0.f;
    VT_CUDA_OK(cudaEventElapsedTime(&ms, h->marks[i - 1].second, h->marks[i].second));
    char line[96];
    snprintf(line, sizeof line, "%s=%.4f\n", h->marks[i].first.c_str(), ms);
    txt += line;
  }
  VT_REQUIRE((int)txt.size() < capacity, "vt_hift_read_timeline: capacity %d too small (need %d)", capacity, (int)txt.size() + 1);
  std::memcpy(out, txt.c_str(), txt.size() + 1);
  return VT_OK;
}

int64_t vt_hift_read_tap(vt_hift* h, const char* tap, int seq, float* out, int64_t capacity, void* workspace,
                         void* stream_v) {
  if (!h || !tap || !h->have_forward) { set_error("vt_hift_read_tap: no forward has run on this handle"); return VT_ERR_INVALID; }
  (void)workspace;
  const Plan& P = h->plan;
  if (seq < 0 || seq >= P.B) { set_error("vt_hift_read_tap: bad sequence index %d", seq); return VT_ERR_INVALID; }
  const Workspace& w = h->ws;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  const std::string name(tap);
  const int T = P.T[seq];
  const void* src = nullptr;
  long long row0 = 0, rows = 0;
  int ld = 0, ch = 0, elem = ELEM_F32;
  const HiftCfg& cfg = h->cfg;
  const int LL = cfg.n_levels - 1;
  const long long last_rows = (long long)cfg.level_mul[LL] * T + 1;
  auto level = [&](int l, const void* p, int e) {
    if (l < 0 || l > LL) { src = nullptr; return; }
    src = p; row0 = P.h_off[l][seq]; rows = (long long)cfg.level_mul[l] * T + (l == LL ? 1 : 0);
    ld = ch = kBase >> (l + 1); elem = e;
  };
  if (name == "f0") { src = w.f0; row0 = P.h_mel_off[seq]; rows = T; ld = ch = 1; }
  else if (name == "s") { src = w.s; row0 = (long long)P.h_mel_off[seq] * cfg.spf; rows = (long long)T * cfg.spf; ld = ch = 1; }
  else if (name == "s_stft") {
    // tensor-core modes keep only the operand-typed rows (kSpecOp wide)
    src = h->use_tc ? w.spec_op : (const void*)w.spec; row0 = P.h_off[LL][seq]; rows = last_rows;
    ld = h->use_tc ? kSpecOp : kSpecCh; ch = kNfft + 2; elem = h->use_tc ? h->act_elem : (int)ELEM_F32;
  }
  else if (name == "conv_post") { src = w.post; row0 = P.h_off[LL][seq]; rows = last_rows; ld = kSpecCh; ch = kNfft + 2; }
  else if (name == "conv_pre") { src = w.xpre; row0 = P.h_offM[seq]; rows = T; ld = ch = kBase; }
  else {
    // per-level taps: ups<l>, x<l>, stage<l>, act0<l>
    const int lv = name.empty() ? -1 : name[name.size() - 1] - '0';
    const std::string stem = name.substr(0, name.size() ? name.size() - 1 : 0);
    if (lv < 0 || lv > LL || (stem != "ups" && stem != "x" && stem != "stage" && stem != "act0")) {
      set_error("vt_hift_read_tap: unknown tap '%s' (levels 0..%d)", tap, LL);
      return VT_ERR_INVALID;
    }
    if (stem == "ups") level(lv, w.U[lv], ELEM_F32);
    else if (stem == "x") level(lv, w.X[lv], ELEM_F32);
    else if (stem == "stage") level(lv, w.Y[lv], ELEM_F32);
    else level(lv, w.A[lv][0], h->act_elem);
  }
  const int64_t n = rows * ch;
  if (!out) return n;
  if (capacity < n) { set_error("vt_hift_read_tap: capacity %lld < %lld", (long long)capacity, (long long)n); return VT_ERR_INVALID; }
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  if (elem == ELEM_F32) k_tap<float><<<blocks, 256, 0, st>>>((const float*)src, row0, ld, ch, rows, out);
  else if (elem == ELEM_F16) k_tap<__half><<<blocks, 256, 0, st>>>((const __half*)src, row0, ld, ch, rows, out);
  else k_tap<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)src, row0, ld, ch, rows, out);
  if (cudaGetLastError() != cudaSuccess) { set_error("vt_hift_read_tap: launch failed"); return VT_ERR_CUDA; }
  return n;
}

}  // extern "C"
