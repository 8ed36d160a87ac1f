// Plan object + C ABI for hot path 1 (convolutional autoencoder).  See include/behavenet_b200.h.
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/behavenet_b200.h"
#include "cae_kernels.cuh"

thread_local char g_bn_err[512] = "";
std::atomic<long long> g_bn_launches{0};
static std::atomic<int> g_tc_mode{1};

extern "C" int bn_abi_version(void) { return BN_ABI_VERSION; }

bool bn_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("BN_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
extern "C" const char* bn_last_error(void) { return g_bn_err; }
extern "C" int64_t bn_launch_count(void) { return (int64_t)g_bn_launches.load(); }
extern "C" int bn_set_tensor_core_mode(int mode) {
  g_tc_mode.store(mode ? 1 : 0);
  return 0;
}
extern "C" int bn_get_tensor_core_mode(void) { return g_tc_mode.load(); }

struct bn_cae_plan {
  bn_cae_desc d;
  int nl;
  ConvGeom enc[BN_MAX_LAYERS], dec[BN_MAX_LAYERS];
  std::vector<TapClass> h_tables;
  std::vector<int> enc_f, enc_d, dec_f, dec_d;   // table start indices
  TapClass* d_tables;
  size_t packed_floats, off_heads;
  size_t enc_sz[BN_MAX_LAYERS + 1];   // per-frame floats: [0] input, [i+1] output of enc layer i
  size_t dec_sz[BN_MAX_LAYERS + 1];   // [0] h0, [i+1] output of dec layer i
  size_t max_act;
  int feat_c, feat_h, feat_w;
};

namespace {

size_t align64(size_t x) { return (x + 63) & ~(size_t)63; }

struct WsLayout {
  size_t enc_act[BN_MAX_LAYERS + 1];
  size_t dec_act[BN_MAX_LAYERS + 1];
  size_t dpre_last, zcopy, gA, gB, partial, partial_floats, total;
  // weight-gradient partial sums: one region per layer (the reductions of a backward call are deferred
  // and batched), encoder and decoder calls share the space; `partial` (above) is the separate scratch of
  // the split-K implicit GEMMs and of bn_cae_layer_op
  size_t wg_enc[BN_MAX_LAYERS], wg_dec[BN_MAX_LAYERS], wg_floats_enc[BN_MAX_LAYERS], wg_floats_dec[BN_MAX_LAYERS];
  // per-CTA column-sum rows of the gradient image that layer i's backward-data kernel stores (= the bias
  // gradient of layer i - 1), reduced in the same batched launch as the weight-gradient slices
  size_t cp_enc[BN_MAX_LAYERS], cp_dec[BN_MAX_LAYERS], cp_floats_enc[BN_MAX_LAYERS], cp_floats_dec[BN_MAX_LAYERS];
};

// rows of the per-CTA column-sum table an implicit-GEMM launch can produce: one per (class, 128-row tile) of
// the im2col kernels, or one per 16 x 8 block of the halo kernel
size_t colpart_rows(const bn_cae_plan* p, int table_idx, int nclasses, int n) {
  long long maxM = 0, maxH = 0, maxW = 0;
  for (int c = 0; c < nclasses; ++c) {
    const TapClass& k = p->h_tables[table_idx + c];
    maxM = std::max(maxM, (long long)k.Hm * k.Wm);
    maxH = std::max<long long>(maxH, k.Hm);
    maxW = std::max<long long>(maxW, k.Wm);
  }
  const long long im2col = (long long)nclasses * (((long long)n * maxM + 127) / 128);
  const long long halo = (long long)n * ((maxW + 7) / 8) * ((maxH + 15) / 16);
  return (size_t)std::max(im2col, halo);
}

WsLayout ws_layout(const bn_cae_plan* p, int n) {
  WsLayout w;
  size_t o = 0;
  w.enc_act[0] = 0;
  for (int i = 1; i <= p->nl; ++i) { w.enc_act[i] = o; o += align64((size_t)n * p->enc_sz[i]); }
  for (int i = 0; i <= p->nl; ++i) { w.dec_act[i] = o; o += align64((size_t)n * p->dec_sz[i]); }
  w.dpre_last = o; o += align64((size_t)n * p->dec_sz[p->nl]);
  w.zcopy = o; o += align64((size_t)n * p->d.n_latents);
  w.gA = o; o += align64((size_t)n * p->max_act);
  w.gB = o; o += align64((size_t)n * p->max_act);
  size_t pf = 0;
  for (int i = 0; i < p->nl; ++i) {
    pf = std::max(pf, bn_wgrad_partial_floats(p->enc[i], n));
    pf = std::max(pf, bn_wgrad_partial_floats(p->dec[i], n));
  }
  w.partial = o; w.partial_floats = pf; o += align64(pf);
  size_t oe = o, od = o;
  for (int i = 0; i < p->nl; ++i) {
    w.wg_floats_enc[i] = bn_wgrad_partial_floats(p->enc[i], n);
    w.wg_enc[i] = oe; oe += align64(w.wg_floats_enc[i]);
    w.wg_floats_dec[i] = bn_wgrad_partial_floats(p->dec[i], n);
    w.wg_dec[i] = od; od += align64(w.wg_floats_dec[i]);
  }
  o = std::max(oe, od);
  oe = od = o;
  for (int i = 0; i < p->nl; ++i) {
    // encoder layer i, backward-data (dgrad form) -> image with enc[i].Cb channels
    w.cp_floats_enc[i] = i > 0 ? colpart_rows(p, p->enc_d[i], p->enc[i].n_dgrad, n) * p->enc[i].Cb : 0;
    w.cp_enc[i] = oe; oe += align64(w.cp_floats_enc[i]);
    // decoder layer i, backward-data (fprop form) -> image with dec[i].Cs channels
    w.cp_floats_dec[i] = i > 0 ? colpart_rows(p, p->dec_f[i], 1, n) * p->dec[i].Cs : 0;
    w.cp_dec[i] = od; od += align64(w.cp_floats_dec[i]);
  }
  o = std::max(oe, od);
  w.total = o;
  return w;
}

void build_classes(ConvGeom& g, std::vector<TapClass>& tab, int& i_f, int& i_d) {
  // fprop-form class: every small-image pixel gathers k*k taps from the big image
  TapClass c;
  memset(&c, 0, sizeof(c));
  c.Hm = g.Hs; c.Wm = g.Ws; c.oy0 = 0; c.ox0 = 0; c.ntaps = g.k * g.k;
  for (int ky = 0; ky < g.k; ++ky)
    for (int kx = 0; kx < g.k; ++kx) {
      int t = ky * g.k + kx;
      c.dy[t] = (signed char)(ky - g.pt);
      c.dx[t] = (signed char)(kx - g.pl);
      c.wt[t] = (unsigned char)t;
    }
  i_f = (int)tab.size();
  tab.push_back(c);
  // dgrad-form classes: big-image pixels with the same residue mod stride share a tap list;
  // big[y] receives small[(y + pt - ky) / s] * W[ky] whenever the division is exact
  i_d = (int)tab.size();
  g.n_dgrad = 0; g.dgrad_maxM = 0; g.dgrad_maxtaps = 0;
  for (int y0 = 0; y0 < g.s; ++y0)
    for (int x0 = 0; x0 < g.s; ++x0) {
      if (y0 >= g.Hb || x0 >= g.Wb) continue;
      TapClass d;
      memset(&d, 0, sizeof(d));
      d.Hm = (g.Hb - y0 + g.s - 1) / g.s;
      d.Wm = (g.Wb - x0 + g.s - 1) / g.s;
      d.oy0 = y0; d.ox0 = x0;
      int nt = 0;
      for (int ky = (y0 + g.pt) % g.s; ky < g.k; ky += g.s)
        for (int kx = (x0 + g.pl) % g.s; kx < g.k; kx += g.s) {
          d.dy[nt] = (signed char)((y0 + g.pt - ky) / g.s);
          d.dx[nt] = (signed char)((x0 + g.pl - kx) / g.s);
          d.wt[nt] = (unsigned char)(ky * g.k + kx);
          ++nt;
        }
      d.ntaps = nt;
      tab.push_back(d);
      g.n_dgrad++;
      g.dgrad_maxM = std::max(g.dgrad_maxM, d.Hm * d.Wm);
      g.dgrad_maxtaps = std::max(g.dgrad_maxtaps, nt);
    }
}

// scratch of the current C-ABI call (the workspace's split-K partial region), used by the
// tensor-core split-K variant; set by every entry point after it lays out the workspace
thread_local float* t_split_buf = nullptr;
thread_local size_t t_split_floats = 0;

int run_igemm(const bn_cae_plan* p, const ImgView& in, const float* w, const float* wt, int wrow,
              const float* bias, float* out, int Ho, int Wo, int Co, const float* dact, int table_idx,
              int nclasses, int maxM, int gs, int os, int n, int act, cudaStream_t st,
              float* colsum = nullptr, int* colsum_fused = nullptr, float* colpart = nullptr,
              size_t colpart_floats = 0) {
  // colsum (optional, [Co], accumulated into): the column sums of `out`, i.e. the bias gradient of the
  // layer below when `out` is a gradient image; *colsum_fused says whether the kernel's epilogue
  // took care of it (otherwise the caller runs bn_launch_colsum over the stored image)
  const TapClass* dcls = p->d_tables + table_idx;
  if (colsum_fused) *colsum_fused = 0;
  if (g_tc_mode.load()) {
    int maxtaps = 0;
    for (int c = 0; c < nclasses; ++c) maxtaps = std::max(maxtaps, p->h_tables[table_idx + c].ntaps);
    int r = bn_launch_igemm_tc(in, wt, wrow, bias, out, Ho, Wo, Co, dact, dcls, p->h_tables.data() + table_idx,
                               nclasses, maxM, maxtaps, gs, os, n, act, t_split_buf, t_split_floats, colsum,
                               colsum_fused, colpart, colpart_floats, st);
    if (r <= 0) return r;
    if (colsum_fused) *colsum_fused = 0;
  }
  return bn_launch_igemm(in, w, bias, out, Ho, Wo, Co, dact, dcls, nclasses, maxM, gs, os, n, act, st);
}

int run_wgrad(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
              size_t partial_floats, float* grad, cudaStream_t st) {
  if (!grad) return 0;
  if (g.Cb <= 4) {
    if (g_tc_mode.load()) {
      int r = bn_launch_thin_wgrad_tc(big, small, g, n, partial, partial_floats, grad, st);
      if (r <= 0) return r;
    }
    int r = bn_launch_thin_wgrad(big, small, g, n, partial, partial_floats, grad, st);
    if (r <= 0) return r;
  }
  if (g_tc_mode.load()) {
    int r = bn_launch_wgrad_tc(big, small, g, n, partial, partial_floats, grad, st);
    if (r <= 0) return r;
  }
  return bn_launch_wgrad(big, small, g, n, partial, partial_floats, grad, st);
}

ImgView input_view(const bn_cae_plan* p, const float* x) {
  // frames arrive NCHW (data_generator.py:258-263); with one channel that is already NHWC
  if (p->d.in_c == 1) return nhwc_view(x, p->d.in_h, p->d.in_w, 1);
  return nchw_view(x, p->d.in_h, p->d.in_w, p->d.in_c);
}

}  // namespace

extern "C" int bn_cae_plan_create(const bn_cae_desc* desc, bn_cae_plan** out) {
  if (!desc || !out) BN_FAIL("bn_cae_plan_create: null argument");
  const bn_cae_desc& d = *desc;
  if (d.n_layers < 1 || d.n_layers > BN_MAX_LAYERS) BN_FAIL("n_layers=%d out of range", d.n_layers);
  if (d.n_heads < 1 || d.n_heads > 2) BN_FAIL("n_heads=%d", d.n_heads);
  if (d.n_latents < 1) BN_FAIL("n_latents=%d", d.n_latents);
  bn_cae_plan* p = new bn_cae_plan();
  p->d = d;
  p->nl = d.n_layers;
  p->d_tables = nullptr;
  int H = d.in_h, W = d.in_w, C = d.in_c;
  p->enc_sz[0] = (size_t)H * W * C;
  size_t off = 0;
  for (int i = 0; i < p->nl; ++i) {
    ConvGeom& g = p->enc[i];
    g.Hb = H; g.Wb = W; g.Cb = C;
    g.Hs = d.enc_h[i]; g.Ws = d.enc_w[i]; g.Cs = d.enc_c[i];
    g.k = d.enc_k[i]; g.s = d.enc_s[i]; g.pt = d.enc_pt[i]; g.pl = d.enc_pl[i];
    if (g.k * g.k > BN_MAX_TAPS || g.k < 1 || g.s < 1) { delete p; BN_FAIL("enc layer %d: kernel %d / stride %d unsupported", i, g.k, g.s); }
    int eh = (H + d.enc_pt[i] + d.enc_pb[i] - g.k) / g.s + 1;
    int ew = (W + d.enc_pl[i] + d.enc_pr[i] - g.k) / g.s + 1;
    if (eh != g.Hs || ew != g.Ws) { delete p; BN_FAIL("enc layer %d: dims (%d,%d) inconsistent with padding (expect %d,%d)", i, g.Hs, g.Ws, eh, ew); }
    g.p_w = 2 * i; g.p_b = 2 * i + 1;
    g.off_wf = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    g.off_wd = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    g.off_wft = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    g.off_wdt = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    H = g.Hs; W = g.Ws; C = g.Cs;
    p->enc_sz[i + 1] = (size_t)H * W * C;
  }
  p->feat_c = C; p->feat_h = H; p->feat_w = W;
  p->off_heads = off; off += align64((size_t)d.n_heads * d.n_latents * C * H * W);
  H = d.dec_h0; W = d.dec_w0; C = d.dec_c0;
  p->dec_sz[0] = (size_t)H * W * C;
  for (int i = 0; i < p->nl; ++i) {
    ConvGeom& g = p->dec[i];
    g.Hs = H; g.Ws = W; g.Cs = C;
    g.Hb = d.dec_h[i]; g.Wb = d.dec_w[i]; g.Cb = d.dec_c[i];
    g.k = d.dec_k[i]; g.s = d.dec_s[i]; g.pt = d.dec_pt[i]; g.pl = d.dec_pl[i];
    if (g.k * g.k > BN_MAX_TAPS || g.k < 1 || g.s < 1) { delete p; BN_FAIL("dec layer %d: kernel %d / stride %d unsupported", i, g.k, g.s); }
    int eh = (H - 1) * g.s + g.k - d.dec_pt[i] - d.dec_pb[i];
    int ew = (W - 1) * g.s + g.k - d.dec_pl[i] - d.dec_pr[i];
    if (eh != g.Hb || ew != g.Wb) { delete p; BN_FAIL("dec layer %d: dims (%d,%d) inconsistent with crop (expect %d,%d)", i, g.Hb, g.Wb, eh, ew); }
    g.p_w = 2 * p->nl + 6 + 2 * i; g.p_b = g.p_w + 1;
    g.off_wf = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    g.off_wd = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    g.off_wft = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    g.off_wdt = off; off += align64((size_t)g.k * g.k * g.Cb * g.Cs);
    H = g.Hb; W = g.Wb; C = g.Cb;
    p->dec_sz[i + 1] = (size_t)H * W * C;
  }
  // (the channel counts may differ: a conditional encoder reads label images next to the frame, aes.py:129-137)
  if (H != d.in_h || W != d.in_w) { delete p; BN_FAIL("decoder output (%d,%d) != input (%d,%d)", H, W, d.in_h, d.in_w); }
  if (C > 4) { delete p; BN_FAIL("n_input_channels=%d > 4 has no fused output-layer kernel", C); }
  p->packed_floats = off;
  p->max_act = 0;
  for (int i = 0; i <= p->nl; ++i) p->max_act = std::max(p->max_act, std::max(p->enc_sz[i], p->dec_sz[i]));
  p->enc_f.resize(p->nl); p->enc_d.resize(p->nl); p->dec_f.resize(p->nl); p->dec_d.resize(p->nl);
  for (int i = 0; i < p->nl; ++i) build_classes(p->enc[i], p->h_tables, p->enc_f[i], p->enc_d[i]);
  for (int i = 0; i < p->nl; ++i) build_classes(p->dec[i], p->h_tables, p->dec_f[i], p->dec_d[i]);
  cudaError_t e = cudaMalloc(&p->d_tables, p->h_tables.size() * sizeof(TapClass));
  if (e == cudaSuccess)
    e = cudaMemcpy(p->d_tables, p->h_tables.data(), p->h_tables.size() * sizeof(TapClass), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->d_tables) cudaFree(p->d_tables);
    delete p;
    BN_FAIL("plan tables: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < p->nl; ++i) {
    p->enc[i].d_fprop = p->d_tables + p->enc_f[i]; p->enc[i].d_dgrad = p->d_tables + p->enc_d[i];
    p->dec[i].d_fprop = p->d_tables + p->dec_f[i]; p->dec[i].d_dgrad = p->d_tables + p->dec_d[i];
  }
  *out = p;
  return 0;
}

extern "C" void bn_cae_plan_destroy(bn_cae_plan* p) {
  if (!p) return;
  if (p->d_tables) cudaFree(p->d_tables);
  delete p;
}

extern "C" size_t bn_cae_packed_bytes(const bn_cae_plan* p) { return p ? p->packed_floats * sizeof(float) : 0; }

extern "C" size_t bn_cae_workspace_bytes(const bn_cae_plan* p, int n) {
  if (!p || n <= 0) return 0;
  return ws_layout(p, n).total * sizeof(float);
}

extern "C" int bn_cae_pack_params(bn_cae_plan* p, const float* const* P, void* d_packed, void* stream) {
  if (!p || !P || !d_packed) BN_FAIL("bn_cae_pack_params: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* pk = (float*)d_packed;
  PackJobs jobs;
  jobs.n = 0;
  static_assert(sizeof(jobs.j) / sizeof(jobs.j[0]) >= 2 * BN_MAX_LAYERS, "PackJobs holds every conv layer of a plan");
  auto add = [&](const ConvGeom& g) {
    if (!P[g.p_w]) return;
    PackJob& jb = jobs.j[jobs.n++];
    jb.src = P[g.p_w];
    jb.wf = pk + g.off_wf; jb.wd = pk + g.off_wd; jb.wft = pk + g.off_wft; jb.wdt = pk + g.off_wdt;
    jb.Cs = g.Cs; jb.Cb = g.Cb; jb.kk = g.k * g.k;
  };
  // widest layers first: their blocks are the long ones
  for (int i = p->nl - 1; i >= 0; --i) add(p->enc[i]);
  for (int i = 0; i < p->nl; ++i) add(p->dec[i]);
  BN_TRY(bn_launch_pack_all(jobs, st));
  const int n2 = 2 * p->nl;
  if (P[n2]) {
    if (p->d.n_heads == 2 && !P[n2 + 2]) BN_FAIL("bn_cae_pack_params: logvar head weight missing");
    BN_TRY(bn_launch_pack_heads(P[n2], p->d.n_heads == 2 ? P[n2 + 2] : nullptr, p->d.n_latents, p->feat_c,
                                p->feat_h, p->feat_w, pk + p->off_heads, st));
  }
  return 0;
}

static int encode_impl(bn_cae_plan* p, int n, const float* d_x, const unsigned char* d_x8, const float* const* P,
                       const void* d_packed, void* d_ws, float* d_mu, float* d_logvar, void* stream) {
  if (!p || (!d_x && !d_x8) || !P || !d_packed || !d_ws || !d_mu) BN_FAIL("bn_cae_encode: null argument");
  if (p->d.n_heads == 2 && !d_logvar) BN_FAIL("bn_cae_encode: logvar output required for a variational encoder");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* pk = (const float*)d_packed;
  float* ws = (float*)d_ws;
  WsLayout L = ws_layout(p, n);
  t_split_buf = ws + L.partial;
  t_split_floats = L.partial_floats;
  ImgView in = input_view(p, d_x);     // d_x == NULL with uint8 frames: only the strides are used
  for (int i = 0; i < p->nl; ++i) {
    const ConvGeom& g = p->enc[i];
    float* out = ws + L.enc_act[i + 1];
    const unsigned char* u8 = i == 0 ? d_x8 : nullptr;
    int thin = g.Cb <= 4 ? bn_launch_thin_fprop(in, g, pk + g.off_wf, g_tc_mode.load() ? pk + g.off_wft : nullptr, P[g.p_b], out, nullptr,
                                                BN_ACT_LEAKY, n, st, u8) : 1;
    if (thin < 0) return thin;
    if (thin > 0 && u8)
      BN_FAIL("bn_cae_encode_u8: the uint8 frame loader exists for first layers with <= 4 input channels, "
              "kernel 5, stride 2 and a multiple of 32 output channels (got C=%d k=%d s=%d C_out=%d)",
              g.Cb, g.k, g.s, g.Cs);
    if (thin > 0)
      BN_TRY(run_igemm(p, in, pk + g.off_wf, pk + g.off_wft, g.k * g.k * g.Cb, P[g.p_b], out, g.Hs, g.Ws,
                       g.Cs, nullptr, p->enc_f[i], 1, g.Hs * g.Ws, g.s, 1, n, BN_ACT_LEAKY, st));
    in = nhwc_view(out, g.Hs, g.Ws, g.Cs);
  }
  const int n2 = 2 * p->nl;
  BN_TRY(bn_launch_heads_fwd(ws + L.enc_act[p->nl], pk + p->off_heads, P[n2 + 1],
                             p->d.n_heads == 2 ? P[n2 + 3] : nullptr, n, p->d.n_latents, p->d.n_heads,
                             p->feat_c * p->feat_h * p->feat_w, d_mu, d_logvar, st));
  return 0;
}

extern "C" int bn_cae_encode(bn_cae_plan* p, int n, const float* d_x, const float* const* P,
                             const void* d_packed, void* d_ws, float* d_mu, float* d_logvar, void* stream) {
  if (!d_x) BN_FAIL("bn_cae_encode: null argument");
  return encode_impl(p, n, d_x, nullptr, P, d_packed, d_ws, d_mu, d_logvar, stream);
}

extern "C" int bn_cae_encode_u8(bn_cae_plan* p, int n, const uint8_t* d_x, const float* const* P,
                                const void* d_packed, void* d_ws, float* d_mu, float* d_logvar, void* stream) {
  if (!d_x) BN_FAIL("bn_cae_encode_u8: null argument");
  return encode_impl(p, n, nullptr, d_x, P, d_packed, d_ws, d_mu, d_logvar, stream);
}

extern "C" int bn_cae_decode(bn_cae_plan* p, int n, const float* d_z, const float* const* P,
                             const void* d_packed, void* d_ws, float* d_xhat, const float* d_target,
                             const float* d_mask, int chunk_size, int frame_offset, int n_total,
                             float grad_coef, double* d_sse, void* stream) {
  if (!p || !d_z || !P || !d_packed || !d_ws) BN_FAIL("bn_cae_decode: null argument");
  if (d_target && !d_sse) BN_FAIL("bn_cae_decode: d_sse required with d_target");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* pk = (const float*)d_packed;
  float* ws = (float*)d_ws;
  WsLayout L = ws_layout(p, n);
  t_split_buf = ws + L.partial;
  t_split_floats = L.partial_floats;
  const int n2 = 2 * p->nl;
  BN_CUDA(cudaMemcpyAsync(ws + L.zcopy, d_z, (size_t)n * p->d.n_latents * sizeof(float), cudaMemcpyDeviceToDevice, st));
  BN_TRY(bn_launch_decff_fwd(d_z, P[n2 + 4], P[n2 + 5], n, p->d.n_latents, p->d.dec_c0, p->d.dec_h0,
                             p->d.dec_w0, ws + L.dec_act[0], st));
  for (int i = 0; i + 1 < p->nl; ++i) {
    const ConvGeom& g = p->dec[i];
    ImgView in = nhwc_view(ws + L.dec_act[i], g.Hs, g.Ws, g.Cs);
    BN_TRY(run_igemm(p, in, pk + g.off_wd, pk + g.off_wdt, g.k * g.k * g.Cs, P[g.p_b],
                     ws + L.dec_act[i + 1], g.Hb, g.Wb, g.Cb, nullptr, p->dec_d[i], g.n_dgrad,
                     g.dgrad_maxM, 1, g.s, n, BN_ACT_LEAKY, st));
  }
  const ConvGeom& g = p->dec[p->nl - 1];
  // training pass (target given, no reconstruction requested): GEMM-over-all-taps + col2im on the tensor
  // cores; a reconstruction the caller asks for always comes from the fp32 kernel (1e-4 contract)
  int fast = (g_tc_mode.load() && d_target && !d_xhat)
                 ? bn_launch_thin_dgrad_tc(ws + L.dec_act[p->nl - 1], g, pk + g.off_wdt, P[g.p_b], n, ws + L.dec_act[p->nl],
                                           d_target, d_mask, chunk_size, frame_offset, n_total, grad_coef, d_sse,
                                           ws + L.dpre_last, st)
                 : 1;
  if (fast < 0) return fast;
  if (fast > 0)
    fast = bn_launch_thin_dgrad5(ws + L.dec_act[p->nl - 1], g, pk + g.off_wd, P[g.p_b], n, ws + L.dec_act[p->nl],
                                 d_xhat, d_target, d_mask, chunk_size, frame_offset, n_total, grad_coef, d_sse,
                                 ws + L.dpre_last, st);
  if (fast < 0) return fast;
  if (fast > 0)
    BN_TRY(bn_launch_thin_dgrad(ws + L.dec_act[p->nl - 1], g, pk + g.off_wd, P[g.p_b], n, ws + L.dec_act[p->nl],
                                d_xhat, d_target, d_mask, chunk_size, frame_offset, n_total, grad_coef, d_sse,
                                ws + L.dpre_last, st));
  return 0;
}

extern "C" int bn_cae_decode_bwd(bn_cae_plan* p, int n, const float* d_dxhat, const float* const* P,
                                 const void* d_packed, void* d_ws, float* const* G, float* d_dz,
                                 void* stream) {
  if (!p || !P || !d_packed || !d_ws || !G) BN_FAIL("bn_cae_decode_bwd: null argument");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* pk = (const float*)d_packed;
  float* ws = (float*)d_ws;
  WsLayout L = ws_layout(p, n);
  t_split_buf = ws + L.partial;
  t_split_floats = L.partial_floats;
  const int n2 = 2 * p->nl;
  const ConvGeom& gl = p->dec[p->nl - 1];
  if (d_dxhat)
    BN_TRY(bn_launch_sigmoid_bwd(d_dxhat, ws + L.dec_act[p->nl], ws + L.dpre_last, n, gl.Cb, gl.Hb, gl.Wb, st));
  float* gcur = ws + L.dpre_last;
  float* pp[2] = {ws + L.gA, ws + L.gB};
  int flip = 0;
  int bias_done = 0;   // the kernel that produced gcur already accumulated its column sums
  bn_wgrad_reduce_defer_begin();        // per-layer partial regions; all reductions in one launch below
  const int rc = [&]() -> int {
  for (int i = p->nl - 1; i >= 0; --i) {
    const ConvGeom& g = p->dec[i];
    ImgView big = nhwc_view(gcur, g.Hb, g.Wb, g.Cb);
    const float* small = ws + L.dec_act[i];
    BN_TRY(run_wgrad(big, small, g, n, ws + L.wg_dec[i], L.wg_floats_dec[i], G[g.p_w], st));
    if (!bias_done) BN_TRY(bn_launch_colsum(gcur, (long long)n * g.Hb * g.Wb, g.Cb, G[g.p_b], st));
    bias_done = 0;
    float* out = pp[flip];
    flip ^= 1;
    int thin = g.Cb <= 4 ? bn_launch_thin_fprop(big, g, pk + g.off_wf, g_tc_mode.load() ? pk + g.off_wft : nullptr, nullptr, out,
                                                i > 0 ? small : nullptr,
                                                BN_ACT_NONE, n, st, nullptr, i > 0 ? G[p->dec[i - 1].p_b] : nullptr,
                                                &bias_done) : 1;
    if (thin < 0) return thin;
    if (thin > 0)
      BN_TRY(run_igemm(p, big, pk + g.off_wf, pk + g.off_wft, g.k * g.k * g.Cb, nullptr, out, g.Hs, g.Ws,
                       g.Cs, i > 0 ? small : nullptr, p->dec_f[i], 1, g.Hs * g.Ws, g.s, 1, n, BN_ACT_NONE, st,
                       i > 0 ? G[p->dec[i - 1].p_b] : nullptr, &bias_done, ws + L.cp_dec[i], L.cp_floats_dec[i]));
    gcur = out;
  }
  return 0;
  }();
  const int rf = bn_wgrad_reduce_flush(st);
  if (rc) return rc;
  if (rf) return rf;
  BN_TRY(bn_launch_decff_bwd(ws + L.zcopy, P[n2 + 4], gcur, n, p->d.n_latents, p->d.dec_c0, p->d.dec_h0,
                             p->d.dec_w0, d_dz, G[n2 + 4], G[n2 + 5], st));
  return 0;
}

namespace {
// phase 0: the whole pass.  Phases 1 + 2 are the same launches in two calls, split where the gradients of the
// heads and of the top conv layer are final (heads backward, top layer weight gradient + bias sums, their batched
// reduction), so that a data-parallel caller can start the all-reduce of that bucket -- three quarters of the
// encoder's parameters in the default architecture -- underneath phase 2 (backward-data of the top layer, then
// every layer below).  The upstream-gradient image of the top layer stays in the workspace between the calls.
int encode_bwd_impl(bn_cae_plan* p, int n, const float* d_x, const float* d_dmu, const float* d_dlogvar,
                    const float* const* P, const void* d_packed, void* d_ws, float* const* G, void* stream,
                    int phase) {
  if (!p || !d_x || !P || !d_packed || !d_ws || !G) BN_FAIL("bn_cae_encode_bwd: null argument");
  if (!d_dmu && !d_dlogvar) BN_FAIL("bn_cae_encode_bwd: no upstream gradient");
  if (phase < 0 || phase > 2) BN_FAIL("bn_cae_encode_bwd: phase %d", phase);
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* pk = (const float*)d_packed;
  float* ws = (float*)d_ws;
  WsLayout L = ws_layout(p, n);
  t_split_buf = ws + L.partial;
  t_split_floats = L.partial_floats;
  const int n2 = 2 * p->nl;
  float* pp[2] = {ws + L.gA, ws + L.gB};
  int flip = 0;
  float* gcur = pp[flip];
  flip ^= 1;
  if (phase != 2)
    BN_TRY(bn_launch_heads_bwd(ws + L.enc_act[p->nl], pk + p->off_heads, d_dmu, d_dlogvar, n, p->d.n_latents,
                               p->feat_c, p->feat_h, p->feat_w, gcur, G[n2], G[n2 + 1],
                               p->d.n_heads == 2 ? G[n2 + 2] : nullptr, p->d.n_heads == 2 ? G[n2 + 3] : nullptr, st));
  int bias_done = 0;   // the kernel that produced gcur already accumulated its column sums
  bn_wgrad_reduce_defer_begin();
  const int rc = [&]() -> int {
  for (int i = p->nl - 1; i >= 0; --i) {
    const ConvGeom& g = p->enc[i];
    ImgView big = i == 0 ? input_view(p, d_x) : nhwc_view(ws + L.enc_act[i], g.Hb, g.Wb, g.Cb);
    if (!(phase == 2 && i == p->nl - 1)) {
      BN_TRY(run_wgrad(big, gcur, g, n, ws + L.wg_enc[i], L.wg_floats_enc[i], G[g.p_w], st));
      if (!bias_done) BN_TRY(bn_launch_colsum(gcur, (long long)n * g.Hs * g.Ws, g.Cs, G[g.p_b], st));
    }
    bias_done = 0;
    if (phase == 1) return 0;
    if (i > 0) {
      float* out = pp[flip];
      flip ^= 1;
      ImgView in = nhwc_view(gcur, g.Hs, g.Ws, g.Cs);
      BN_TRY(run_igemm(p, in, pk + g.off_wd, pk + g.off_wdt, g.k * g.k * g.Cs, nullptr, out, g.Hb, g.Wb,
                       g.Cb, ws + L.enc_act[i], p->enc_d[i], g.n_dgrad, g.dgrad_maxM, 1, g.s, n,
                       BN_ACT_NONE, st, G[p->enc[i - 1].p_b], &bias_done, ws + L.cp_enc[i], L.cp_floats_enc[i]));
      gcur = out;
    }
  }
  return 0;
  }();
  const int rf = bn_wgrad_reduce_flush(st);
  return rc ? rc : rf;
}
}  // namespace

extern "C" int bn_cae_encode_bwd(bn_cae_plan* p, int n, const float* d_x, const float* d_dmu,
                                 const float* d_dlogvar, const float* const* P, const void* d_packed,
                                 void* d_ws, float* const* G, void* stream) {
  return encode_bwd_impl(p, n, d_x, d_dmu, d_dlogvar, P, d_packed, d_ws, G, stream, 0);
}

extern "C" int bn_cae_encode_bwd_phase(bn_cae_plan* p, int n, const float* d_x, const float* d_dmu,
                                       const float* d_dlogvar, const float* const* P, const void* d_packed,
                                       void* d_ws, float* const* G, void* stream, int phase) {
  return encode_bwd_impl(p, n, d_x, d_dmu, d_dlogvar, P, d_packed, d_ws, G, stream, phase);
}

// Single-layer entry point: kernel-level parity tests (tensor-core vs CUDA-core kernels on the same
// tensors) and per-kernel timing for the roofline line of bench.py.
extern "C" int bn_cae_layer_op(bn_cae_plan* p, int side, int layer, int op, int n, const float* d_in,
                               const float* d_in2, float* d_out, const float* const* P,
                               const void* d_packed, void* d_ws, void* stream) {
  if (!p || !d_in || !d_out || !d_packed || !d_ws) BN_FAIL("bn_cae_layer_op: null argument");
  if (side < 0 || side > 1 || layer < 0 || layer >= p->nl || op < 0 || op > 3) BN_FAIL("bn_cae_layer_op: bad selector");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* pk = (const float*)d_packed;
  float* ws = (float*)d_ws;
  WsLayout L = ws_layout(p, n);
  t_split_buf = ws + L.partial;
  t_split_floats = L.partial_floats;
  const ConvGeom& g = side == 0 ? p->enc[layer] : p->dec[layer];
  const int tf = side == 0 ? p->enc_f[layer] : p->dec_f[layer];
  const int td = side == 0 ? p->enc_d[layer] : p->dec_d[layer];
  const float* bias = (P && op == 0) ? P[g.p_b] : nullptr;
  const bool fprop_form = (side == 0 && op == 0) || (side == 1 && op == 1);
  if (op == 3) {
    // last decoder layer forward with the fused loss (chunks of 200 frames, gaussian-ll coefficient):
    // d_in = small image, d_in2 = target (n, C, H, W); d_out = [x_hat (n C H W) | dL/dpre (n H W C) | sse doubles]
    if (side != 1 || layer != p->nl - 1 || !d_in2) BN_FAIL("bn_cae_layer_op: op 3 is the last decoder layer with a target");
    const size_t tot = (size_t)n * g.Cb * g.Hb * g.Wb;
    double* sse = reinterpret_cast<double*>(d_out + 2 * tot);
    const int nchunks = (n + 199) / 200;
    BN_CUDA(cudaMemsetAsync(sse, 0, sizeof(double) * nchunks, st));
    int fast = g_tc_mode.load() ? bn_launch_thin_dgrad_tc(d_in, g, pk + g.off_wdt, P ? P[g.p_b] : nullptr, n, d_out, d_in2,
                                                          nullptr, 200, 0, n, 1.f, sse, d_out + tot, st)
                                : 1;
    if (fast > 0)
      fast = bn_launch_thin_dgrad5(d_in, g, pk + g.off_wd, P ? P[g.p_b] : nullptr, n, d_out, nullptr, d_in2, nullptr, 200, 0,
                                   n, 1.f, sse, d_out + tot, st);
    if (fast > 0) BN_FAIL("bn_cae_layer_op: op 3 needs the kernel-5 / stride-2 thin output layer");
    return fast;
  }
  if (op == 2) {
    if (!d_in2) BN_FAIL("bn_cae_layer_op: wgrad needs both images");
    return run_wgrad(nhwc_view(d_in, g.Hb, g.Wb, g.Cb), d_in2, g, n, ws + L.partial, L.partial_floats, d_out, st);
  }
  if (fprop_form) {
    if (g.Cb <= 4) {
      // thin first / last layer: same routing as bn_cae_encode / bn_cae_decode_bwd (d_in is NHWC here)
      int thin = bn_launch_thin_fprop(nhwc_view(d_in, g.Hb, g.Wb, g.Cb), g, pk + g.off_wf,
                                      g_tc_mode.load() ? pk + g.off_wft : nullptr, bias, d_out, nullptr,
                                      op == 0 ? BN_ACT_LEAKY : BN_ACT_NONE, n, st);
      if (thin <= 0) return thin;
    }
    return run_igemm(p, nhwc_view(d_in, g.Hb, g.Wb, g.Cb), pk + g.off_wf, pk + g.off_wft, g.k * g.k * g.Cb, bias,
                     d_out, g.Hs, g.Ws, g.Cs, nullptr, tf, 1, g.Hs * g.Ws, g.s, 1, n,
                     op == 0 ? BN_ACT_LEAKY : BN_ACT_NONE, st);
  }
  if (side == 1 && op == 0 && layer == p->nl - 1) {
    return bn_launch_thin_dgrad(d_in, g, pk + g.off_wd, bias, n, ws + L.dec_act[p->nl], d_out, nullptr, nullptr, 0,
                                0, n, 0.f, nullptr, ws + L.dpre_last, st);
  }
  return run_igemm(p, nhwc_view(d_in, g.Hs, g.Ws, g.Cs), pk + g.off_wd, pk + g.off_wdt, g.k * g.k * g.Cs, bias,
                   d_out, g.Hb, g.Wb, g.Cb, nullptr, td, g.n_dgrad, g.dgrad_maxM, 1, g.s, n,
                   op == 0 ? BN_ACT_LEAKY : BN_ACT_NONE, st);
}
