// model.cu -- acoustic model files -> host model -> packed device images.
//
// File formats and load-time arithmetic follow the reference:
//   .mc  HmmSet::read_mc            aku/HmmSet.cc:157-180, Mixture::read aku/Distributions.cc:2419-2434
//        (weights re-normalised to sum 1: Mixture::normalize_weights :2068-2075)
//   .ph  HmmSet::read_legacy_ph     aku/HmmSet.cc:209-329  (state index == mixture index,
//        num_states = highest referenced mixture index + 1)
//   .gk  PDFPool::read_gk           aku/Distributions.cc:2812-2910, DiagonalGaussian::read :1132-1150
//        precision = 1/cov if cov > 0 else 0;  constant = log(sqrt(prod precision)) when the
//        product is > 0, otherwise the raw product is kept (DiagonalGaussian::set_constant :1274-1288)
#include "ctx.hpp"
#include "kernels.hpp"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <fstream>
#include <sstream>

namespace akugpu {

namespace {

struct Tokens {
  std::string buf;
  size_t pos = 0;
  std::string path;
  explicit Tokens(const std::string &p) : path(p) {
    std::ifstream in(p.c_str(), std::ios::binary);
    if (!in) throw Error(AKUGPU_E_IO, "could not open " + p);
    std::ostringstream ss;
    ss << in.rdbuf();
    buf = ss.str();
  }
  void skip() { while (pos < buf.size() && isspace((unsigned char)buf[pos])) pos++; }
  bool eof() { skip(); return pos >= buf.size(); }
  std::string word() {
    skip();
    size_t b = pos;
    while (pos < buf.size() && !isspace((unsigned char)buf[pos])) pos++;
    if (b == pos) throw Error(AKUGPU_E_MODEL, "unexpected end of file in " + path);
    return buf.substr(b, pos - b);
  }
  long integer() {
    skip();
    char *e;
    long v = strtol(buf.c_str() + pos, &e, 10);
    if (e == buf.c_str() + pos) throw Error(AKUGPU_E_MODEL, "expected integer in " + path);
    pos = e - buf.c_str();
    return v;
  }
  double real() {
    skip();
    char *e;
    double v = strtod(buf.c_str() + pos, &e);
    if (e == buf.c_str() + pos) throw Error(AKUGPU_E_MODEL, "expected number in " + path);
    pos = e - buf.c_str();
    return v;
  }
};

}  // namespace

void model_read_files(const std::string &gk_path, const std::string &mc_path, const std::string &ph_path, HostModel &hm)
{
  // --- .mc ---
  std::vector<std::vector<int32_t>> mc_idx;
  std::vector<std::vector<double>> mc_w;
  {
    Tokens t(mc_path);
    long n = t.integer();
    if (n < 0) throw Error(AKUGPU_E_MODEL, "negative mixture count in " + mc_path);
    mc_idx.resize(n);
    mc_w.resize(n);
    for (long i = 0; i < n; i++) {
      long k = t.integer();
      for (long j = 0; j < k; j++) {
        mc_idx[i].push_back((int32_t)t.integer());
        mc_w[i].push_back(t.real());
      }
    }
  }
  // --- .ph --- only the state -> mixture binding matters for scoring
  int n_states = 0;
  {
    Tokens t(ph_path);
    if (t.word() != "PHONE") throw Error(AKUGPU_E_MODEL, ph_path + ": first token is not PHONE");
    long phones = t.integer();
    for (long h = 0; h < phones; h++) {
      t.integer();                 // index
      long states = t.integer() - 2;
      t.word();                    // label
      t.integer(); t.integer();    // -1 -2
      for (long s = 0; s < states; s++) {
        long pdf = t.integer();
        if (pdf + 1 > n_states) n_states = (int)pdf + 1;
      }
      for (long s = -2; s < states; s++) {
        t.integer();               // source
        long ntr = t.integer();
        for (long k = 0; k < ntr; k++) { t.integer(); t.real(); }
      }
    }
  }
  if (n_states > (int)mc_idx.size())
    throw Error(AKUGPU_E_MODEL, fmt("%s refers to mixture %d but %s has only %d", ph_path.c_str(), n_states - 1,
                                    mc_path.c_str(), (int)mc_idx.size()));
  // --- .gk ---
  {
    Tokens t(gk_path);
    long G = t.integer(), D = t.integer();
    std::string type = t.word();
    if (G < 0 || D <= 0) throw Error(AKUGPU_E_MODEL, "bad header in " + gk_path);
    hm.G = (int)G; hm.D = (int)D;
    hm.mean.assign((size_t)G * D, 0.0);
    hm.cov.assign((size_t)G * D, 0.0);
    bool variable = (type == "variable");
    if (!variable && type != "diagonal_cov") {
      if (type == "full_cov") throw Error(AKUGPU_E_MODEL, "full-covariance pools are not supported by the diagonal scorer");
      throw Error(AKUGPU_E_MODEL, "Unknown model type " + type);
    }
    for (long g = 0; g < G; g++) {
      if (variable) {
        std::string gt = t.word();
        if (gt != "diag")
          throw Error(AKUGPU_E_MODEL, "Gaussian type '" + gt + "' is not supported by the diagonal scorer");
      }
      for (long d = 0; d < D; d++) hm.mean[g * D + d] = t.real();
      for (long d = 0; d < D; d++) hm.cov[g * D + d] = t.real();
    }
  }
  hm.S = n_states;
  hm.mix_off.assign(1, 0);
  hm.mix_gauss.clear();
  hm.mix_w.clear();
  for (int s = 0; s < n_states; s++) {
    double sum = 0;
    for (size_t k = 0; k < mc_w[s].size(); k++) sum += mc_w[s][k];
    for (size_t k = 0; k < mc_w[s].size(); k++) {
      if (mc_idx[s][k] < 0 || mc_idx[s][k] >= hm.G)
        throw Error(AKUGPU_E_MODEL, fmt("mixture %d refers to Gaussian %d of %d", s, mc_idx[s][k], hm.G));
      hm.mix_gauss.push_back(mc_idx[s][k]);
      hm.mix_w.push_back(mc_w[s][k] / sum);
    }
    hm.mix_off.push_back((int32_t)hm.mix_gauss.size());
  }
}

template <class T>
static void upload(DevBuf &b, const std::vector<T> &v, cudaStream_t st)
{
  b.reserve(std::max<size_t>(v.size() * sizeof(T), 16));
  if (!v.empty()) AKU_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

void model_pack(akugpu_ctx *ctx)
{
  const HostModel &hm = ctx->hm;
  const int S = hm.S, G = hm.G, D = hm.D;
  // Gaussian-level quantities in double, as the reference computes them at load time.
  std::vector<double> prec((size_t)G * D), cst(G);
  for (int g = 0; g < G; g++) {
    double c = 1;
    for (int d = 0; d < D; d++) {
      double cv = hm.cov[(size_t)g * D + d];
      double p = cv > 0 ? 1 / cv : 0;
      prec[(size_t)g * D + d] = p;
      c *= p;
    }
    if (c > 0) c = log(sqrt(c));
    cst[g] = c;
  }
  // ---- fp64 image: the reference's own arrays ----
  {
    PackedF64 &p = ctx->p64;
    upload(p.mean, hm.mean, ctx->stream);
    upload(p.prec, prec, ctx->stream);
    upload(p.cst, cst, ctx->stream);
    upload(p.mix_off, hm.mix_off, ctx->stream);
    upload(p.mix_gauss, hm.mix_gauss, ctx->stream);
    upload(p.mix_w, hm.mix_w, ctx->stream);
    ctx->have_p64 = true;
  }
  // ---- fp32 image: slots of 16 components dealt into 8 warp queues (gmm_kernels.cu) ----
  PackedF32 &p = ctx->p32;
  const int NW = 8, GR = 16, TC = NW * GR, META_INTS = 8;
  p.packed_ffma2 = ctx->scorer_variant != 1;
  p.ranges.clear();
  p.DP = (D + 1) / 2;
  p.tile_floats = (size_t)p.DP * TC * 4 + TC + META_INTS;

  std::vector<double> cen(2 * p.DP, 0.0);
  if (G > 0)
    for (int d = 0; d < D; d++) {
      double s = 0;
      for (int g = 0; g < G; g++) s += hm.mean[(size_t)g * D + d];
      cen[d] = s / G;
    }
  // Deal the states to the shortest queue, in state order; a state's slots stay consecutive.
  struct Slot { int state, k0, first, last; };
  std::vector<std::vector<Slot>> queue(NW);
  for (int s = 0; s < S; s++) {
    const int K = hm.mix_off[s + 1] - hm.mix_off[s];
    const int ns = std::max(1, (K + GR - 1) / GR);
    int q = 0;
    for (int w = 1; w < NW; w++) if (queue[w].size() < queue[q].size()) q = w;
    for (int i = 0; i < ns; i++) queue[q].push_back(Slot{s, i * GR, i == 0, i == ns - 1});
  }
  size_t T = 0;
  for (int w = 0; w < NW; w++) T = std::max(T, queue[w].size());
  p.n_tiles = (int)T;
  p.clean.assign(T, 1);
  std::vector<float> img(T * p.tile_floats, 0.f);
  for (size_t t = 0; t < T; t++) {
    float *P = img.data() + t * p.tile_floats;
    float *C = P + (size_t)p.DP * TC * 4;
    int32_t *meta = reinterpret_cast<int32_t *>(C + TC);
    for (int i = 0; i < TC; i++) C[i] = 1.0e30f;
    for (int w = 0; w < NW; w++) {
      if (t >= queue[w].size()) { meta[w] = -1; continue; }   // padding slot: computed, never written
      const Slot &sl = queue[w][t];
      meta[w] = (sl.state << 2) | (sl.first << 1) | sl.last;
      if (!sl.first) p.clean[t] = 0;
      const int K = hm.mix_off[sl.state + 1] - hm.mix_off[sl.state];
      for (int j = 0; j < GR && sl.k0 + j < K; j++) {
        const int k = hm.mix_off[sl.state] + sl.k0 + j;
        const int g = hm.mix_gauss[k];
        const double wgt = hm.mix_w[k];
        const double c = (wgt > 0 ? log(wgt) : -1.0e30) + cst[g];
        const int comp = w * GR + j;
        C[comp] = (c > -1.0e30) ? (float)(-c) : 1.0e30f;
        for (int d = 0; d < D; d++) {
          float sf = (float)sqrt(0.5 * prec[(size_t)g * D + d]);
          float mf = (float)(-(hm.mean[(size_t)g * D + d] - cen[d]) * (double)sf);
          float *q = P + ((size_t)(d >> 1) * TC + comp) * 4;
          q[d & 1] = sf;
          q[2 + (d & 1)] = mf;
        }
      }
    }
  }
  std::vector<float> cenf(cen.begin(), cen.end());
  upload(p.params, img, ctx->stream);
  upload(p.center, cenf, ctx->stream);
  upload(p.center64, cen, ctx->stream);
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_model = true;
}

}  // namespace akugpu
