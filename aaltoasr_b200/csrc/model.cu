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
#include "tc_common.cuh"
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

// ---------------------------------------------------------------------------------------------
// Gaussian clustering (phone_probs -C file.gcl --eval-minc x --eval-ming y; aku/phone_probs.cc:112-117).
// .gcl: number of clusters, then `gauss_index cluster_index` pairs (PDFPool::read_clustering, aku/Distributions.cc:3115-3147).
// The reference reads pairs in a `while (in) { in >> g >> c; ... }` loop: after the last pair the stream is still good, the
// next extraction fails at end of file without touching g and c, and the last pair is recorded a SECOND time.  That
// duplicate counts in the cluster's size (the min-Gaussians budget) and weighs twice in its centre, so it is kept.
void model_set_clustering(HostModel &hm, int n_clusters, const int32_t *gauss_index, const int32_t *cluster_index, int64_t n_pairs)
{
  const int G = hm.G, D = hm.D;
  if (hm.n_full > 0) throw Error(AKUGPU_E_MODEL, "Gaussian clustering is supported for diagonal pools only");
  if (n_clusters < 0 || (double)n_clusters > 0.3 * G)
    throw Error(AKUGPU_E_MODEL, fmt("PDFPool::read_clustering(): Number of clusters (%d) seems insensible compared to the number of Gaussians (%d).",
                                    n_clusters, G));
  hm.clear_clustering();
  hm.n_clusters = n_clusters;
  hm.cluster_gauss.assign(n_clusters, std::vector<int32_t>());
  hm.gauss_cluster.assign(G, -1);
  for (int64_t i = 0; i < n_pairs; i++) {
    const int g = gauss_index[i], c = cluster_index[i];
    if (g >= G || g < 0) throw Error(AKUGPU_E_MODEL, "PDFPool::read_clustering(): Gauss index out of bounds\n");
    if (c >= n_clusters || c < 0) throw Error(AKUGPU_E_MODEL, "PDFPool::read_clustering(): Cluster index out of bounds\n");
    if (hm.gauss_cluster[g] >= 0 && hm.gauss_cluster[g] != c)
      throw Error(AKUGPU_E_MODEL, fmt("Gaussian %d is assigned to clusters %d and %d: not supported", g, hm.gauss_cluster[g], c));
    hm.gauss_cluster[g] = c;
    hm.cluster_gauss[c].push_back(g);
  }
  // centres: Gaussian::merge with unit weights, diagonal kept (DiagonalGaussian::set_covariance :1208-1234), in the
  // reference's order of operations: sum of (cov + mu mu), sum of mu, scale by the reciprocal, subtract mean*mean
  hm.c_mean.assign((size_t)n_clusters * D, 0.0);
  hm.c_cov.assign((size_t)n_clusters * D, 0.0);
  for (int c = 0; c < n_clusters; c++) {
    const std::vector<int32_t> &L = hm.cluster_gauss[c];
    double wsum = 0;
    for (size_t i = 0; i < L.size(); i++) wsum += 1.0;
    const bool even = wsum < 1e-15;
    const double cw = even ? 1.0 / (double)L.size() : 1.0;      // empty cluster: 0/0 in the reference as well
    if (even) wsum = 1;
    double *mu = &hm.c_mean[(size_t)c * D], *cv = &hm.c_cov[(size_t)c * D];
    for (size_t i = 0; i < L.size(); i++)
      for (int d = 0; d < D; d++) {
        const double m = hm.mean[(size_t)L[i] * D + d];
        const double cur = hm.cov[(size_t)L[i] * D + d] + m * m;
        cv[d] += cw * cur;
        mu[d] += cw * m;
      }
    const double r = 1.0 / wsum;
    for (int d = 0; d < D; d++) { mu[d] *= r; cv[d] *= r; }
    for (int d = 0; d < D; d++) cv[d] += -1.0 * mu[d] * mu[d];
  }
}

void model_read_clustering(const std::string &gcl_path, HostModel &hm)
{
  std::ifstream in(gcl_path.c_str());
  if (!in) throw Error(AKUGPU_E_IO, "PDFPool::read_clustering(): could not open " + gcl_path);
  int n = 0;
  in >> n;
  std::vector<int32_t> gi, ci;
  int g = 0, c = 0;
  bool any = false;
  while (in) {                       // see above: the failed read after the last pair repeats it
    int a, b;
    if (in >> a >> b) { g = a; c = b; any = true; }
    else if (!any) break;            // no pair at all: the reference would use uninitialised indices
    gi.push_back(g);
    ci.push_back(c);
  }
  model_set_clustering(hm, n, gi.data(), ci.data(), (int64_t)gi.size());
}

void model_read_files(const std::string &gk_path, const std::string &mc_path, const std::string &ph_path, HostModel &hm)
{
  hm.clear_clustering();
  hm.clear_cmllr();
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
    hm.ph_label.clear();
    hm.ph_states.clear();
    for (long h = 0; h < phones; h++) {
      t.integer();                 // index
      long states = t.integer() - 2;
      hm.ph_label.push_back(t.word());     // label (UNIT_PHONE transforms of a speaker file are resolved with it)
      hm.ph_states.push_back(std::vector<int32_t>());
      t.integer(); t.integer();    // -1 -2
      for (long s = 0; s < states; s++) {
        long pdf = t.integer();
        if (pdf + 1 > n_states) n_states = (int)pdf + 1;
        hm.ph_states.back().push_back((int32_t)pdf);
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
    if (!variable && type != "diagonal_cov" && type != "full_cov") throw Error(AKUGPU_E_MODEL, "Unknown model type " + type);
    hm.full_index.assign(G, -1);
    hm.full_cov.clear();
    hm.n_full = 0;
    for (long g = 0; g < G; g++) {
      bool full = (type == "full_cov");
      if (variable) {
        std::string gt = t.word();
        if (gt == "full") full = true;
        else if (gt != "diag")    // precision_subspace / pcgmm / scgmm need USE_SUBSPACE_COV, which no build defines
          throw Error(AKUGPU_E_MODEL, "Unknown model type\n" + gt);
      }
      for (long d = 0; d < D; d++) hm.mean[g * D + d] = t.real();
      if (full) {
        hm.full_index[g] = hm.n_full++;
        for (long d = 0; d < D * D; d++) hm.full_cov.push_back(t.real());
      } else {
        for (long d = 0; d < D; d++) hm.cov[g * D + d] = t.real();
      }
    }
    if (hm.n_full == 0) hm.full_index.clear();
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

// ---- full-covariance load-time algebra (FullCovarianceGaussian::set_covariance, aku/Distributions.cc:1560-1586)
// Cholesky A = L L^T of a symmetric matrix; false if not positive definite (the reference's is_spd test,
// aku/LinearAlgebra.cc:421-434, asks for all eigenvalues > 0, which is the same condition).
bool host_cholesky(const std::vector<double> &A, int n, std::vector<double> &Lw)
{
  Lw.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= Lw[(size_t)j * n + k] * Lw[(size_t)j * n + k];
    if (!(d > 0)) return false;
    d = sqrt(d);
    Lw[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double v = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) v -= Lw[(size_t)i * n + k] * Lw[(size_t)j * n + k];
      Lw[(size_t)i * n + j] = v / d;
    }
  }
  return true;
}
// LU with partial pivoting + inverse, the algorithm behind LinearAlgebra::inverse (aku/LinearAlgebra.cc:508-515).
void host_lu_inverse(const std::vector<double> &M, int n, std::vector<double> &inv)
{
  std::vector<double> A(M);
  std::vector<int> piv(n);
  for (int k = 0; k < n; k++) {
    int p = k; double best = fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; i++) if (fabs(A[(size_t)i * n + k]) > best) { best = fabs(A[(size_t)i * n + k]); p = i; }
    piv[k] = p;
    if (p != k) for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
    if (A[(size_t)k * n + k] != 0.0)
      for (int i = k + 1; i < n; i++) {
        A[(size_t)i * n + k] /= A[(size_t)k * n + k];
        for (int j = k + 1; j < n; j++) A[(size_t)i * n + j] -= A[(size_t)i * n + k] * A[(size_t)k * n + j];
      }
  }
  inv.assign((size_t)n * n, 0.0);
  std::vector<double> b(n);
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) b[i] = (i == c) ? 1.0 : 0.0;
    for (int k = 0; k < n; k++) if (piv[k] != k) std::swap(b[k], b[piv[k]]);
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) b[i] -= A[(size_t)i * n + j] * b[j];
    for (int i = n - 1; i >= 0; i--) { for (int j = i + 1; j < n; j++) b[i] -= A[(size_t)i * n + j] * b[j]; b[i] /= A[(size_t)i * n + i]; }
    for (int i = 0; i < n; i++) inv[(size_t)i * n + c] = b[i];
  }
}

template <class T>
static void upload(DevBuf &b, const std::vector<T> &v, cudaStream_t st)
{
  b.reserve(std::max<size_t>(v.size() * sizeof(T), 16));
  if (!v.empty()) AKU_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

// The tensor-core scorer is the default throughput kernel (variant 0 or 3) for pools that are all diagonal or all
// full with at most 64 components per state; variants 1/2 force the FP32-pipe kernel.
// Device image of the clustering: the centres as a pool of one-component states for gmm_diag_f64, cluster ids, sizes.
void model_pack_clustering(akugpu_ctx *ctx)
{
  const HostModel &hm = ctx->hm;
  PackedF64 &p = ctx->p64;
  const int C = hm.n_clusters, D = hm.D;
  if (C <= 0) return;
  std::vector<double> prec((size_t)C * D), cst(C), w(C, 1.0);
  std::vector<int32_t> off(C + 1), idx(C), sz(C);
  for (int c = 0; c < C; c++) {
    double k = 1;
    for (int d = 0; d < D; d++) {
      const double cv = hm.c_cov[(size_t)c * D + d], pr = cv > 0 ? 1 / cv : 0;
      prec[(size_t)c * D + d] = pr;
      k *= pr;
    }
    if (k > 0) k = log(sqrt(k));
    cst[c] = k;
    off[c] = c; idx[c] = c; sz[c] = (int32_t)hm.cluster_gauss[c].size();
  }
  off[C] = C;
  upload(p.c_mean, hm.c_mean, ctx->stream);
  upload(p.c_prec, prec, ctx->stream);
  upload(p.c_cst, cst, ctx->stream);
  upload(p.c_mix_off, off, ctx->stream);
  upload(p.c_mix_gauss, idx, ctx->stream);
  upload(p.c_mix_w, w, ctx->stream);
  upload(p.g2c, hm.gauss_cluster, ctx->stream);
  upload(p.c_size, sz, ctx->stream);
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
}

static bool tc_wanted(akugpu_ctx *ctx)
{
  const HostModel &hm = ctx->hm;
  if (ctx->scorer_variant == 1 || ctx->scorer_variant == 2) return false;
  // diagonal pools go to the fp16x2 kernel (gmm_tc16.cu); the bf16x3 image is then packed only on demand
  if (ctx->scorer_variant != 4 && tc16_supported(hm)) return false;
  if (hm.n_full != 0 && hm.n_full != hm.G) return false;
  for (int s = 0; s < hm.S; s++)
    if (hm.mix_off[s + 1] - hm.mix_off[s] > 64) return false;
  return hm.S > 0 && hm.G > 0;
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
    if (hm.n_tr > 0) upload(p.g_tr, hm.g_tr, ctx->stream);
    // full-covariance Gaussians: exponential parameters (recompute_exponential_parameters, :1530-1547)
    p.n_full = hm.n_full;
    p.L = D * (D + 3) / 2;
    if (hm.n_full > 0) {
      const int nf = hm.n_full, L = p.L;
      std::vector<double> theta((size_t)L * nf, 0.0), fnorm(nf, 0.0), fcst(nf, 0.0);
      std::vector<int32_t> fg(nf), dg;
      std::vector<double> P, chol, Pm(D);
      for (int g = 0; g < G; g++) {
        const int fi = hm.full_index[g];
        if (fi < 0) { dg.push_back(g); continue; }
        fg[fi] = g;
        std::vector<double> cov(hm.full_cov.begin() + (size_t)fi * D * D, hm.full_cov.begin() + (size_t)(fi + 1) * D * D);
        if (!host_cholesky(cov, D, chol)) continue;   // not SPD: precision = 0, constant = 0 (set_covariance :1578-1582)
        host_lu_inverse(cov, D, P);
        if (!host_cholesky(P, D, chol)) continue;
        double det = 1;                           // spd_determinant (aku/LinearAlgebra.cc:26-38)
        for (int i = 0; i < D; i++) det *= chol[(size_t)i * D + i];
        det *= det;
        fcst[fi] = log(sqrt(det));
        const double *mu = &hm.mean[(size_t)g * D];
        double dot = 0;
        for (int i = 0; i < D; i++) {
          double s = 0;
          for (int j = 0; j < D; j++) s += P[(size_t)i * D + j] * mu[j];
          Pm[i] = s;
        }
        for (int i = 0; i < D; i++) dot += Pm[i] * mu[i];
        fnorm[fi] = -0.5 * dot;
        for (int i = 0; i < D; i++) theta[(size_t)i * nf + fi] = Pm[i];
        int pos = D;                              // map_m2v (aku/LinearAlgebra.cc:220-238): lower triangle, off-diagonals x sqrt(2)
        for (int i = 0; i < D; i++)
          for (int j = 0; j <= i; j++, pos++)
            theta[(size_t)pos * nf + fi] = -0.5 * ((i == j) ? P[(size_t)i * D + j] : sqrt(2.0) * P[(size_t)i * D + j]);
      }
      upload(p.theta, theta, ctx->stream);
      upload(p.full_norm, fnorm, ctx->stream);
      upload(p.full_cst, fcst, ctx->stream);
      upload(p.full_gauss, fg, ctx->stream);
      upload(p.diag_gauss, dg, ctx->stream);
    }
    ctx->have_p64 = true;
  }
  if (hm.n_full > 0) {   // the fp32 image covers diagonal pools only; full-covariance models score in double
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->p32.n_tiles = 0;
    ctx->ptc.ready = false;
    ctx->ptc16.ready = false;
    if (tc_wanted(ctx)) model_pack_tc(ctx);
    else if (ctx->scorer_variant == 0 || ctx->scorer_variant == 3) {
      model_pack_tc16(ctx);              // all-full pools: the streaming fp16x2 kernel
      if (!ctx->ptc16.ready && hm.n_full == hm.G && ctx->ptc16.q_max <= TC_Q_MAX) {
        bool ok = true;
        for (int s = 0; s < hm.S && ok; s++) ok = hm.mix_off[s + 1] - hm.mix_off[s] <= 64;
        if (ok) model_pack_tc(ctx);
      }
    }
    ctx->have_model = true;
    return;
  }
  // ---- tensor-core images first: the fp16x2 packer decides which states (if any) it leaves to the FP32-pipe kernel ----
  ctx->ptc.ready = false;
  ctx->ptc16.ready = false;
  ctx->ptc16.hybrid = false;
  if (tc_wanted(ctx)) model_pack_tc(ctx);
  else if (ctx->scorer_variant == 0 || ctx->scorer_variant == 3) model_pack_tc16(ctx);
  // packing can decline (a component constant outside the fp16 range): fall back to the bf16x3 image
  if (!ctx->ptc16.ready && !ctx->ptc.ready && (ctx->scorer_variant == 0 || ctx->scorer_variant == 3) &&
      (ctx->ptc16.q_max <= TC_Q_MAX || ctx->scorer_variant == 3)) {
    const HostModel &h = ctx->hm;
    bool ok = h.S > 0 && h.G > 0;
    for (int s = 0; s < h.S && ok; s++) ok = h.mix_off[s + 1] - h.mix_off[s] <= 64;
    if (ok) model_pack_tc(ctx);
  }
  const bool only_bad = ctx->ptc16.ready && ctx->ptc16.hybrid;       // FP32 image = the states the tensor-core image skips
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
    if (only_bad && !ctx->ptc16.bad_state[s]) continue;
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
