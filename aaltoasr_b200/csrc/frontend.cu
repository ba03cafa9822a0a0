// frontend.cu -- B200 feature front-end: config parser, module graph and kernels.
//
// Mirrors the reference's FeatureGenerator / FeatureModule plugin chain, but computes
// whole utterances (all frames of a batch per launch) instead of one frame per call:
//   config text  FeatureGenerator::load_configuration  aku/FeatureGenerator.cc:97-219
//                ModuleConfig::read                    aku/ModuleConfig.cc:166-203
//   audiofile    AudioFileModule                       aku/FeatureModules.cc:328-440
//   fft          FFTModule                             aku/FeatureModules.cc:476-566
//   mel          MelModule                             aku/FeatureModules.cc:776-849
//   power        PowerModule                           aku/FeatureModules.cc:875-885
//   mel_power    MelPowerModule                        aku/FeatureModules.cc:908-919
//   dct          DCTModule                             aku/FeatureModules.cc:938-979
//   delta        DeltaModule                           aku/FeatureModules.cc:999-1037
//   normalization NormalizationModule                  aku/FeatureModules.cc:1057-1142
//   lin_transform LinTransformModule                   aku/FeatureModules.cc:1167-1269
//   merge        MergerModule                          aku/FeatureModules.cc:1340-1364
//   mean_subtractor MeanSubtractorModule               aku/FeatureModules.cc:1385-1454
//   concat       ConcatModule                          aku/FeatureModules.cc:1473-1501
//
// The reference mixes float and double on purpose-by-accident; every kernel below
// keeps the same type at the same place (float pre-emphasis, float Hamming table made
// with cosf, float FFT, sqrtf/logf, float mel accumulators, cosf DCT basis with double
// accumulation, float power accumulator, double deltas ...) and blocks FMA contraction
// where the x86 build has none, so that everything except the FFT's internal rounding
// order is bit-for-bit the reference's arithmetic.
//
// Layout: every module output is a row-major double matrix [rows][dim] in HBM.  A row is
// one (utterance, frame) pair; each utterance carries a halo of H frames on both sides
// (H = largest context any module needs) so that context modules never special-case
// borders: only the base module clamps the frame index (first/last window replicated,
// aku/FeatureModules.cc:381-397), exactly like the reference.
#include "ctx.hpp"
#include "kernels.hpp"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <sstream>

namespace akugpu {

// ===================================================================================
// config parsing
namespace {

struct Block {
  std::vector<std::pair<std::string, std::string>> kv;
  const std::string *find(const std::string &k) const {
    for (auto &p : kv) if (p.first == k) return &p.second;
    return nullptr;
  }
};

std::string trim(const std::string &s) {
  size_t b = 0, e = s.size();
  while (e > b && (s[e - 1] == ' ' || s[e - 1] == '\t' || s[e - 1] == '\r' || s[e - 1] == '\n')) e--;
  while (b < e && (s[b] == ' ' || s[b] == '\t')) b++;
  return s.substr(b, e - b);
}
std::vector<std::string> split_ws(const std::string &s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
    size_t b = i;
    while (i < s.size() && s[i] != ' ' && s[i] != '\t') i++;
    if (i > b) out.push_back(s.substr(b, i - b));
  }
  return out;
}
long to_long(const std::string &s) {
  char *e;
  long v = strtol(s.c_str(), &e, 10);
  if (s.empty() || *e) throw Error(AKUGPU_E_CONFIG, "invalid integer value: " + s);
  return v;
}
float to_float(const std::string &s) {   // aku/str.cc:261-282: strtod narrowed to float
  char *e;
  float v = (float)strtod(s.c_str(), &e);
  if (s.empty() || *e) throw Error(AKUGPU_E_CONFIG, "invalid float value: " + s);
  return v;
}
bool get_int(const Block &b, const char *k, int &v) { auto p = b.find(k); if (!p) return false; v = (int)to_long(*p); return true; }
bool get_float(const Block &b, const char *k, float &v) { auto p = b.find(k); if (!p) return false; v = to_float(*p); return true; }
bool get_fvec(const Block &b, const char *k, std::vector<float> &v) {
  auto p = b.find(k);
  if (!p) return false;
  auto f = split_ws(*p);
  v.resize(f.size());
  for (size_t i = 0; i < f.size(); i++) v[i] = to_float(f[i]);
  return true;
}

void parse_blocks(const std::string &text, std::vector<Block> &blocks, bool bare)
{
  std::istringstream in(text);
  std::string line;
  int lineno = 0;
  if (bare) {   // key/value lines only (set_parameters)
    Block b;
    while (std::getline(in, line)) {
      line = trim(line);
      if (line.empty()) continue;
      size_t sp = line.find_first_of(" \t");
      if (sp == std::string::npos) throw Error(AKUGPU_E_CONFIG, "value missing for option: " + line);
      b.kv.push_back(std::make_pair(line.substr(0, sp), trim(line.substr(sp))));
    }
    blocks.push_back(b);
    return;
  }
  while (std::getline(in, line)) {
    lineno++;
    line = trim(line);
    if (line.empty()) continue;
    if (line != "module") throw Error(AKUGPU_E_CONFIG, fmt("expected keyword 'module' on line %d: ", lineno) + line);
    Block b;
    bool first = true, closed = false;
    while (std::getline(in, line)) {
      lineno++;
      line = trim(line);
      if (line.empty()) continue;
      if (first) {
        if (line != "{") throw Error(AKUGPU_E_CONFIG, "'{' expected in module config file: " + line);
        first = false;
        continue;
      }
      if (line == "}") { closed = true; break; }
      size_t sp = line.find_first_of(" \t");
      if (sp == std::string::npos) throw Error(AKUGPU_E_CONFIG, "value missing for option: " + line);
      std::string key = line.substr(0, sp);
      if (b.find(key)) throw Error(AKUGPU_E_CONFIG, "value redefined: " + line);
      b.kv.push_back(std::make_pair(key, trim(line.substr(sp))));
    }
    if (!closed) throw Error(AKUGPU_E_CONFIG, "unexpected end of module config file");
    blocks.push_back(b);
  }
}

template <class T>
std::shared_ptr<DevBuf> upload_vec(const std::vector<T> &v, cudaStream_t st)
{
  auto b = std::make_shared<DevBuf>();
  b->reserve(std::max<size_t>(16, v.size() * sizeof(T)));
  if (!v.empty()) {
    AKU_CUDA(cudaMemcpyAsync(b->p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    AKU_CUDA(cudaStreamSynchronize(st));
  }
  return b;
}

void check_lin_transform(Module &m, int src_dim)
{
  // LinTransformModule::check_transform_parameters, aku/FeatureModules.cc:1203-1241
  if (m.matrix.empty()) {
    m.matrix_defined = false;
  } else {
    m.matrix_defined = true;
    if ((int)m.matrix.size() != m.dim * src_dim) throw Error(AKUGPU_E_CONFIG, "LinTransformModule: Invalid matrix dimension");
  }
  if (m.bias.empty()) {
    m.bias_defined = false;
  } else {
    m.bias_defined = true;
    if ((int)m.bias.size() != m.dim) throw Error(AKUGPU_E_CONFIG, "LinTransformModule: Invalid bias dimension");
  }
}

void read_normalization(Module &m, const Block &b)
{
  // NormalizationModule::set_module_config / set_parameters, aku/FeatureModules.cc:1057-1110
  get_fvec(b, "mean", m.v_mean);
  if ((int)m.v_mean.size() != m.dim) throw Error(AKUGPU_E_CONFIG, "NormalizationModule: Invalid mean dimension");
  if (b.find("var") && b.find("scale"))
    throw Error(AKUGPU_E_CONFIG, "NormalizationModule: Both scale and var can not be defined simultaneously");
  if (get_fvec(b, "var", m.v_scale)) {
    if ((int)m.v_scale.size() != m.dim) throw Error(AKUGPU_E_CONFIG, "Normalization module: Invalid variance dimension");
    for (int i = 0; i < m.dim; i++) m.v_scale[i] = 1 / sqrtf(m.v_scale[i]);
  } else if (get_fvec(b, "scale", m.v_scale)) {
    if ((int)m.v_scale.size() != m.dim) throw Error(AKUGPU_E_CONFIG, "NormalizationModule: Invalid scale dimension");
  }
}

void upload_module_params(akugpu_ctx *ctx, Module &m, const std::vector<Module> &mods)
{
  cudaStream_t st = ctx->stream;
  switch (m.type) {
    case M_FFT: {
      const int N = mods[m.src[0]].dim;
      std::vector<float> win(N);
      for (int i = 0; i < N; i++) win[i] = .54 - .46 * cosf(2 * M_PI * i / (N - 1.0));   // aku/FeatureModules.cc:488-490
      m.d_a = upload_vec(win, st);
      std::vector<float> tw(2 * (size_t)N);   // e^{-2 pi i k / N}
      for (int k = 0; k < N; k++) {
        double a = -2.0 * M_PI * k / N;
        tw[2 * k] = (float)cos(a);
        tw[2 * k + 1] = (float)sin(a);
      }
      m.d_b = upload_vec(tw, st);
      break;
    }
    case M_MEL: {
      // Triangle weights and their sums, tap by tap in the reference's float arithmetic (MelModule::generate,
      // aku/FeatureModules.cc:809-843: scale = (t - beg) / (end - beg) rising, (end - t) / (end - beg) falling; the
      // host's SSE float operations round exactly like the device's __f*_rn), so that the kernels do not divide.
      // d_a: float scales in tap order; d_b: int {first spectrum bin, taps, offset into d_a} per mel bin; d_c: float sums.
      std::vector<float> scales, sums;
      std::vector<int> desc;
      for (int b = 0; b < m.dim; b++) {
        float sum = 0, scale;
        float beg = m.bin_edges[b] - 1.f;
        float end = m.bin_edges[b + 1];
        int t = (int)fmaxf(ceilf(beg), 0.0f);
        const int t0 = t, off = (int)scales.size();
        while (t < end) {
          scale = ((float)t - beg) / (end - beg);
          scales.push_back(scale);
          sum = sum + scale;
          t++;
        }
        beg = end;
        end = m.bin_edges[b + 2];
        while (t < end) {
          scale = (end - (float)t) / (end - beg);
          scales.push_back(scale);
          sum = sum + scale;
          t++;
        }
        desc.push_back(t0); desc.push_back(t - t0); desc.push_back(off);
        sums.push_back(sum);
      }
      if (scales.empty()) scales.push_back(0.f);
      m.d_a = upload_vec(scales, st);
      m.d_b = upload_vec(desc, st);
      m.d_c = upload_vec(sums, st);
      m.d_d = upload_vec(std::vector<double>(scales.begin(), scales.end()), st);
      break;
    }
    case M_DCT: {
      const int sd = mods[m.src[0]].dim;
      const int bias = m.zeroth ? 1 : 0;
      std::vector<float> tab((size_t)m.dim * sd, 1.0f);   // row 0 of a zeroth-component DCT is a plain sum
      for (int i = 0; i < m.dim - bias; i++)
        for (int b = 0; b < sd; b++) tab[(size_t)(i + bias) * sd + b] = cosf((i + 1) * (b + 0.5) * M_PI / sd);   // :977
      m.d_a = upload_vec(tab, st);
      m.d_d = upload_vec(std::vector<double>(tab.begin(), tab.end()), st);
      break;
    }
    case M_VTLN: {
      // Warped bin positions (float), then the interpolation table, in the reference's mixed float/double arithmetic:
      // create_blin_bins :1662-1675, create_pwlin_bins :1634-1660, create_slapt_bins :1677-1695, create_sinc_coef_table
      // :1697-1723 (util::sinc aku/util.hh:151-159).  d_a: float coefficients in tap order; d_b: int {first source bin,
      // taps, offset} per bin; d_c: float warped positions (linear interpolation when sinc_interpolation_rad == 0).
      const int dim = m.dim;
      std::vector<float> &bins = m.vtln_bins;
      bins.assign(dim, 0.f);
      int t;
      if (m.use_slapt) {
        for (t = 0; t < dim - 1; t++) {
          double nf = M_PI * (double)t / (dim - 1);
          bins[t] = t;
          for (int i = 0; i < (int)m.slapt_params.size(); i++) bins[t] += m.slapt_params[i] * sin((i + 1) * nf) * (dim - 1);
        }
        bins[t] = dim - 1;
      } else if (m.use_pwlin) {
        float border, slope = 0, point = 0;
        bool limit = false;
        border = m.pwlin_turn_point * (float)(dim - 1);
        for (t = 0; t < dim - 1; t++) {
          if (!limit) bins[t] = m.warp_factor * (float)t;
          else bins[t] = slope * (float)t + point;
          if (!limit && (t >= border || bins[t] >= border)) {
            slope = ((float)dim - 1 - bins[t]) / ((float)dim - 1 - t);
            point = (1 - slope) * (float)(dim - 1);
            limit = true;
          }
        }
        bins[t] = (float)(dim - 1);
      } else {
        for (t = 0; t < dim - 1; t++) {
          double nf = M_PI * (double)t / (dim - 1);
          bins[t] = t + 2 * atan2((m.warp_factor - 1) * sin(nf), 1 + (1 - m.warp_factor) * cos(nf)) / M_PI * (dim - 1);
        }
        bins[t] = dim - 1;
      }
      std::vector<float> coef;
      std::vector<int> desc;
      if (m.all_pass) {
        // All-pass VTLN: the warp as a matrix T on the cepstrum of the spectrum, final = IDCT * (T * DCT), applied as full
        // rows of coefficients (create_all_pass_blin_transform :1717-1757, create_all_pass_slapt_transform :1759-1868,
        // set_all_pass_transform :1870-1904).  Everything in double; sums over the inner index in ascending order.
        auto conv = [](const std::vector<double> &a, const std::vector<double> &bb) {      // (a * b)[j] = sum_{p+q=j} a[p] b[q]
          std::vector<double> out(a.size() + bb.size() - 1, 0.0);
          for (size_t j = 0; j < out.size(); j++) {
            const size_t lo = j >= bb.size() - 1 ? j - (bb.size() - 1) : 0, hi = std::min(j, a.size() - 1);
            double tt = 0;
            for (size_t p = lo; p <= hi; p++) tt += a[p] * bb[j - p];
            out[j] = tt;
          }
          return out;
        };
        std::vector<double> T((size_t)dim * dim, 0.0);
        T[0] = 1.0;
        if (!m.use_slapt) {
          const double alpha = (double)(float)(m.warp_factor - 1);
          std::vector<double> q1(dim), q(dim, 0.0), qn(dim);
          q1[0] = -alpha;
          double temp = 1 - alpha * alpha;
          for (int i = 1; i < dim; i++) { q1[i] = temp; temp *= alpha; }
          q[0] = 1;
          for (int i = 1; i < dim; i++) {      // column i: the i-th convolution power of q1, truncated to dim terms
            for (int j = 0; j < dim; j++) {
              double tt = 0;
              for (int k = 0; k <= j; k++) tt += q[k] * q1[j - k];
              qn[j] = tt;
            }
            q = qn;
            T[(size_t)0 * dim + i] = 2 * q[0];
            for (int j = 1; j < dim; j++) T[(size_t)j * dim + i] = q[j];
          }
        } else {
          const int P = (int)m.slapt_params.size();
          std::vector<double> f1(2 * P + 1, 0.0);
          for (int i = 0; i < P; i++) {
            f1[i] = -m.slapt_params[P - i - 1] * M_PI / 2;
            f1[i + P + 1] = m.slapt_params[i] * M_PI / 2;
          }
          std::vector<double> q(2 * dim + 1, 0.0), cur_f(1, 1.0);
          int center = 0;
          double cur_m = 1;
          for (int i = 0; i <= 10; i++) {        // exp of the sequence: sum over i of f1^(*i) / i!
            if (i > 0) cur_m = cur_m / (double)i;
            for (int j = std::max(0, dim - center); j < std::min(2 * dim + 1, dim + center + 1); j++) q[j] = q[j] + cur_m * cur_f[j - dim + center];
            cur_f = conv(cur_f, f1);
            center = ((int)cur_f.size() - 1) / 2;
          }
          q.pop_back(); q.pop_back();
          const std::vector<double> q1 = q;
          for (int i = 1; i < dim; i++) {
            T[(size_t)0 * dim + i] = 2 * q[dim - 1];
            for (int j = 1; j < dim; j++) T[(size_t)j * dim + i] = q[dim + j - 1] + q[dim - j - 1];
            const std::vector<double> full = conv(q, q1);
            q.assign(full.begin() + (dim - 1), full.begin() + (3 * dim - 2));
          }
        }
        std::vector<double> dct((size_t)dim * dim), idct((size_t)dim * dim), tmp((size_t)dim * dim), fin((size_t)dim * dim);
        for (int i = 0; i < dim; i++)
          for (int j = 0; j < dim; j++) {
            dct[(size_t)i * dim + j] = cos(i * (j + 0.5) * M_PI / dim);
            idct[(size_t)i * dim + j] = j == 0 ? 1.0 / dim : cos((i + 0.5) * j * M_PI / dim) * 2 / dim;
          }
        for (int i = 0; i < dim; i++)
          for (int j = 0; j < dim; j++) {
            double acc = 0;
            for (int p2 = 0; p2 < dim; p2++) acc += T[(size_t)i * dim + p2] * dct[(size_t)p2 * dim + j];
            tmp[(size_t)i * dim + j] = acc;
          }
        for (int i = 0; i < dim; i++)
          for (int j = 0; j < dim; j++) {
            double acc = 0;
            for (int p2 = 0; p2 < dim; p2++) acc += idct[(size_t)i * dim + p2] * tmp[(size_t)p2 * dim + j];
            fin[(size_t)i * dim + j] = acc;
          }
        for (int bi = 0; bi < dim; bi++) {
          desc.push_back(0); desc.push_back(dim); desc.push_back((int)coef.size());
          for (int j = 0; j < dim; j++) coef.push_back((float)fin[(size_t)bi * dim + j]);
        }
      } else if (m.sinc_rad > 0) {
        auto sinc = [](float x) -> float {
          const double PI = 3.14159265358979323846;
          if (fabs(x) < 1e-8) return 1;
          double y = PI * x;
          return sin(y) / y;
        };
        for (int bi = 0; bi < dim; bi++) {
          int cent = (int)(bins[bi] + 0.5);
          int min_i = std::max(cent - m.sinc_rad, 0);
          int max_i = std::min(cent + m.sinc_rad + 1, dim);
          desc.push_back(min_i); desc.push_back(std::max(0, max_i - min_i)); desc.push_back((int)coef.size());
          for (int i = min_i; i < max_i; i++) {
            float tt = sinc(i - bins[bi]);
            if (m.lanczos) {
              if (fabs(i - bins[bi]) < m.sinc_rad) tt *= sinc((i - bins[bi]) / (float)m.sinc_rad);
              else tt = 0;
            }
            coef.push_back(tt);
          }
        }
      }
      if (coef.empty()) coef.push_back(0.f);
      if (desc.empty()) desc.assign(3, 0);
      m.d_a = upload_vec(coef, st);
      m.d_b = upload_vec(desc, st);
      m.d_c = upload_vec(bins, st);
      break;
    }
    case M_SR_NORM: {
      // SRNormModule::set_speech_rate (aku/FeatureModules.cc:2004-2036): Lanczos taps over the input frames for every
      // output frame, float arithmetic as the reference.  d_a: float taps; d_b: int {first input frame, taps, offset}.
      const float in_cent = (float)(m.in_frames - 1) / 2, out_cent = (float)(m.out_frames - 1) / 2;
      auto sinc = [](float x) -> float {
        const double PI = 3.14159265358979323846;
        if (fabs(x) < 1e-8) return 1;
        double y = PI * x;
        return sin(y) / y;
      };
      std::vector<float> coef;
      std::vector<int> desc;
      for (int i = 0; i < m.out_frames; i++) {
        float target_pos = (i - out_cent) / m.speech_rate + in_cent;
        int cent = (int)roundf(target_pos);
        int interp_start = std::max(cent - m.lanczos_order, 0);
        int interp_end = std::min(cent + m.lanczos_order + 1, m.in_frames);
        desc.push_back(interp_start); desc.push_back(std::max(0, interp_end - interp_start)); desc.push_back((int)coef.size());
        for (int j = interp_start; j < interp_end; j++) {
          float t = sinc(j - target_pos);
          if (fabs(j - target_pos) < m.lanczos_order) t *= sinc((j - target_pos) / (float)m.lanczos_order);
          else t = 0;
          coef.push_back(t);
        }
      }
      if (coef.empty()) coef.push_back(0.f);
      m.d_a = upload_vec(coef, st);
      m.d_b = upload_vec(desc, st);
      break;
    }
    case M_QUANTEQ: {
      // active only when alpha, gamma and quant_max are all set (QuantEqModule::generate :2127-2133); d_a = the three
      // vectors back to back
      std::vector<float> all;
      if (!m.q_alpha.empty() && !m.q_gamma.empty() && !m.q_max.empty()) {
        if ((int)m.q_alpha.size() < m.dim || (int)m.q_gamma.size() < m.dim || (int)m.q_max.size() < m.dim)
          throw Error(AKUGPU_E_CONFIG, "QuantEqModule: alpha / gamma / quant_max need one value per channel");
        all.insert(all.end(), m.q_alpha.begin(), m.q_alpha.begin() + m.dim);
        all.insert(all.end(), m.q_gamma.begin(), m.q_gamma.begin() + m.dim);
        all.insert(all.end(), m.q_max.begin(), m.q_max.begin() + m.dim);
      }
      if (all.empty()) all.push_back(0.f);
      m.d_a = upload_vec(all, st);
      break;
    }
    case M_NORMALIZATION:
      m.d_a = upload_vec(m.v_mean, st);
      m.d_b = upload_vec(m.v_scale, st);
      break;
    case M_LIN_TRANSFORM:
      m.d_a = upload_vec(m.matrix, st);
      m.d_b = upload_vec(m.bias, st);
      break;
    default: break;
  }
}

}  // namespace

void frontend_parse(akugpu_ctx *ctx, const std::string &text)
{
  std::vector<Block> blocks;
  parse_blocks(text, blocks, false);
  if (blocks.empty()) throw Error(AKUGPU_E_CONFIG, "feature configuration defines no modules");
  Frontend fe;
  std::map<std::string, int> by_name;
  for (size_t bi = 0; bi < blocks.size(); bi++) {
    const Block &b = blocks[bi];
    const std::string *type = b.find("type"), *name = b.find("name");
    if (!type) throw Error(AKUGPU_E_CONFIG, fmt("type not defined for module %d", (int)bi));
    if (!name) throw Error(AKUGPU_E_CONFIG, fmt("name not defined for module %d", (int)bi));
    if (name->find_first_of(" \t\n") != std::string::npos) throw Error(AKUGPU_E_CONFIG, "module name may not contain whitespaces");
    Module m;
    m.name = *name;
    static const struct { const char *s; ModType t; } types[] = {
        {"audiofile", M_AUDIOFILE}, {"fft", M_FFT}, {"mel", M_MEL}, {"power", M_POWER}, {"mel_power", M_MEL_POWER},
        {"dct", M_DCT}, {"delta", M_DELTA}, {"merge", M_MERGE}, {"concat", M_CONCAT},
        {"normalization", M_NORMALIZATION}, {"lin_transform", M_LIN_TRANSFORM}, {"mean_subtractor", M_MEAN_SUBTRACTOR},
        {"pre", M_PRE}, {"vtln", M_VTLN}, {"sr_norm", M_SR_NORM}, {"quanteq", M_QUANTEQ}};
    bool known = false;
    for (auto &t : types) if (*type == t.s) { m.type = t.t; known = true; }
    if (!known) throw Error(AKUGPU_E_CONFIG, "Unknown module type '" + *type + "'");
    if (fe.mods.empty() && m.type != M_AUDIOFILE && m.type != M_PRE) throw Error(AKUGPU_E_CONFIG, "first module should be a base module");
    if (!fe.mods.empty() && (m.type == M_AUDIOFILE || m.type == M_PRE)) throw Error(AKUGPU_E_CONFIG, "base module must be the first module: " + m.name);
    if (by_name.count(m.name)) throw Error(AKUGPU_E_CONFIG, "multiple definitions of module name: " + m.name);
    const std::string *sources = b.find("sources");
    if (fe.mods.empty() && sources) throw Error(AKUGPU_E_CONFIG, "can not define sources for the first module");
    if (!fe.mods.empty() && !sources) throw Error(AKUGPU_E_CONFIG, "sources not defined for module: " + m.name);
    if (sources) {
      for (auto &s : split_ws(*sources)) {
        auto it = by_name.find(s);
        if (it == by_name.end()) throw Error(AKUGPU_E_CONFIG, "unknown source module: " + s);
        m.src.push_back(it->second);
      }
      if (m.src.empty()) throw Error(AKUGPU_E_CONFIG, "sources not defined for module: " + m.name);
      if (m.type != M_MERGE && m.src.size() != 1)   // FeatureModule::add_source, aku/FeatureModules.cc:160-166
        throw Error(AKUGPU_E_CONFIG, "multiple sources are not allowed for module: " + m.name);
    }
    const int sdim = m.src.empty() ? 0 : fe.mods[m.src.back()].dim;
    switch (m.type) {
      case M_AUDIOFILE: {
        if (!get_int(b, "sample_rate", m.sample_rate)) throw Error(AKUGPU_E_CONFIG, "AudioFileModule: Must set sample rate");
        m.emph = 0.97; get_float(b, "pre_emph_coef", m.emph);
        m.frame_rate = 125; get_float(b, "frame_rate", m.frame_rate);
        m.window_advance = m.sample_rate / m.frame_rate;
        m.window_width = (int)(2 * m.sample_rate / m.frame_rate);
        get_int(b, "window_width", m.window_width);
        m.dim = m.window_width;
        int raw = 0; get_int(b, "raw", raw);   // container format is the host reader's business
        m.copy_borders = 1; get_int(b, "copy_borders", m.copy_borders);
        if (m.window_width < 2 || m.window_advance <= 0) throw Error(AKUGPU_E_CONFIG, "AudioFileModule: bad window geometry");
        break;
      }
      case M_PRE: {       // PreModule::set_module_config, aku/FeatureModules.cc:671-689: features read from a file
        m.frame_rate = 125; m.sample_rate = 16000;
        get_int(b, "sample_rate", m.sample_rate);
        get_float(b, "frame_rate", m.frame_rate);
        m.legacy_file = 0; get_int(b, "legacy_file", m.legacy_file);   // 1-byte header (akugpu_frontend_pre_legacy): the host reader's business
        if (!get_int(b, "dim", m.dim)) throw Error(AKUGPU_E_CONFIG, "PreModule: Must set dimension");
        if (m.dim < 1) throw Error(AKUGPU_E_CONFIG, "PreModule: Must set dimension");
        m.window_advance = m.sample_rate / m.frame_rate;
        break;
      }
      case M_FFT:
        if (fe.mods[m.src[0]].type != M_AUDIOFILE)
          throw Error(AKUGPU_E_CONFIG, "fft module must read the audiofile module directly");
        m.magnitude = 1; get_int(b, "magnitude", m.magnitude);
        m.log = 0; get_int(b, "log", m.log);
        m.dim = sdim / 2 + 1;
        break;
      case M_MEL: {
        m.root = 0; get_int(b, "root", m.root);
        const int sr = fe.mods[0].sample_rate;
        m.dim = (int)((21 + 2) * log10f(1 + sr / 1400.0) / log10f(1 + 16000 / 1400.0) - 2);   // :784-785
        int edges = m.dim + 2;
        float rate = sr;
        float mel_step = 2595 * log10f(1.0 + rate / 1400.0) / edges;
        m.bin_edges.resize(edges);
        for (int i = 0; i < edges; i++)
          m.bin_edges[i] = 1400.0 * (pow(10, (i + 1) * mel_step / 2595) - 1) * (sdim - 1) / rate;   // :799-801
        break;
      }
      case M_POWER: case M_MEL_POWER: m.dim = 1; break;
      case M_DCT:
        m.dim = 12; get_int(b, "dim", m.dim);
        if (m.dim < 1) throw Error(AKUGPU_E_CONFIG, "DCTModule: Dimension must be > 0");
        m.zeroth = 0; get_int(b, "zeroth", m.zeroth);
        break;
      case M_DELTA:
        m.dim = sdim;
        m.width = 2; get_int(b, "width", m.width);
        m.norm = 2 * m.width * (m.width + 1) * (2 * m.width + 1) / 6;   // int arithmetic, then float (:1007)
        get_float(b, "normalization", m.norm);
        if (m.width < 1) throw Error(AKUGPU_E_CONFIG, "DeltaModule: Delta width must be > 0");
        m.left = m.right = m.width;
        break;
      case M_MERGE:
        m.dim = 0;
        for (int s : m.src) m.dim += fe.mods[s].dim;
        break;
      case M_CONCAT:
        get_int(b, "left", m.left); get_int(b, "right", m.right);
        if (m.left < 0 || m.right < 0) throw Error(AKUGPU_E_CONFIG, "ConcatModule: context spans must be >= 0");
        m.dim = sdim * (1 + m.left + m.right);
        break;
      case M_NORMALIZATION:
        m.dim = sdim;
        m.v_mean.assign(m.dim, 0.f);
        m.v_scale.assign(m.dim, 1.f);
        read_normalization(m, b);
        break;
      case M_LIN_TRANSFORM:
        m.dim = sdim;
        get_fvec(b, "matrix", m.matrix); get_fvec(b, "bias", m.bias);
        get_int(b, "dim", m.dim);
        if (m.dim < 1) throw Error(AKUGPU_E_CONFIG, "LinTransformModule: Dimension must be > 0");
        check_lin_transform(m, sdim);
        break;
      case M_VTLN: {      // VtlnModule::set_module_config, aku/FeatureModules.cc:1530-1573
        m.dim = sdim;
        m.use_pwlin = 0; m.pwlin_turn_point = 0.8f;
        get_int(b, "pwlin_vtln", m.use_pwlin); get_float(b, "pwlin_turnpoint", m.pwlin_turn_point);
        m.use_slapt = 0; get_int(b, "slapt", m.use_slapt);
        if (m.use_pwlin && m.use_slapt) throw Error(AKUGPU_E_CONFIG, "VtlnModule: Can not use both pwlin_vtln and slapt!");
        m.sinc_rad = 8; get_int(b, "sinc_interpolation_rad", m.sinc_rad);
        m.all_pass = 0; get_int(b, "all-pass", m.all_pass);
        if (m.use_pwlin && m.all_pass) throw Error(AKUGPU_E_CONFIG, "VtlnModule: Can not use both pwlin_vtln and all-pass!");
        m.lanczos = m.all_pass ? 0 : 1; get_int(b, "lanczos_window", m.lanczos);
        m.lanczos = m.lanczos > 0 ? 1 : 0;
        if (m.lanczos && m.all_pass) throw Error(AKUGPU_E_CONFIG, "VtlnModule: Can not use both lanczos_window and all-pass!");
        if (m.all_pass && m.sinc_rad <= 0)     // the reference would then interpolate with warped bins it never computed
          throw Error(AKUGPU_E_CONFIG, "VtlnModule: all-pass needs sinc_interpolation_rad > 0");
        m.warp_factor = 1.0f;
        m.slapt_params.assign(1, 0.0f);
        break;
      }
      case M_SR_NORM: {   // SRNormModule::set_module_config, aku/FeatureModules.cc:1954-1989
        m.in_frames = 0; m.out_frames = 0;
        get_int(b, "in_frames", m.in_frames); get_int(b, "out_frames", m.out_frames);
        if (m.in_frames == 0 || m.out_frames == 0) throw Error(AKUGPU_E_CONFIG, "SRNormModule: Must set both in_frames and out_frames.");
        if (m.in_frames < 0 || m.out_frames < 0 || sdim % m.in_frames != 0)
          throw Error(AKUGPU_E_CONFIG, "SRNormModule: in_frames does not match with the input dimension");
        m.frame_dim = sdim / m.in_frames;
        m.dim = m.out_frames * m.frame_dim;
        m.lanczos_order = 4; get_int(b, "lanczos_order", m.lanczos_order);
        if (m.lanczos_order < 1) throw Error(AKUGPU_E_CONFIG, "SRNormModule: lanczos_order must be positive.");
        m.speech_rate = 1.0f; get_float(b, "speech_rate", m.speech_rate);
        break;
      }
      case M_QUANTEQ:     // QuantEqModule::set_module_config :2086-2093 (quant_train only matters to the trainer)
        m.dim = sdim;
        m.q_alpha.clear(); m.q_gamma.clear(); m.q_max.clear();
        break;
      case M_MEAN_SUBTRACTOR: {
        m.dim = sdim;
        int l = 75, r = 75;
        get_int(b, "left", l); get_int(b, "right", r);
        if (l + 1 < 1 || r + 1 < 1) throw Error(AKUGPU_E_CONFIG, "MeanSubtractorModule: context widths must be >= 0");
        m.left = l; m.right = r;          // frames actually averaged: [t-l, t+r]
        m.ms_width = l + r + 1;
        break;
      }
    }
    upload_module_params(ctx, m, fe.mods);
    by_name[m.name] = (int)fe.mods.size();
    fe.mods.push_back(m);
  }
  fe.last = (int)fe.mods.size() - 1;
  fe.configured = true;
  ctx->fe = fe;
}

void frontend_set_parameters(akugpu_ctx *ctx, const std::string &module, const std::string &text)
{
  Frontend &fe = ctx->fe;
  int idx = -1;
  for (size_t i = 0; i < fe.mods.size(); i++) if (fe.mods[i].name == module) idx = (int)i;
  if (idx < 0) throw Error(AKUGPU_E_ARG, "unknown module requested: " + module);
  Module &m = fe.mods[idx];
  std::vector<Block> blocks;
  parse_blocks(text, blocks, true);
  const Block &b = blocks[0];
  if (m.type == M_NORMALIZATION) {
    read_normalization(m, b);
  } else if (m.type == M_SR_NORM) {     // SRNormModule::set_parameters :1991-1996
    m.speech_rate = 1.0f;
    get_float(b, "speech_rate", m.speech_rate);
  } else if (m.type == M_QUANTEQ) {     // QuantEqModule::set_parameters :2095-2104
    m.q_alpha.clear(); m.q_gamma.clear(); m.q_max.clear();
    get_fvec(b, "alpha", m.q_alpha); get_fvec(b, "gamma", m.q_gamma); get_fvec(b, "quant_max", m.q_max);
  } else if (m.type == M_VTLN) {        // VtlnModule::set_parameters, aku/FeatureModules.cc:1575-1592
    if (m.use_slapt) {
      m.slapt_params.assign(1, 0.0f);
      get_fvec(b, "slapt_coef", m.slapt_params);
    } else {
      m.warp_factor = 1.0f;
      get_float(b, "warp_factor", m.warp_factor);
    }
  } else if (m.type == M_LIN_TRANSFORM) {
    m.matrix.clear(); m.bias.clear();
    get_fvec(b, "matrix", m.matrix); get_fvec(b, "bias", m.bias);
    check_lin_transform(m, fe.mods[m.src[0]].dim);
  } else {
    return;   // FeatureModule::set_parameters default is a no-op (aku/FeatureModule.hh:107)
  }
  upload_module_params(ctx, m, fe.mods);
}

int64_t frontend_num_frames(const Frontend &fe, int64_t n_samples)
{
  // Sequential generate(0),generate(1),.. reports eof on the first frame whose window
  // [ws, ws+W+1) runs past the file (aku/FeatureModules.cc:399-404), ws = (int)(f*advance).
  const Module &a = fe.mods[0];
  if (a.type == M_PRE) return n_samples;        // a `pre` base module: one frame per stored row (PreModule::last_frame :651-662)
  if (n_samples < a.window_width + 1) return 0;
  int64_t f = (int64_t)((float)(n_samples - a.window_width - 1) / a.window_advance);
  while (f > 0 && (int64_t)(int)((int)f * a.window_advance) + a.window_width + 1 > n_samples) f--;
  while ((int64_t)(int)((int)(f + 1) * a.window_advance) + a.window_width + 1 <= n_samples) f++;
  return f + 1;
}

// ===================================================================================
// kernels
struct UttDesc {
  int64_t pcm_off;     // first sample of the utterance in the PCM buffer
  int64_t n_samples;
  int64_t row_off;     // first row (frame start-H) of the utterance in every module buffer
  int64_t out_off;     // first row of the utterance in the gathered output
  int n_frames;        // valid frames 0..n_frames-1
  int start;           // reference frame number of the first output row
  int n_rows_out;      // output rows (frames start .. start+n_rows_out-1)
  int pad;
};

constexpr int fe_log2(int v) { return v <= 1 ? 0 : 1 + fe_log2(v >> 1); }
// threads cooperating on one frame: the largest power of two <= min(256, max(32, M/2))
constexpr int fe_threads_per_frame(int M) { return M / 2 >= 256 ? 256 : (M / 2 <= 32 ? 32 : (1 << fe_log2(M / 2))); }

__device__ __forceinline__ void row_to_frame(const int *__restrict__ row_utt, const UttDesc *__restrict__ utts,
                                             int64_t r, int H, int &u, int &t)
{
  u = row_utt[r];
  t = utts[u].start - H + (int)(r - utts[u].row_off);
}

// row r of the module buffers belongs to the utterance whose [row_off, next row_off) holds it: binary search.
__global__ void fe_build_row_utt(const UttDesc *__restrict__ utts, int n_utts, int64_t n_rows, int *__restrict__ row_utt)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int lo = 0, hi = n_utts;                      // largest u with utts[u].row_off <= r
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (utts[mid].row_off <= r) lo = mid; else hi = mid; }
  row_utt[r] = lo;
}

// PreModule::generate (aku/FeatureModules.cc:705-755): the base module's rows are stored float32 features; frames before
// the file repeat the first row, frames after it the last one.  raw: all utterances' rows back to back.
__global__ void fe_pre_base(const float *__restrict__ raw, int dim, const UttDesc *__restrict__ utts, const int *__restrict__ row_utt,
                            int64_t n_rows, int H, double *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  const int64_t r = i / dim;
  const int c = (int)(i - r * dim);
  int u, t;
  row_to_frame(row_utt, utts, r, H, u, t);
  const UttDesc ud = utts[u];
  const int tc = min(max(t, 0), ud.n_frames - 1);
  out[i] = (double)raw[(ud.pcm_off + tc) * dim + c];
}

// MelModule::generate, aku/FeatureModules.cc:806-849, one bin from the precomputed triangle weights (see
// upload_module_params).  SRC yields the spectrum value as a double.
struct MelTable { const float *scale; const int *desc; const float *sum; };
template <class SRC>
__device__ __forceinline__ double mel_bin(SRC data, int sdim, const MelTable mt, int b, int root)
{
  const int t0 = mt.desc[3 * b], n = mt.desc[3 * b + 1];
  const float *sc = mt.scale + mt.desc[3 * b + 2];
  float val = 0;
  for (int i = 0; i < n; i++)       // float accumulator fed through a double product, as the reference's `val += scale * data`
    val = (float)__dadd_rn((double)val, __dmul_rn((double)sc[i], data(min(t0 + i, sdim - 1))));
  const float sum = mt.sum[b];
  if (root) return pow((double)__fdiv_rn(val, sum), 0.1);
  return (double)logf(__fadd_rn(__fdiv_rn(val, sum), 1.f));
}

// Fused path fft -> {mel -> dct, power} -> merge (the static part of an MFCC chain): instead of a spectrum matrix the
// FFT kernel keeps the frame's spectrum in shared memory and writes the merged row (dct columns + power column).
struct FuseStatic {
  int on = 0;
  int mel_dim = 0, mel_root = 0;
  MelTable mel = {nullptr, nullptr, nullptr};
  int dct_dim = 0, dct_col = 0;
  const float *dct_table = nullptr;
  const double *mel_scale_d = nullptr, *dct_table_d = nullptr;   // the same tables widened to double
  int pow_col = -1;            // -1: no power module
  int odim = 0;
  double *out = nullptr;       // [rows][odim]
};
constexpr int FUSE_MAX_MEL = 64;

// audiofile + fft for windows of 2^k or 3 * 2^k samples (16 kHz -> 256, 48 kHz -> 768, ...).  One CTA = FPB frames;
// the real signal is packed even/odd into an M = N/2-point complex FFT in shared memory: R = 1 or 3 interleaved
// radix-2 DIT sub-transforms of P = M/R points, one radix-3 combination pass when R = 3, then the real-FFT split.
template <int N>
__global__ void __launch_bounds__(256)
fe_spectrum_fft(const int16_t *__restrict__ pcm, const UttDesc *__restrict__ utts, const int *__restrict__ row_utt,
                int64_t n_rows, int H, float adv, float emph, int copy_borders, const float *__restrict__ window,
                const float2 *__restrict__ tw, int magnitude, int do_log, double *__restrict__ out, const FuseStatic fuse)
{
  constexpr int M = N / 2;                 // complex FFT size
  constexpr int R = (N % 3 == 0) ? 3 : 1;  // odd radix
  constexpr int P = M / R;                 // power-of-two sub-transform
  static_assert((P & (P - 1)) == 0 && P >= 16, "window must be 2^k or 3 * 2^k samples");
  constexpr int TPF = fe_threads_per_frame(M);
  constexpr int FPB = 256 / TPF;
  constexpr int LOGP = fe_log2(P);
  __shared__ float2 z[FPB][M + 1];
  const int fl = threadIdx.x / TPF, tl = threadIdx.x % TPF;
  const int64_t r = (int64_t)blockIdx.x * FPB + fl;
  const bool valid = r < n_rows;
  int u = 0, t = 0;
  if (valid) row_to_frame(row_utt, utts, r, H, u, t);
  const UttDesc ud = utts[u];
  // window start: AudioFileModule::generate, aku/FeatureModules.cc:378-397
  int tc = t;
  if (copy_borders) tc = min(max(t, 0), ud.n_frames - 1);
  const int ws = (int)(tc * adv);
  const int16_t *x = pcm + ud.pcm_off;
  for (int k = tl; k < M; k += TPF) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      int i = 2 * k + e;
      int64_t p0 = (int64_t)ws + i, p1 = p0 + 1;
      float s0 = (valid && p0 >= 0 && p0 < ud.n_samples) ? (float)x[p0] : 0.f;
      float s1 = (valid && p1 >= 0 && p1 < ud.n_samples) ? (float)x[p1] : 0.f;
      float pe = __fsub_rn(s1, __fmul_rn(emph, s0));                    // float pre-emphasis (:429)
      v[e] = (float)__dmul_rn((double)window[i], (double)pe);           // float window * double sample -> float (:531)
    }
    // decimation in time by R, each sub-sequence in bit-reversed order
    const int sub = k % R, m = k / R;
    z[fl][sub * P + (int)(__brev((unsigned)m) >> (32 - LOGP))] = make_float2(v[0], v[1]);
  }
  __syncthreads();
  // radix-2 DIT on the R sub-transforms; twiddle e^{-2 pi i j / len} = tw[j * N / len]
  for (int len = 2; len <= P; len <<= 1) {
    const int half = len >> 1;
    for (int b = tl; b < M / 2; b += TPF) {
      const int sub = b / (P / 2), bb = b - sub * (P / 2);
      int j = bb & (half - 1);
      int i0 = sub * P + ((bb - j) << 1) + j, i1 = i0 + half;
      float2 w = tw[j * (N / len)];
      float2 a = z[fl][i0], c = z[fl][i1];
      float2 wc = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
      z[fl][i0] = make_float2(a.x + wc.x, a.y + wc.y);
      z[fl][i1] = make_float2(a.x - wc.x, a.y - wc.y);
    }
    __syncthreads();
  }
  if (R == 3) {
    // Z[k + qP] = Y0[k] + w3^q W_M^k Y1[k] + w3^2q W_M^2k Y2[k],  W_M^k = tw[2k],  w3 = e^{-2 pi i / 3}
    const float c3 = -0.5f, s3 = -0.86602540378443864676f;
    for (int k = tl; k < P; k += TPF) {
      const float2 y0 = z[fl][k], y1 = z[fl][P + k], y2 = z[fl][2 * P + k];
      const float2 w1 = tw[2 * k], w2 = tw[4 * k];
      const float2 t1 = make_float2(y1.x * w1.x - y1.y * w1.y, y1.x * w1.y + y1.y * w1.x);
      const float2 t2 = make_float2(y2.x * w2.x - y2.y * w2.y, y2.x * w2.y + y2.y * w2.x);
      const float2 sum = make_float2(t1.x + t2.x, t1.y + t2.y), dif = make_float2(t1.x - t2.x, t1.y - t2.y);
      // w3 t1 + conj(w3) t2 = c3 * sum + i s3 * dif ;  conj(w3) t1 + w3 t2 = c3 * sum - i s3 * dif
      const float2 base = make_float2(y0.x + c3 * sum.x, y0.y + c3 * sum.y);
      const float2 rot = make_float2(-s3 * dif.y, s3 * dif.x);
      z[fl][k] = make_float2(y0.x + sum.x, y0.y + sum.y);
      z[fl][P + k] = make_float2(base.x + rot.x, base.y + rot.y);
      z[fl][2 * P + k] = make_float2(base.x - rot.x, base.y - rot.y);
    }
    __syncthreads();
  }
  // split: X[k] = (Z[k]+conj(Z[M-k]))/2 - i e^{-2 pi i k/N} (Z[k]-conj(Z[M-k]))/2 ,  k = 0..M
  __shared__ float pw[FPB][M + 2];         // fused path: the frame's spectrum stays on chip
  __shared__ double melv[FPB][FUSE_MAX_MEL];
  {
    double *o = out + r * (M + 1);
    for (int k = tl; k <= M; k += TPF) {
      float2 a = z[fl][k == M ? 0 : k], b = z[fl][k == 0 ? 0 : M - k];
      float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
      float2 w = (k == M) ? make_float2(-1.f, 0.f) : tw[k];
      // -i * w * d
      float2 wd = make_float2(w.x * d.x - w.y * d.y, w.x * d.y + w.y * d.x);
      float re = e.x + wd.y, im = e.y - wd.x;
      float p = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));       // float power (:534-537)
      if (magnitude) p = sqrtf(p);
      if (do_log) p = logf(p);
      if (fuse.on) pw[fl][k] = p;
      else if (valid) o[k] = (double)p;
    }
  }
  if (!fuse.on) return;
  __syncthreads();
  // mel bins (one thread each) next to the power sum (one thread): the operations of fe_mel / fe_power on the
  // on-chip spectrum.  power: float(double(p) + double(s)) of two floats is the fp32 sum (the double sum is exact or
  // the smaller term is below a quarter ulp), so the float accumulator of PowerModule (:875-885) is a chain of FADDs.
  {
    const float *sp = pw[fl];
    if (tl < fuse.mel_dim)
      melv[fl][tl] = mel_bin([sp](int t) { return (double)sp[t]; }, M + 1, fuse.mel, tl, fuse.mel_root);
  }
  if (fuse.pow_col >= 0 && threadIdx.x < FPB) {           // the FPB power sums share warp 0
    const float *sp = pw[threadIdx.x];
    const int64_t rr = (int64_t)blockIdx.x * FPB + threadIdx.x;
    float power = 0;
    for (int i = 0; i <= M; i++) power = __fadd_rn(power, sp[i]);
    if (rr < n_rows) fuse.out[rr * fuse.odim + fuse.pow_col] = log(__dadd_rn((double)power, 1e-10));
  }
  __syncthreads();
  if (valid && tl < fuse.dct_dim) {                       // DCTModule::generate (:956-979)
    const float *tb = fuse.dct_table + (size_t)tl * fuse.mel_dim;
    double acc = 0.0;
    for (int b = 0; b < fuse.mel_dim; b++) acc = __dadd_rn(acc, __dmul_rn(melv[fl][b], (double)tb[b]));
    fuse.out[r * fuse.odim + fuse.dct_col + tl] = acc;
  }
}

// audiofile + fft with ONE WARP PER FRAME for power-of-two windows (128 ... 2048 samples): the N/2-point complex FFT lives
// in registers -- lane l holds the elements l, l + 32, ... -- the radix-2 DIT stages whose partners sit in another lane
// exchange them with shuffles (both lanes form the same twiddle product: same operations, same bits as a shared-memory
// butterfly), the later stages are register-to-register; no CTA barrier anywhere.  The real-FFT split goes once through
// a per-warp shared-memory row (conjugate-symmetric gather), and the fused epilogue (mel bins on 21 lanes, the sequential
// float power sum on the last lane, the dct on 12 lanes) is warp-local as well.  Float / double placement as in
// fe_spectrum_fft; the mel and dct tables are read widened to double, the spectrum is kept widened next to its floats
// (two conversions per mel tap instead of four, none in the dct).
// One frame's spectrum by one warp: window, pre-emphasis, N/2-point complex FFT in registers, real-FFT split.  The
// values go to the warp's shared-memory row(s) (fused paths) or, widened, to row r of `out`.
template <int N>
__device__ __forceinline__ void wfft_frame(const int16_t *__restrict__ pcm, const UttDesc *__restrict__ utts, const int *__restrict__ row_utt,
                                           int64_t r, bool valid, int H, float adv, float emph, int copy_borders,
                                           const float *__restrict__ window, const float2 *__restrict__ tw, int magnitude, int do_log,
                                           double *__restrict__ out, float2 *zsw, float *pwf_row, double *pwd_row, int lane)
{
  constexpr int M = N / 2, NR = M / 32, LOGM = fe_log2(M);
  int u = 0, t = 0;
  if (valid) row_to_frame(row_utt, utts, r, H, u, t);
  const UttDesc ud = utts[u];
  int tc = t;
  if (copy_borders) tc = min(max(t, 0), ud.n_frames - 1);
  const int ws = (int)(tc * adv);                              // window start: AudioFileModule::generate, aku/FeatureModules.cc:378-397
  const int16_t *x = pcm + ud.pcm_off;
  // frames that lie inside the audio (all but the first / last of an utterance) need no range test per sample
  const bool inside = valid && ws >= 0 && (int64_t)ws + N < ud.n_samples;
  const int16_t *xw = x + ws;
  float2 z[NR];
#pragma unroll
  for (int rr = 0; rr < NR; rr++) {
    const int pos = rr * 32 + lane;                            // position after the bit reversal of the decimation in time
    const int k = (int)(__brev((unsigned)pos) >> (32 - LOGM));
    float s0, s1, s2;
    if (inside) {
      s0 = (float)xw[2 * k]; s1 = (float)xw[2 * k + 1]; s2 = (float)xw[2 * k + 2];
    } else {
      const int64_t p0 = (int64_t)ws + 2 * k;
      s0 = (valid && p0 >= 0 && p0 < ud.n_samples) ? (float)x[p0] : 0.f;
      s1 = (valid && p0 + 1 >= 0 && p0 + 1 < ud.n_samples) ? (float)x[p0 + 1] : 0.f;
      s2 = (valid && p0 + 2 >= 0 && p0 + 2 < ud.n_samples) ? (float)x[p0 + 2] : 0.f;
    }
    const float2 wn = __ldg(reinterpret_cast<const float2 *>(window) + k);
    // float pre-emphasis (:429); float window * double sample -> float (:531): the double product of two floats is
    // exact, so its rounding to float is the fp32 product
    z[rr] = make_float2(__fmul_rn(wn.x, __fsub_rn(s1, __fmul_rn(emph, s0))), __fmul_rn(wn.y, __fsub_rn(s2, __fmul_rn(emph, s1))));
  }
  // stages whose partner is in another lane: half = 1 .. 16.  The upper lane of a pair multiplies its element by the
  // twiddle, the lower one by (1, 0) -- exact --, both send the product: lower = own + received, upper = received - own
  // (one fused multiply-add by +-1: the rounding of the sum / difference).  Same operations, same bits as a butterfly
  // that computes the twiddle product on both sides.
#pragma unroll
  for (int half = 1; half < 32; half <<= 1) {
    const bool hi = (lane & half) != 0;
    float2 wv = __ldg(tw + (lane & (half - 1)) * (N / (2 * half)));         // twiddle e^{-2 pi i j / len} = tw[j * N / len]
    if (!hi) wv = make_float2(1.f, 0.f);
    const float sg = hi ? -1.f : 1.f;
#pragma unroll
    for (int rr = 0; rr < NR; rr++) {
      const float2 c = z[rr];
      const float2 t = make_float2(c.x * wv.x - c.y * wv.y, c.x * wv.y + c.y * wv.x);
      const float2 o = make_float2(__shfl_xor_sync(0xffffffffu, t.x, half), __shfl_xor_sync(0xffffffffu, t.y, half));
      z[rr] = make_float2(__fmaf_rn(t.x, sg, o.x), __fmaf_rn(t.y, sg, o.y));
    }
  }
  // stages inside the lane: half = 32 .. M / 2, register distance half / 32
#pragma unroll
  for (int hr = 1; hr < NR; hr <<= 1) {
    const int half = hr * 32;
#pragma unroll
    for (int rr = 0; rr < NR; rr++) {
      if (rr & hr) continue;
      const int j = ((rr & (hr - 1)) << 5) | lane;
      const float2 wv = __ldg(tw + j * (N / (2 * half)));
      const float2 a = z[rr], c = z[rr + hr];
      const float2 wc = make_float2(c.x * wv.x - c.y * wv.y, c.x * wv.y + c.y * wv.x);
      z[rr] = make_float2(a.x + wc.x, a.y + wc.y);
      z[rr + hr] = make_float2(a.x - wc.x, a.y - wc.y);
    }
  }
#pragma unroll
  for (int rr = 0; rr < NR; rr++) zsw[rr * 32 + lane] = z[rr];
  __syncwarp();
  // split: X[k] = (Z[k]+conj(Z[M-k]))/2 - i e^{-2 pi i k/N} (Z[k]-conj(Z[M-k]))/2 ,  k = 0..M
  {
    double *o = out + r * (M + 1);
    for (int k = lane; k <= M; k += 32) {
      const float2 a = zsw[k == M ? 0 : k], b = zsw[k == 0 ? 0 : M - k];
      const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
      const float2 wv = (k == M) ? make_float2(-1.f, 0.f) : __ldg(tw + k);
      const float2 wd = make_float2(wv.x * d.x - wv.y * d.y, wv.x * d.y + wv.y * d.x);      // -i * w * d below
      const float re = e.x + wd.y, im = e.y - wd.x;
      float p = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));       // float power (:534-537)
      if (magnitude) p = sqrtf(p);
      if (do_log) p = logf(p);
      if (pwf_row) { pwf_row[k] = p; if (pwd_row) pwd_row[k] = (double)p; }
      else if (valid) o[k] = (double)p;
    }
  }
}

template <int N>
__global__ void __launch_bounds__(256)
fe_spectrum_wfft(const int16_t *__restrict__ pcm, const UttDesc *__restrict__ utts, const int *__restrict__ row_utt,
                 int64_t n_rows, int H, float adv, float emph, int copy_borders, const float *__restrict__ window,
                 const float2 *__restrict__ tw, int magnitude, int do_log, double *__restrict__ out, const FuseStatic fuse)
{
  constexpr int M = N / 2;
  constexpr int WPB = 8;                                        // frames (warps) per CTA
  static_assert((M & (M - 1)) == 0 && M >= 32 && M <= 256, "window must be a power of two, 64 ... 512 samples");
  __shared__ float2 zs[WPB][M];
  __shared__ __align__(16) float pwf[WPB][M + 4];
  __shared__ double pwd[WPB][M + 2];
  __shared__ double melv[WPB][FUSE_MAX_MEL];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * WPB + w;
  const bool valid = r < n_rows;
  wfft_frame<N>(pcm, utts, row_utt, r, valid, H, adv, emph, copy_borders, window, tw, magnitude, do_log, out, zs[w],
                fuse.on ? pwf[w] : nullptr, fuse.on ? pwd[w] : nullptr, lane);
  if (!fuse.on) return;
  __syncwarp();
  // mel bins: MelModule::generate (:806-849) from the precomputed triangle weights; float accumulator fed through a
  // double product, as the reference's `val += scale * data`
  for (int b = lane; b < fuse.mel_dim; b += 32) {
    const int t0 = fuse.mel.desc[3 * b], n = fuse.mel.desc[3 * b + 1];
    const double *sc = fuse.mel_scale_d + fuse.mel.desc[3 * b + 2];
    const double *pw = pwd[w] + t0;
    const int n_in = min(n, M + 1 - t0);                       // taps inside the spectrum; the others read its last bin
    float val = 0;
    int i = 0;
#pragma unroll 4
    for (; i < n_in; i++) val = (float)__dadd_rn((double)val, __dmul_rn(sc[i], pw[i]));
    const double last = pwd[w][M];
    for (i = max(i, 0); i < n; i++) val = (float)__dadd_rn((double)val, __dmul_rn(sc[i], last));
    const float sum = fuse.mel.sum[b];
    melv[w][b] = fuse.mel_root ? pow((double)__fdiv_rn(val, sum), 0.1) : (double)logf(__fadd_rn(__fdiv_rn(val, sum), 1.f));
  }
  // power: the float accumulator of PowerModule (:875-885) is a chain of FADDs (see fe_spectrum_fft); the last lane has
  // no mel bin to compute in the usual configurations
  if (fuse.pow_col >= 0 && lane == 31) {
    float power = 0;
    const float4 *p4 = reinterpret_cast<const float4 *>(pwf[w]);
#pragma unroll 4
    for (int i = 0; i < M / 4; i++) {
      const float4 q = p4[i];
      power = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(power, q.x), q.y), q.z), q.w);
    }
    power = __fadd_rn(power, pwf[w][M]);
    if (valid) fuse.out[r * fuse.odim + fuse.pow_col] = log(__dadd_rn((double)power, 1e-10));
  }
  __syncwarp();
  if (valid)
    for (int c = lane; c < fuse.dct_dim; c += 32) {          // DCTModule::generate (:956-979)
      const double *tb = fuse.dct_table_d + (size_t)c * fuse.mel_dim;
      double acc = 0.0;
      for (int b = 0; b < fuse.mel_dim; b++) acc = __dadd_rn(acc, __dmul_rn(melv[w][b], tb[b]));
      fuse.out[r * fuse.odim + fuse.dct_col + c] = acc;
    }
}

static size_t bfft_smem_bytes(int N)
{
  const int M = N / 2, FPC = 32, WPB = 8, MAXB = 32, MAXC = 32;
  const int pw_floats = std::max(FPC * (M + 1), 2 * FPC * (MAXC + 1));
  return sizeof(double) * (MAXB + 1) * FPC + sizeof(float2) * WPB * M + sizeof(float) * pw_floats;
}
// The same chain with the epilogue turned by 90 degrees (fused MFCC graphs with at most 32 mel bins).  In the kernel
// above a frame's epilogue runs on the lanes of ONE warp: 21 lanes walk their triangles (38 sequential taps for the
// widest), one lane adds the 129 spectrum values in the prescribed order, 12 lanes form the dct -- ~600 warp instructions
// per frame at a quarter of the lanes.  Here a CTA of four warps first transforms 32 frames (eight per warp, spectra kept
// in shared memory) and then runs the epilogue with LANE = FRAME: a mel bin, the power sum or a dct coefficient is one
// task for 32 frames at once (tasks handed out longest first through a shared counter), every lane performing exactly
// the operation sequence of the kernel above for its frame -- same bits, ~100 instructions per frame.
template <int N>
__global__ void __launch_bounds__(256, 6)
fe_spectrum_bfft(const int16_t *__restrict__ pcm, const UttDesc *__restrict__ utts, const int *__restrict__ row_utt,
                 int64_t n_rows, int H, float adv, float emph, int copy_borders, const float *__restrict__ window,
                 const float2 *__restrict__ tw, int magnitude, int do_log, const FuseStatic fuse)
{
  constexpr int M = N / 2;
  constexpr int WPB = 8, FPW = 4, FPC = WPB * FPW;              // 32 frames per CTA
  constexpr int MAXB = 32, MAXC = 32;
  static_assert((M & (M - 1)) == 0 && M >= 32 && M <= 256, "window of 128, 256 or 512 samples");
  constexpr int PW_FLOATS = FPC * (M + 1) > 2 * FPC * (MAXC + 1) ? FPC * (M + 1) : 2 * FPC * (MAXC + 1);
  // dynamic shared memory (57 KB at 512 samples): melv [bin][frame] (the last row carries the power column) | zs, one
  // row per warp | pw = spectra [frame][M + 1] (odd row length: lane = frame reads are conflict free), later the merged
  // rows [frame][MAXC + 1] doubles
  extern __shared__ __align__(16) unsigned char bfft_sm[];
  double (*melv)[FPC] = reinterpret_cast<double (*)[FPC]>(bfft_sm);
  float2 (*zs)[M] = reinterpret_cast<float2 (*)[M]>(bfft_sm + sizeof(double) * (MAXB + 1) * FPC);
  float *pw = reinterpret_cast<float *>(bfft_sm + sizeof(double) * (MAXB + 1) * FPC + sizeof(float2) * WPB * M);
  static_assert(PW_FLOATS > 0, "");
  __shared__ int next_task;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * FPC;
  if (threadIdx.x == 0) next_task = 0;
  for (int i = 0; i < FPW; i++) {
    const int fl = w * FPW + i;
    const int64_t r = r0 + fl;
    wfft_frame<N>(pcm, utts, row_utt, r, r < n_rows, H, adv, emph, copy_borders, window, tw, magnitude, do_log, nullptr, zs[w],
                  pw + fl * (M + 1), nullptr, lane);
    __syncwarp();                                               // zs[w] is rewritten by the warp's next frame
  }
  __syncthreads();
  // tasks: 0 = power sum (the longest), then the mel bins in pairs from the widest triangles down: two independent
  // accumulator chains per lane
  const float *sp = pw + lane * (M + 1);
  const int has_pow = fuse.pow_col >= 0 ? 1 : 0;
  const int n_tasks = (fuse.mel_dim + 1) / 2 + has_pow;
  for (;;) {
    int task = 0;
    if (lane == 0) task = atomicAdd(&next_task, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= n_tasks) break;
    if (has_pow && task == 0) {
      // the float accumulator of PowerModule (:875-885): a chain of FADDs (see fe_spectrum_fft)
      float power = 0;
#pragma unroll 4
      for (int k = 0; k <= M; k++) power = __fadd_rn(power, sp[k]);
      melv[MAXB][lane] = log(__dadd_rn((double)power, 1e-10));
      continue;
    }
    // MelModule::generate (:806-849): float accumulator fed through a double product, `val += scale * data`
    const int b0 = fuse.mel_dim - 1 - 2 * (task - has_pow), b1 = b0 - 1;     // b1 = -1: an odd bin out
    const int t00 = __ldg(fuse.mel.desc + 3 * b0), n0 = __ldg(fuse.mel.desc + 3 * b0 + 1);
    const double *sc0 = fuse.mel_scale_d + __ldg(fuse.mel.desc + 3 * b0 + 2);
    const int t01 = b1 >= 0 ? __ldg(fuse.mel.desc + 3 * b1) : 0, n1 = b1 >= 0 ? __ldg(fuse.mel.desc + 3 * b1 + 1) : 0;
    const double *sc1 = fuse.mel_scale_d + (b1 >= 0 ? __ldg(fuse.mel.desc + 3 * b1 + 2) : 0);
    float v0 = 0, v1 = 0;
    const int nn = max(n0, n1);
    for (int i = 0; i < nn; i++) {                              // a tap past the spectrum reads its last bin
      if (i < n0) v0 = (float)__dadd_rn((double)v0, __dmul_rn(__ldg(sc0 + i), (double)sp[min(t00 + i, M)]));
      if (i < n1) v1 = (float)__dadd_rn((double)v1, __dmul_rn(__ldg(sc1 + i), (double)sp[min(t01 + i, M)]));
    }
    const float s0 = __ldg(fuse.mel.sum + b0);
    melv[b0][lane] = fuse.mel_root ? pow((double)__fdiv_rn(v0, s0), 0.1) : (double)logf(__fadd_rn(__fdiv_rn(v0, s0), 1.f));
    if (b1 >= 0) {
      const float s1 = __ldg(fuse.mel.sum + b1);
      melv[b1][lane] = fuse.mel_root ? pow((double)__fdiv_rn(v1, s1), 0.1) : (double)logf(__fadd_rn(__fdiv_rn(v1, s1), 1.f));
    }
  }
  __syncthreads();                                              // the spectra are dead: their memory takes the merged rows
  double *orow = reinterpret_cast<double *>(pw);                // [frame][MAXC + 1]
  for (int c = w; c < fuse.dct_dim; c += WPB) {                // DCTModule::generate (:956-979), one coefficient x 32 frames
    const double *tb = fuse.dct_table_d + (size_t)c * fuse.mel_dim;
    double acc = 0.0;
    for (int b = 0; b < fuse.mel_dim; b++) acc = __dadd_rn(acc, __dmul_rn(melv[b][lane], __ldg(tb + b)));
    orow[lane * (MAXC + 1) + fuse.dct_col + c] = acc;
  }
  if (has_pow && w == 0) orow[lane * (MAXC + 1) + fuse.pow_col] = melv[MAXB][lane];
  __syncthreads();
  const int n_here = (int)min((int64_t)FPC, n_rows - r0);
  double *o = fuse.out + r0 * fuse.odim;
  for (int i = threadIdx.x; i < n_here * fuse.odim; i += blockDim.x) o[i] = orow[(i / fuse.odim) * (MAXC + 1) + i % fuse.odim];
}

// Fused tail  X -> delta -> delta -> merge(X, d1, d2) -> output rows without the halo (DeltaModule::generate :1019-1037
// applied twice, MergerModule :1352-1364): one thread per (output row, column of X); the same operations in the same
// order as the per-module kernels, the intermediate rows recomputed instead of stored.
constexpr int FUSE_MAX_WIDTH = 4;
template <class T>
__global__ void fe_delta2_merge(const double *__restrict__ src, int dim, int64_t n_rows, const UttDesc *__restrict__ utts,
                                const int *__restrict__ row_utt, int H, int w1, float norm1, int w2, float norm2,
                                T *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  const int64_t r = i / dim;                               // row in the haloed matrices
  const int c = (int)(i - r * dim);
  const UttDesc ud = utts[row_utt[r]];
  const int64_t k = r - ud.row_off - H;                    // frame within the utterance's output
  if (k < 0 || k >= ud.n_rows_out) return;                 // halo rows only feed their neighbours
  const int64_t q = ud.out_off + k;
  const double *x = src + r * dim + c;
  double d1[2 * FUSE_MAX_WIDTH + 1];
  for (int j = -w2; j <= w2; j++) {
    double acc = 0;
    for (int kk = 1; kk <= w1; kk++)
      acc = __dadd_rn(acc, __dmul_rn((double)kk, __dsub_rn(x[(int64_t)(j + kk) * dim], x[(int64_t)(j - kk) * dim])));
    d1[j + FUSE_MAX_WIDTH] = __ddiv_rn(acc, (double)norm1);
  }
  double acc = 0;
  for (int kk = 1; kk <= w2; kk++)
    acc = __dadd_rn(acc, __dmul_rn((double)kk, __dsub_rn(d1[kk + FUSE_MAX_WIDTH], d1[-kk + FUSE_MAX_WIDTH])));
  T *o = out + q * (3 * dim);
  o[c] = (T)x[0];
  o[dim + c] = (T)d1[FUSE_MAX_WIDTH];
  o[2 * dim + c] = (T)__ddiv_rn(acc, (double)norm2);
}

// The same fused tail, tiled: a CTA takes 64 consecutive rows of X into shared memory (+ the halo), forms the first delta
// ONCE per (row, column) -- the kernel above recomputes it for each of the 2 w2 + 1 rows that use it: six double divisions
// and a 64-bit index division per output element, 640 instructions -- and the second delta from that tile; the rows'
// output positions are looked up once per row.  Operations and their order per value are the ones above: same bits.
constexpr int DELTA_TILE_ROWS = 64;
template <class T>
__global__ void __launch_bounds__(256)
fe_delta2_merge_tiled(const double *__restrict__ src, int dim, int64_t n_rows, const UttDesc *__restrict__ utts,
                      const int *__restrict__ row_utt, int H, int w1, float norm1, int w2, float norm2, T *__restrict__ out)
{
  constexpr int R = DELTA_TILE_ROWS;
  extern __shared__ double delta_sm[];
  __shared__ long long qrow[R];
  const int hx = w1 + w2, hd = w2;
  double *xs = delta_sm;                                  // rows r0 - hx .. r0 + R + hx of X
  double *ds = xs + (R + 2 * hx) * dim;                   // rows r0 - hd .. r0 + R + hd of the first delta
  const int64_t r0 = (int64_t)blockIdx.x * R;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const int64_t r = r0 + i;
    long long q = -1;
    if (r < n_rows) {
      const UttDesc ud = utts[row_utt[r]];
      const int64_t k = r - ud.row_off - H;                // frame within the utterance's output
      if (k >= 0 && k < ud.n_rows_out) q = ud.out_off + k; // halo rows only feed their neighbours
    }
    qrow[i] = q;
  }
  {
    const int64_t e0 = (r0 - hx) * dim, e_end = n_rows * dim;
    const int nx = (R + 2 * hx) * dim;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      const int64_t e = e0 + i;
      xs[i] = (e >= 0 && e < e_end) ? src[e] : 0.0;       // rows outside the matrix feed no output row
    }
  }
  __syncthreads();
  {
    const int nd = (R + 2 * hd) * dim;
    for (int i = threadIdx.x; i < nd; i += blockDim.x) {
      const double *x = xs + i + w1 * dim;
      double acc = 0;
      for (int kk = 1; kk <= w1; kk++) acc = __dadd_rn(acc, __dmul_rn((double)kk, __dsub_rn(x[kk * dim], x[-kk * dim])));
      ds[i] = __ddiv_rn(acc, (double)norm1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * dim; i += blockDim.x) {
    const int rl = i / dim, c = i - rl * dim;
    const long long q = qrow[rl];
    if (q < 0) continue;
    const double *d = ds + i + hd * dim;
    double acc = 0;
    for (int kk = 1; kk <= w2; kk++) acc = __dadd_rn(acc, __dmul_rn((double)kk, __dsub_rn(d[kk * dim], d[-kk * dim])));
    T *o = out + q * (3 * dim);
    o[c] = (T)xs[i + hx * dim];
    o[dim + c] = (T)d[0];
    o[2 * dim + c] = (T)__ddiv_rn(acc, (double)norm2);
  }
}

// Generic window length: direct DFT with the twiddle table (O(N^2); correctness path for the
// non-power-of-two windows 44.1/48 kHz configurations produce).
__global__ void __launch_bounds__(128)
fe_spectrum_dft(const int16_t *__restrict__ pcm, const UttDesc *__restrict__ utts, const int *__restrict__ row_utt,
                int64_t n_rows, int H, float adv, float emph, int copy_borders, const float *__restrict__ window,
                const float2 *__restrict__ tw, int N, int magnitude, int do_log, double *__restrict__ out)
{
  extern __shared__ float xs[];
  const int64_t r = blockIdx.x;
  int u, t;
  row_to_frame(row_utt, utts, r, H, u, t);
  const UttDesc ud = utts[u];
  int tc = t;
  if (copy_borders) tc = min(max(t, 0), ud.n_frames - 1);
  const int ws = (int)(tc * adv);
  const int16_t *x = pcm + ud.pcm_off;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    int64_t p0 = (int64_t)ws + i, p1 = p0 + 1;
    float s0 = (p0 >= 0 && p0 < ud.n_samples) ? (float)x[p0] : 0.f;
    float s1 = (p1 >= 0 && p1 < ud.n_samples) ? (float)x[p1] : 0.f;
    float pe = __fsub_rn(s1, __fmul_rn(emph, s0));
    xs[i] = (float)__dmul_rn((double)window[i], (double)pe);
  }
  __syncthreads();
  const int nb = N / 2 + 1;
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int n = 0; n < N; n++) {
      float2 w = tw[idx];
      re = fmaf(xs[n], w.x, re);
      im = fmaf(xs[n], w.y, im);
      idx += k;
      if (idx >= N) idx -= N;
    }
    float p = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
    if (magnitude) p = sqrtf(p);
    if (do_log) p = logf(p);
    out[r * nb + k] = (double)p;
  }
}

// MelModule::generate, aku/FeatureModules.cc:806-849; one thread per (row, bin).
__global__ void fe_mel(const double *__restrict__ src, int sdim, int64_t n_rows, const MelTable mt, int dim,
                       int root, double *__restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  int64_t r = i / dim;
  int b = (int)(i - r * dim);
  const double *data = src + r * sdim;
  out[r * dim + b] = mel_bin([data](int t) { return data[t]; }, sdim, mt, b, root);
}

// PowerModule (:875-885) and MelPowerModule (:908-919); one thread per row, float accumulator.
__global__ void fe_power(const double *__restrict__ src, int sdim, int64_t n_rows, int use_exp, double *__restrict__ out)
{
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const double *s = src + r * sdim;
  float power = 0;
  for (int i = 0; i < sdim; i++) power = (float)__dadd_rn((double)power, use_exp ? exp(s[i]) : s[i]);   // double(float)+double: see fe_spectrum_fft
  out[r] = log(__dadd_rn((double)power, 1e-10));
}

// DCTModule::generate (:956-979); table[i][b] = cosf(...) as float, double accumulate in b order.
__global__ void fe_dct(const double *__restrict__ src, int sdim, int64_t n_rows, const float *__restrict__ table, int dim,
                       double *__restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  int64_t r = i / dim;
  int c = (int)(i - r * dim);
  const double *s = src + r * sdim;
  const float *tb = table + (size_t)c * sdim;
  double acc = 0.0;
  for (int b = 0; b < sdim; b++) acc = __dadd_rn(acc, __dmul_rn(s[b], (double)tb[b]));
  out[r * dim + c] = acc;
}

// Rows of a neighbouring frame stay inside the utterance's own row block because the halo H
// covers every module's context; the clamp only protects rows nobody consumes.
__device__ __forceinline__ int64_t nb_row(int64_t r, int k, int64_t lo, int64_t hi)
{
  int64_t q = r + k;
  return q < lo ? lo : (q > hi ? hi : q);
}

// DeltaModule::generate (:1019-1037)
__global__ void fe_delta(const double *__restrict__ src, int dim, int64_t n_rows, const int *__restrict__ row_utt,
                         const UttDesc *__restrict__ utts, int H, int width, float norm, double *__restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  int64_t r = i / dim;
  int c = (int)(i - r * dim);
  const UttDesc ud = utts[row_utt[r]];
  const int64_t lo = ud.row_off, hi = ud.row_off + ud.n_rows_out + 2 * H - 1;
  double acc = 0;
  for (int k = 1; k <= width; k++) {
    double rv = src[nb_row(r, k, lo, hi) * dim + c], lv = src[nb_row(r, -k, lo, hi) * dim + c];
    acc = __dadd_rn(acc, __dmul_rn((double)k, __dsub_rn(rv, lv)));
  }
  out[r * dim + c] = __ddiv_rn(acc, (double)norm);
}

// merge / concat: copy `sdim` columns of the source row shifted by `shift` frames to column `col0`.
__global__ void fe_copy(const double *__restrict__ src, int sdim, int64_t n_rows, const int *__restrict__ row_utt,
                        const UttDesc *__restrict__ utts, int H, int shift, double *__restrict__ out, int odim, int col0)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * sdim) return;
  int64_t r = i / sdim;
  int c = (int)(i - r * sdim);
  int64_t q = r;
  if (shift) {
    const UttDesc ud = utts[row_utt[r]];
    q = nb_row(r, shift, ud.row_off, ud.row_off + ud.n_rows_out + 2 * H - 1);
  }
  out[r * odim + col0 + c] = src[q * sdim + c];
}

// VtlnModule::generate (aku/FeatureModules.cc:1906-1934): every output bin is a short dot product of the source spectrum
// with the Lanczos-windowed sinc taps around its warped position (double accumulator, float coefficients, result clamped
// at 0 as a float), or a linear interpolation between the two neighbouring bins when sinc_interpolation_rad is 0.
__global__ void fe_vtln(const double *__restrict__ src, int dim, int64_t n_rows, const float *__restrict__ coef,
                        const int *__restrict__ desc, const float *__restrict__ bins, int rad, double *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  const int64_t r = i / dim;
  const int b = (int)(i - r * dim);
  const double *data = src + r * dim;
  if (rad > 0) {
    const int di = desc[3 * b], n = desc[3 * b + 1];
    const float *cf = coef + desc[3 * b + 2];
    double t = 0;
    for (int k = 0; k < n; k++) t = __dadd_rn(t, __dmul_rn(data[di + k], (double)cf[k]));
    out[i] = (double)fmaxf((float)t, 0.0f);
  } else {
    const float vb = bins[b];
    const float p = __fsub_rn(ceilf(vb), vb);
    out[i] = __dadd_rn(__dmul_rn((double)p, data[(int)floorf(vb)]), __dmul_rn((double)__fsub_rn(1.f, p), data[(int)ceilf(vb)]));
  }
}

// SRNormModule::generate (aku/FeatureModules.cc:2039-2061): the input row holds in_frames stacked frames; every output
// frame is a Lanczos-weighted sum of input frames (float taps, double data and accumulator), clamped at 0 as a float.
__global__ void fe_srnorm(const double *__restrict__ src, int sdim, int frame_dim, int out_frames, int64_t n_rows,
                          const float *__restrict__ coef, const int *__restrict__ desc, double *__restrict__ out)
{
  const int odim = out_frames * frame_dim;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * odim) return;
  const int64_t r = i / odim;
  const int c = (int)(i - r * odim);
  const int of = c / frame_dim, d = c - of * frame_dim;
  const double *data = src + r * sdim;
  const int fi = desc[3 * of], n = desc[3 * of + 1];
  const float *cf = coef + desc[3 * of + 2];
  double t = 0;
  for (int j = 0; j < n; j++) t = __dadd_rn(t, __dmul_rn((double)cf[j], data[(fi + j) * frame_dim + d]));
  out[i] = (double)fmaxf((float)t, 0.0f);
}

// QuantEqModule::generate (:2122-2140), operation for operation -- including the reference's parenthesisation, which puts
// the linear term into the exponent: q * (alpha * pow(x / q, gamma + (1 - alpha) * (x / q))).
__global__ void fe_quanteq(const double *__restrict__ src, int dim, int64_t n_rows, const float *__restrict__ par, int active,
                           double *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  if (!active) { out[i] = src[i]; return; }
  const int k = (int)(i % dim);
  const float alpha = par[k], gamma = par[dim + k], q = par[2 * dim + k];
  const double u = __ddiv_rn(src[i], (double)q);
  const double e = __dadd_rn((double)gamma, __dmul_rn((double)__fsub_rn(1.f, alpha), u));
  out[i] = __dmul_rn((double)q, __dmul_rn((double)alpha, pow(u, e)));
}

// NormalizationModule::generate (:1136-1142)
__global__ void fe_norm(const double *__restrict__ src, int dim, int64_t n_rows, const float *__restrict__ mean,
                        const float *__restrict__ scale, double *__restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  int c = (int)(i % dim);
  out[i] = __dmul_rn(__dsub_rn(src[i], (double)mean[c]), (double)scale[c]);
}

// LinTransformModule::generate (:1244-1269)
__global__ void fe_lintrans(const double *__restrict__ src, int sdim, int64_t n_rows, const float *__restrict__ mat,
                            int has_mat, const float *__restrict__ bias, int has_bias, int dim, double *__restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  int64_t r = i / dim;
  int c = (int)(i - r * dim);
  const double *s = src + r * sdim;
  double acc;
  if (has_mat) {
    acc = 0;
    const float *m = mat + (size_t)c * sdim;
    for (int j = 0; j < sdim; j++) acc = __dadd_rn(acc, __dmul_rn((double)m[j], s[j]));
  } else {
    acc = s[c];
  }
  if (has_bias) acc = __dadd_rn(acc, (double)bias[c]);
  out[r * dim + c] = acc;
}

// MeanSubtractorModule::generate, full-window branch (:1436-1447): mean over frames t-left..t+right.
__global__ void fe_meansub(const double *__restrict__ src, int dim, int64_t n_rows, const int *__restrict__ row_utt,
                           const UttDesc *__restrict__ utts, int H, int left, int right, double *__restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  int64_t r = i / dim;
  int c = (int)(i - r * dim);
  const UttDesc ud = utts[row_utt[r]];
  const int64_t lo = ud.row_off, hi = ud.row_off + ud.n_rows_out + 2 * H - 1;
  double m = 0;
  for (int k = -left; k <= right; k++) m = __dadd_rn(m, src[nb_row(r, k, lo, hi) * dim + c]);
  m = __ddiv_rn(m, (double)(left + right + 1));
  out[i] = __dsub_rn(src[i], m);
}

// Drop the halo: rows (u, start..start+n_rows_out-1) -> dense output, float or double.
template <class T>
__global__ void fe_gather(const double *__restrict__ src, int dim, const UttDesc *__restrict__ utts,
                          const int *__restrict__ row_utt, int64_t n_rows, int H, T *__restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  int64_t r = i / dim;
  int c = (int)(i - r * dim);
  const UttDesc ud = utts[row_utt[r]];
  int64_t k = r - ud.row_off - H;
  if (k < 0 || k >= ud.n_rows_out) return;
  out[(ud.out_off + k) * dim + c] = (T)src[i];
}

// ===================================================================================
// graph execution
namespace {

inline unsigned grid1(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

struct Plan {
  std::vector<char> needed;
  int H = 0;
};

Plan make_plan(const Frontend &fe, int target)
{
  Plan p;
  const int n = (int)fe.mods.size();
  p.needed.assign(n, 0);
  std::vector<int> el(n, 0), er(n, 0);
  p.needed[target] = 1;
  for (int m = target; m >= 0; m--) {
    if (!p.needed[m]) continue;
    const Module &mod = fe.mods[m];
    for (int s : mod.src) {
      p.needed[s] = 1;
      el[s] = std::max(el[s], el[m] + mod.left);
      er[s] = std::max(er[s], er[m] + mod.right);
    }
  }
  for (int m = 0; m < n; m++) if (p.needed[m]) p.H = std::max(p.H, std::max(el[m], er[m]));
  return p;
}

// Runs the graph for a set of utterance descriptors whose PCM is device resident.
void run_graph(akugpu_ctx *ctx, const void *d_in, std::vector<UttDesc> &utts, int target, void *d_out, int out_f64,
               int64_t out_row_base)
{
  const Frontend &fe = ctx->fe;
  if (target < 0) target = fe.last;
  Plan plan = make_plan(fe, target);
  const int H = plan.H;
  int64_t n_rows = 0;
  for (auto &u : utts) { u.row_off = n_rows; n_rows += (int64_t)u.n_rows_out + 2 * H; }
  if (n_rows == 0) return;
  cudaStream_t st = ctx->stream;
  DevBuf &d_utts = ctx->d_fe[0], &d_rowutt = ctx->d_fe[1];
  d_utts.reserve(utts.size() * sizeof(UttDesc));
  d_rowutt.reserve((size_t)n_rows * sizeof(int));
  AKU_CUDA(cudaMemcpyAsync(d_utts.p, utts.data(), utts.size() * sizeof(UttDesc), cudaMemcpyHostToDevice, st));
  // the pageable source above must stay alive until the copy is done
  AKU_CUDA(cudaStreamSynchronize(st));
  const UttDesc *du = d_utts.as<UttDesc>();
  // row -> utterance table, built on the device (a million-entry host loop + copy per call used to be a third of the
  // front-end stage of a config-2 step)
  fe_build_row_utt<<<grid1(n_rows, 256), 256, 0, st>>>(du, (int)utts.size(), n_rows, d_rowutt.as<int>());
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
  const int *dr = d_rowutt.as<int>();

  const int16_t *d_pcm = reinterpret_cast<const int16_t *>(d_in);      // audiofile base; a `pre` base reads float rows
  const bool pre_base = fe.mods[0].type == M_PRE;
  // ---- fusion patterns (the canonical MFCC chain); anything else runs module by module ----
  const int nm = (int)fe.mods.size();
  std::vector<std::vector<int>> cons(nm);
  for (int m = 1; m <= target; m++)
    if (plan.needed[m]) for (int sidx : fe.mods[m].src) cons[sidx].push_back(m);
  std::vector<char> skip(nm, 0);
  const Module &base = fe.mods[0];
  const char *no_fuse = getenv("AKUGPU_FE_NOFUSE");     // tests compare the fused kernels with the module-by-module path
  // (1) fft -> {mel -> dct, power} -> merge
  int f_fft = -1, f_mel = -1, f_dct = -1, f_pow = -1, f_out = -1;
  for (int m = 1; m <= target && f_fft < 0 && !no_fuse; m++) {
    if (!plan.needed[m] || fe.mods[m].type != M_FFT || m == target) continue;
    const int N = base.window_width;
    const bool fft_ok = N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048 || N == 192 || N == 384 || N == 768 || N == 1536;
    if (!fft_ok) break;
    int mel = -1, pw = -1;
    bool ok = true;
    for (int c : cons[m]) {
      if (fe.mods[c].type == M_MEL && mel < 0) mel = c;
      else if (fe.mods[c].type == M_POWER && pw < 0) pw = c;
      else ok = false;
    }
    if (!ok || mel < 0 || mel == target || cons[mel].size() != 1 || fe.mods[cons[mel][0]].type != M_DCT) break;
    const int dct = cons[mel][0];
    const int tpf = fe_threads_per_frame(N / 2);
    if (fe.mods[mel].dim > tpf || fe.mods[mel].dim > FUSE_MAX_MEL || fe.mods[dct].dim > tpf) break;
    if (pw < 0) {
      if (dct == target) break;      // the dct rows themselves are the result: nothing to merge into, keep it simple
      f_fft = m; f_mel = mel; f_dct = dct; f_out = dct;
    } else {
      if (pw == target || dct == target || cons[pw].size() != 1 || cons[dct].size() != 1 || cons[pw][0] != cons[dct][0]) break;
      const int mg = cons[pw][0];
      const Module &M = fe.mods[mg];
      if (M.type != M_MERGE || M.src.size() != 2) break;
      f_fft = m; f_mel = mel; f_dct = dct; f_pow = pw; f_out = mg;
    }
  }
  if (f_fft >= 0) { skip[f_mel] = 1; if (f_pow >= 0) { skip[f_pow] = 1; skip[f_dct] = 1; } if (f_out != f_dct) skip[f_out] = 1; else skip[f_dct] = 1; }
  // (2) X -> delta -> delta -> merge(X, d1, d2) == target
  int g_x = -1, g_d1 = -1, g_d2 = -1;
  if (!no_fuse && fe.mods[target].type == M_MERGE && fe.mods[target].src.size() == 3) {
    const std::vector<int> &sv = fe.mods[target].src;
    const int x = sv[0], d1 = sv[1], d2 = sv[2];
    if (x > 0 && fe.mods[d1].type == M_DELTA && fe.mods[d2].type == M_DELTA && fe.mods[d1].src[0] == x && fe.mods[d2].src[0] == d1 &&
        cons[d1].size() == 2 && cons[d2].size() == 1 && fe.mods[d1].width <= FUSE_MAX_WIDTH && fe.mods[d2].width <= FUSE_MAX_WIDTH &&
        fe.mods[x].type != M_FFT && fe.mods[x].type != M_AUDIOFILE) {
      g_x = x; g_d1 = d1; g_d2 = d2;
      skip[d1] = skip[d2] = skip[target] = 1;
    }
  }

  // one buffer per module that is materialised (the audiofile base module is fused into fft and has none)
  std::vector<std::shared_ptr<DevBuf>> buf(fe.mods.size());
  if (pre_base) {
    if (ctx->fe_bufs.empty()) ctx->fe_bufs.resize(1);
    if (!ctx->fe_bufs[0]) ctx->fe_bufs[0] = std::make_shared<DevBuf>();
    buf[0] = ctx->fe_bufs[0];
    buf[0]->reserve((size_t)n_rows * base.dim * sizeof(double));
    fe_pre_base<<<grid1(n_rows * base.dim, 256), 256, 0, st>>>(reinterpret_cast<const float *>(d_in), base.dim, du, dr, n_rows, H,
                                                              buf[0]->as<double>());
    AKU_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  for (int m = 1; m <= target; m++) {
    if (!plan.needed[m]) continue;
    if ((skip[m] && m != f_out) || m == f_fft) continue;
    if ((int)ctx->fe_bufs.size() <= m) ctx->fe_bufs.resize(m + 1);
    if (!ctx->fe_bufs[m]) ctx->fe_bufs[m] = std::make_shared<DevBuf>();
    buf[m] = ctx->fe_bufs[m];
    buf[m]->reserve((size_t)n_rows * fe.mods[m].dim * sizeof(double));
  }
  for (int m = 1; m <= target; m++) {
    if (!plan.needed[m] || skip[m]) continue;
    const Module &mod = fe.mods[m];
    if (mod.type != M_FFT && !pre_base)
      for (int sidx : mod.src)
        if (sidx == 0) throw Error(AKUGPU_E_CONFIG, "the audiofile module can only feed an fft module");
    double *o = buf[m] ? buf[m]->as<double>() : nullptr;
    const double *s0 = mod.src.empty() ? nullptr : (buf[mod.src[0]] ? buf[mod.src[0]]->as<double>() : nullptr);
    const int sdim = mod.src.empty() ? 0 : fe.mods[mod.src[0]].dim;
    const int64_t ne = n_rows * mod.dim;
    switch (mod.type) {
      case M_FFT: {
        const int N = base.window_width;
        const float *win = mod.d_a->as<float>();
        const float2 *tw = mod.d_b->as<float2>();
        FuseStatic fuse;
        if (m == f_fft) {
          const Module &mel = fe.mods[f_mel], &dct = fe.mods[f_dct];
          fuse.on = 1;
          fuse.mel_dim = mel.dim; fuse.mel_root = mel.root;
          fuse.mel = MelTable{mel.d_a->as<float>(), mel.d_b->as<int>(), mel.d_c->as<float>()};
          fuse.dct_dim = dct.dim; fuse.dct_table = dct.d_a->as<float>();
          fuse.mel_scale_d = mel.d_d->as<double>(); fuse.dct_table_d = dct.d_d->as<double>();
          fuse.odim = fe.mods[f_out].dim;
          fuse.out = buf[f_out]->as<double>();
          if (f_pow >= 0) {   // column order of the merge
            const bool dct_first = fe.mods[f_out].src[0] == f_dct;
            fuse.dct_col = dct_first ? 0 : 1;
            fuse.pow_col = dct_first ? dct.dim : 0;
          }
        }
#define SPEC_FFT(NN)                                                                                                \
  case NN: {                                                                                                        \
    constexpr int FPB_ = 256 / fe_threads_per_frame(NN / 2);                                                        \
    fe_spectrum_fft<NN><<<grid1(n_rows, FPB_), 256, 0, st>>>(d_pcm, du, dr, n_rows, H, base.window_advance,         \
                                                             base.emph, base.copy_borders, win, tw, mod.magnitude,  \
                                                             mod.log, o, fuse);                                     \
    break;                                                                                                          \
  }
#define SPEC_WFFT(NN)                                                                                               \
  case NN: {                                                                                                        \
    constexpr int WPB_ = 8;                                                                                         \
    if (fuse.on && fuse.mel_dim <= 32 && fuse.odim <= 32 && !one_warp_epilogue) {                                   \
      const size_t bsm = bfft_smem_bytes(NN);                                                                       \
      ensure_dynamic_smem(ctx, (const void *)fe_spectrum_bfft<NN>, bsm);                                            \
      fe_spectrum_bfft<NN><<<grid1(n_rows, 32), 256, bsm, st>>>(d_pcm, du, dr, n_rows, H, base.window_advance,      \
                                                               base.emph, base.copy_borders, win, tw,               \
                                                               mod.magnitude, mod.log, fuse);                       \
      break;                                                                                                        \
    }                                                                                                               \
    fe_spectrum_wfft<NN><<<grid1(n_rows, WPB_), WPB_ * 32, 0, st>>>(d_pcm, du, dr, n_rows, H, base.window_advance,  \
                                                                   base.emph, base.copy_borders, win, tw,           \
                                                                   mod.magnitude, mod.log, o, fuse);                \
    break;                                                                                                          \
  }
        const char *old_fft = getenv("AKUGPU_FE_OLDFFT");             // the shared-memory kernel, for comparisons
        static const bool one_warp_epilogue = getenv("AKUGPU_FE_WFFT1") != nullptr;   // the frame-per-warp epilogue, for comparisons
        if (old_fft && (N & (N - 1)) == 0) {
          switch (N) {
            SPEC_FFT(128)
            SPEC_FFT(256)
            SPEC_FFT(512)
            SPEC_FFT(1024)
            SPEC_FFT(2048)
          }
        } else
        switch (N) {       // measured (profiles/r02_config3_frontend.txt): one warp per frame wins up to 512 samples
          SPEC_WFFT(128)
          SPEC_WFFT(256)
          SPEC_WFFT(512)
          SPEC_FFT(1024)
          SPEC_FFT(2048)
          SPEC_FFT(192)
          SPEC_FFT(384)
          SPEC_FFT(768)
          SPEC_FFT(1536)
          default:
            fe_spectrum_dft<<<(unsigned)n_rows, 128, N * sizeof(float), st>>>(d_pcm, du, dr, n_rows, H, base.window_advance,
                                                                             base.emph, base.copy_borders, win, tw, N,
                                                                             mod.magnitude, mod.log, o);
        }
#undef SPEC_FFT
#undef SPEC_WFFT
        break;
      }
      case M_MEL:
        fe_mel<<<grid1(ne, 256), 256, 0, st>>>(s0, sdim, n_rows, MelTable{mod.d_a->as<float>(), mod.d_b->as<int>(), mod.d_c->as<float>()},
                                               mod.dim, mod.root, o);
        break;
      case M_POWER:
        fe_power<<<grid1(n_rows, 128), 128, 0, st>>>(s0, sdim, n_rows, 0, o);
        break;
      case M_MEL_POWER:
        fe_power<<<grid1(n_rows, 128), 128, 0, st>>>(s0, sdim, n_rows, 1, o);
        break;
      case M_DCT:
        fe_dct<<<grid1(ne, 256), 256, 0, st>>>(s0, sdim, n_rows, mod.d_a->as<float>(), mod.dim, o);
        break;
      case M_DELTA:
        fe_delta<<<grid1(ne, 256), 256, 0, st>>>(s0, mod.dim, n_rows, dr, du, H, mod.width, mod.norm, o);
        break;
      case M_MERGE: {
        int col = 0;
        for (int s : mod.src) {
          const int sd = fe.mods[s].dim;
          fe_copy<<<grid1(n_rows * sd, 256), 256, 0, st>>>(buf[s]->as<double>(), sd, n_rows, dr, du, H, 0, o, mod.dim, col);
          ctx->launches++;
          col += sd;
        }
        ctx->launches--;
        break;
      }
      case M_CONCAT: {
        int col = 0;
        for (int k = -mod.left; k <= mod.right; k++) {
          fe_copy<<<grid1(n_rows * sdim, 256), 256, 0, st>>>(s0, sdim, n_rows, dr, du, H, k, o, mod.dim, col);
          ctx->launches++;
          col += sdim;
        }
        ctx->launches--;
        break;
      }
      case M_NORMALIZATION:
        fe_norm<<<grid1(ne, 256), 256, 0, st>>>(s0, mod.dim, n_rows, mod.d_a->as<float>(), mod.d_b->as<float>(), o);
        break;
      case M_LIN_TRANSFORM:
        fe_lintrans<<<grid1(ne, 256), 256, 0, st>>>(s0, sdim, n_rows, mod.d_a->as<float>(), mod.matrix_defined ? 1 : 0,
                                                    mod.d_b->as<float>(), mod.bias_defined ? 1 : 0, mod.dim, o);
        break;
      case M_MEAN_SUBTRACTOR:
        fe_meansub<<<grid1(ne, 256), 256, 0, st>>>(s0, mod.dim, n_rows, dr, du, H, mod.left, mod.right, o);
        break;
      case M_SR_NORM:
        fe_srnorm<<<grid1(ne, 256), 256, 0, st>>>(s0, sdim, mod.frame_dim, mod.out_frames, n_rows, mod.d_a->as<float>(),
                                                  mod.d_b->as<int>(), o);
        break;
      case M_QUANTEQ:
        fe_quanteq<<<grid1(ne, 256), 256, 0, st>>>(s0, mod.dim, n_rows, mod.d_a->as<float>(),
                                                   (!mod.q_alpha.empty() && !mod.q_gamma.empty() && !mod.q_max.empty()) ? 1 : 0, o);
        break;
      case M_VTLN:
        fe_vtln<<<grid1(ne, 256), 256, 0, st>>>(s0, mod.dim, n_rows, mod.d_a->as<float>(), mod.d_b->as<int>(), mod.d_c->as<float>(),
                                                mod.sinc_rad, o);
        break;
      default:
        throw Error(AKUGPU_E_CONFIG, "module '" + mod.name + "' cannot be evaluated here");
    }
    AKU_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (target == 0 && !pre_base) throw Error(AKUGPU_E_CONFIG, "the raw audiofile module output is not materialised on the GPU");
  const int dim = fe.mods[target].dim;
  if (g_x >= 0) {
    const Module &D1 = fe.mods[g_d1], &D2 = fe.mods[g_d2];
    const int xd = fe.mods[g_x].dim;
    const size_t tile_smem = ((size_t)(DELTA_TILE_ROWS + 2 * (D1.width + D2.width)) + (DELTA_TILE_ROWS + 2 * D2.width)) * xd * sizeof(double);
    static const bool untiled = getenv("AKUGPU_FE_DELTA_UNTILED") != nullptr;      // the element-per-thread kernel, for comparisons
    if (tile_smem <= 160 * 1024 && !untiled) {
      const unsigned g = grid1(n_rows, DELTA_TILE_ROWS);
      if (out_f64) {
        ensure_dynamic_smem(ctx, (const void *)fe_delta2_merge_tiled<double>, tile_smem);
        fe_delta2_merge_tiled<double><<<g, 256, tile_smem, st>>>(buf[g_x]->as<double>(), xd, n_rows, du, dr, H, D1.width, D1.norm, D2.width,
                                                                 D2.norm, (double *)d_out + out_row_base * dim);
      } else {
        ensure_dynamic_smem(ctx, (const void *)fe_delta2_merge_tiled<float>, tile_smem);
        fe_delta2_merge_tiled<float><<<g, 256, tile_smem, st>>>(buf[g_x]->as<double>(), xd, n_rows, du, dr, H, D1.width, D1.norm, D2.width,
                                                                D2.norm, (float *)d_out + out_row_base * dim);
      }
    } else if (out_f64)
      fe_delta2_merge<double><<<grid1(n_rows * xd, 256), 256, 0, st>>>(buf[g_x]->as<double>(), xd, n_rows, du, dr, H, D1.width, D1.norm,
                                                                      D2.width, D2.norm, (double *)d_out + out_row_base * dim);
    else
      fe_delta2_merge<float><<<grid1(n_rows * xd, 256), 256, 0, st>>>(buf[g_x]->as<double>(), xd, n_rows, du, dr, H, D1.width, D1.norm,
                                                                     D2.width, D2.norm, (float *)d_out + out_row_base * dim);
  } else if (out_f64)
    fe_gather<double><<<grid1(n_rows * dim, 256), 256, 0, st>>>(buf[target]->as<double>(), dim, du, dr, n_rows, H,
                                                                (double *)d_out + out_row_base * dim);
  else
    fe_gather<float><<<grid1(n_rows * dim, 256), 256, 0, st>>>(buf[target]->as<double>(), dim, du, dr, n_rows, H,
                                                               (float *)d_out + out_row_base * dim);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace

void frontend_run_range(akugpu_ctx *ctx, const void *d_pcm, int64_t n_samples, int start, int end, int target,
                        void *d_out, int out_f64)
{
  const int64_t nf = frontend_num_frames(ctx->fe, n_samples);
  if (nf <= 0) throw Error(AKUGPU_E_ARG, "audio shorter than frame");
  std::vector<UttDesc> utts(1);
  UttDesc &u = utts[0];
  u.pcm_off = 0; u.n_samples = n_samples; u.row_off = 0; u.out_off = 0;
  u.n_frames = (int)nf; u.start = start; u.n_rows_out = end - start; u.pad = 0;
  run_graph(ctx, d_pcm, utts, target, d_out, out_f64, 0);
}

void frontend_run_batch(akugpu_ctx *ctx, const void *d_pcm, const std::vector<int64_t> &utt_off,
                        const std::vector<int64_t> &frame_off, void *d_out, int out_f64)
{
  const int n = (int)utt_off.size() - 1;
  const int64_t max_rows = 262144;   // rows per pass: bounds the module buffers (spectrum: rows*129*8 B)
  int u0 = 0;
  while (u0 < n) {
    std::vector<UttDesc> utts;
    int64_t rows = 0;
    int u1 = u0;
    while (u1 < n && (u1 == u0 || rows + (frame_off[u1 + 1] - frame_off[u1]) <= max_rows)) {
      UttDesc u;
      u.pcm_off = utt_off[u1]; u.n_samples = utt_off[u1 + 1] - utt_off[u1]; u.row_off = 0;
      u.out_off = frame_off[u1] - frame_off[u0];
      u.n_frames = (int)(frame_off[u1 + 1] - frame_off[u1]); u.start = 0; u.n_rows_out = u.n_frames; u.pad = 0;
      rows += u.n_frames;
      utts.push_back(u);
      u1++;
    }
    run_graph(ctx, d_pcm, utts, -1, d_out, out_f64, frame_off[u0]);
    u0 = u1;
  }
}

}  // namespace akugpu
