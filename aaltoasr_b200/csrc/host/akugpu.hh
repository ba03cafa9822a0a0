// akugpu.hh -- C++ host-side adapters over the C ABI (include/akugpu.h), header only.
//
// Same names, argument meaning and error behaviour as the reference classes on the accelerated
// path, so that existing callers (aku/phone_probs.cc, decoder/decode-stream.cc style loops) compile
// against them with a type swap:
//   akugpu::FeatureGenerator  ~ aku::FeatureGenerator   aku/FeatureGenerator.hh:23-123
//   akugpu::HmmSet            ~ aku::HmmSet (scoring)   aku/HmmSet.hh:94-571
// Errors are thrown as std::string, as the reference does (e.g. aku/FeatureModules.cc:334).
//
// The reference pulls ONE frame per call through ring buffers and scores every Gaussian for that
// frame.  Here open() computes the whole utterance's features on the GPU once, and
// HmmSet::set_utterance() scores the whole utterance once; generate(f) / state_likelihood(s) are then
// served from those matrices.  Frames outside the file go through akugpu_features_range, which
// replicates the first/last window exactly like the reference's base module.
#ifndef AKUGPU_HOST_HH
#define AKUGPU_HOST_HH

#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../../include/akugpu.h"

namespace akugpu {

inline void check(akugpu_ctx *ctx, int rc) { if (rc != 0) throw std::string(akugpu_last_error(ctx)); }

// One context per host thread / GPU.
class Engine {
public:
  explicit Engine(int device = 0) : m_ctx(akugpu_create(device)) {
    if (!m_ctx) throw std::string(akugpu_last_error(NULL));
  }
  ~Engine() { akugpu_destroy(m_ctx); }
  akugpu_ctx *ctx() const { return m_ctx; }
private:
  Engine(const Engine &);
  Engine &operator=(const Engine &);
  akugpu_ctx *m_ctx;
};

// RIFF/WAVE PCM16 mono or headerless RAW PCM16, what aku::AudioReader accepts (aku/AudioReader.cc:86-155).
inline void read_audio(const std::string &path, int config_rate, bool force_raw, std::vector<int16_t> &pcm, int &rate)
{
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) throw std::string("AudioReader::open(): could not open file:") + path;
  std::vector<unsigned char> b;
  unsigned char buf[65536];
  size_t k;
  while ((k = fread(buf, 1, sizeof buf, fp)) > 0) b.insert(b.end(), buf, buf + k);
  fclose(fp);
  size_t off = 0, len = b.size();
  rate = config_rate;
  if (!force_raw && b.size() >= 12 && !memcmp(&b[0], "RIFF", 4) && !memcmp(&b[8], "WAVE", 4)) {
    size_t p = 12;
    int fmt = -1, ch = 0, bits = 0;
    bool found = false;
    while (p + 8 <= b.size()) {
      uint32_t n = b[p + 4] | (b[p + 5] << 8) | (b[p + 6] << 16) | ((uint32_t)b[p + 7] << 24);
      if (!memcmp(&b[p], "fmt ", 4) && p + 24 <= b.size()) {
        fmt = b[p + 8] | (b[p + 9] << 8);
        ch = b[p + 10] | (b[p + 11] << 8);
        rate = b[p + 12] | (b[p + 13] << 8) | (b[p + 14] << 16) | ((uint32_t)b[p + 15] << 24);
        bits = b[p + 22] | (b[p + 23] << 8);
      } else if (!memcmp(&b[p], "data", 4)) {
        off = p + 8; len = n; if (off + len > b.size()) len = b.size() - off;
        found = true;
        break;
      }
      p += 8 + n + (n & 1);
    }
    if (!found || fmt != 1 || bits != 16) throw std::string("AudioReader: sample format not PCM16: ") + path;
    if (ch != 1) throw std::string("AudioReader: sorry, audio files with multiple channels not supported");
  }
  pcm.resize(len / 2);
  for (size_t i = 0; i < pcm.size(); i++) pcm[i] = (int16_t)(b[off + 2 * i] | (b[off + 2 * i + 1] << 8));
}

class FeatureGenerator {
public:
  explicit FeatureGenerator(Engine &e) : m_e(e), m_frames(0), m_dim(0), m_eof(false) {}
  void load_configuration(const std::string &path) {
    check(m_e.ctx(), akugpu_frontend_load_config(m_e.ctx(), path.c_str()));
    m_dim = akugpu_frontend_dim(m_e.ctx());
  }
  void open(const std::string &filename) {
    int rate = 0;
    read_audio(filename, sample_rate(), false, m_pcm, rate);
    if (rate != sample_rate()) {     // aku/FeatureModules.cc:254-261
      char msg[256];
      snprintf(msg, sizeof msg, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree.", rate, sample_rate());
      throw std::string(msg);
    }
    open_pcm();
  }
  void open_pcm(const std::vector<int16_t> &pcm) { m_pcm = pcm; open_pcm(); }
  void close() { m_pcm.clear(); m_feats.clear(); m_frames = 0; }
  // Feature vector of `frame` (doubles, like aku::FeatureVec); valid until the next generate().
  const double *generate(int frame) {
    m_eof = frame >= m_frames;
    if (frame >= 0 && frame < m_frames) return &m_feats[(size_t)frame * m_dim];
    m_tmp.resize(m_dim);
    int dim = 0;
    check(m_e.ctx(), akugpu_features_range(m_e.ctx(), m_pcm.data(), (int64_t)m_pcm.size(), frame, frame + 1, NULL,
                                           m_tmp.data(), 1, &dim));
    return m_tmp.data();
  }
  const std::vector<double> &features() const { return m_feats; }   // all frames, [frames x dim]
  bool eof() const { return m_eof; }
  int last_frame() const { return m_frames - 1; }
  int num_frames() const { return m_frames; }
  int dim() const { return m_dim; }
  int sample_rate() const { return akugpu_frontend_sample_rate(m_e.ctx()); }
  float frame_rate() const { return akugpu_frontend_frame_rate(m_e.ctx()); }
  const std::vector<int16_t> &pcm() const { return m_pcm; }
private:
  void open_pcm() {
    int64_t uo[2] = {0, (int64_t)m_pcm.size()}, fo[2] = {0, 0};
    check(m_e.ctx(), akugpu_features(m_e.ctx(), NULL, uo, 1, NULL, 1, fo));
    m_frames = (int)fo[1];
    m_feats.resize((size_t)m_frames * m_dim);
    check(m_e.ctx(), akugpu_features(m_e.ctx(), m_pcm.data(), uo, 1, m_feats.data(), 1, fo));
    m_eof = false;
  }
  Engine &m_e;
  std::vector<int16_t> m_pcm;
  std::vector<double> m_feats, m_tmp;
  int m_frames, m_dim;
  bool m_eof;
};

class HmmSet {
public:
  explicit HmmSet(Engine &e, int precision = AKUGPU_F64) : m_e(e), m_prec(precision), m_S(0), m_cur(-1) {}
  void read_all(const std::string &base) { check(m_e.ctx(), akugpu_model_read(m_e.ctx(), base.c_str())); m_S = num_states(); }
  void read_files(const std::string &gk, const std::string &mc, const std::string &ph) {
    check(m_e.ctx(), akugpu_model_read_files(m_e.ctx(), gk.c_str(), mc.c_str(), ph.c_str()));
    m_S = num_states();
  }
  int num_states() const { return akugpu_model_num_states(m_e.ctx()); }
  int dim() const { return akugpu_model_dim(m_e.ctx()); }
  // Gaussian clustering approximation (aku/HmmSet.cc:1354-1366)
  void read_clustering(const std::string &filename) { check(m_e.ctx(), akugpu_model_read_clustering(m_e.ctx(), filename.c_str())); }
  void set_clustering_min_evals(double min_clusters = 1.0, double min_gaussians = 1.0) {
    check(m_e.ctx(), akugpu_model_set_clustering_min_evals(m_e.ctx(), min_clusters, min_gaussians));
  }
  // Scores every frame of an utterance at once; the per-frame calls below index into the result.
  void set_utterance(const double *feats, int64_t n_frames) {
    m_lik.resize((size_t)n_frames * m_S);
    if (m_prec == AKUGPU_F64) {
      check(m_e.ctx(), akugpu_gmm_score(m_e.ctx(), feats, 1, n_frames, AKUGPU_F64, m_lik.data()));
    } else {
      std::vector<float> ll((size_t)n_frames * m_S);
      check(m_e.ctx(), akugpu_gmm_score(m_e.ctx(), feats, 1, n_frames, AKUGPU_F32, ll.data()));
      for (size_t i = 0; i < ll.size(); i++) { double v = exp((double)ll[i]); m_lik[i] = v < 1e-50 ? 1e-50 : v; }
    }
    m_cur = -1;
  }
  // The in-process decoder feed of decoder/decode-stream.cc:191-207: (float) log(max(likelihood, tiny)) for every state
  // of the frames [first, first + n) of `feats` (n = 1: the per-frame loop), ready for Toolbox::set_one_frame.
  void log_probs(const double *feats, int64_t n_frames, std::vector<float> &out, double tiny = 1e-30) {
    out.resize((size_t)n_frames * m_S);
    check(m_e.ctx(), akugpu_gmm_logprobs(m_e.ctx(), feats, 1, n_frames, m_prec, tiny, out.data()));
  }
  void reset_cache() { m_cur = -1; }
  void precompute_likelihoods(int frame) { m_cur = frame; }
  // Linear likelihood floored at 1e-50, as aku::HmmSet::state_likelihood returns (aku/HmmSet.cc:470-481).
  double state_likelihood(int state) const { return m_lik[(size_t)m_cur * m_S + state]; }
  double state_likelihood(int state, int frame) const { return m_lik[(size_t)frame * m_S + state]; }
private:
  Engine &m_e;
  int m_prec, m_S, m_cur;
  std::vector<double> m_lik;
};

}  // namespace akugpu
#endif
