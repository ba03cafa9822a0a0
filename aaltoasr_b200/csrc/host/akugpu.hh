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

#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <algorithm>
#include <map>
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
// Headerless data is little-endian unless the audiofile module's configuration says `endian big` (aku/AudioReader.cc:94-108).
inline void parse_audio(const std::vector<unsigned char> &b, const std::string &path, int config_rate, bool force_raw,
                        std::vector<int16_t> &pcm, int &rate, bool raw_big_endian = false)
{
  bool is_raw = true;
  size_t off = 0, len = b.size();
  rate = config_rate;
  if (!force_raw && b.size() >= 12 && !memcmp(&b[0], "RIFF", 4) && !memcmp(&b[8], "WAVE", 4)) {
    is_raw = false;
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
        // WAVE_FORMAT_EXTENSIBLE: the sample format is the first field of the SubFormat GUID (libsndfile reads these too)
        if (fmt == 0xFFFE && n >= 40 && p + 8 + 26 <= b.size()) fmt = b[p + 8 + 24] | (b[p + 8 + 25] << 8);
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
  const int lo = (is_raw && raw_big_endian) ? 1 : 0;
  for (size_t i = 0; i < pcm.size(); i++) pcm[i] = (int16_t)(b[off + 2 * i + lo] | (b[off + 2 * i + 1 - lo] << 8));
}
inline void read_audio(const std::string &path, int config_rate, bool force_raw, std::vector<int16_t> &pcm, int &rate,
                       bool raw_big_endian = false)
{
  FILE *fp = path == "-" ? stdin : fopen(path.c_str(), "rb");    // "-" = standard input, like io::Stream (aku/io.cc:52-60)
  if (!fp) throw std::string("AudioReader::open(): could not open file:") + path;
  std::vector<unsigned char> b;
  unsigned char buf[65536];
  size_t k;
  while ((k = fread(buf, 1, sizeof buf, fp)) > 0) b.insert(b.end(), buf, buf + k);
  if (fp != stdin) fclose(fp);
  parse_audio(b, path, config_rate, force_raw, pcm, rate, raw_big_endian);
}

// The two keys of the audiofile module that concern the FILE rather than the signal processing (AudioFileModule::
// set_module_config, aku/FeatureModules.cc:345-356): `raw 1` = never look for a header, `endian big|little` = byte
// order of headerless data.  The library parses the rest of the configuration; the container is the host's business.
inline void config_audio_format(const std::string &cfg_path, bool &raw, bool &big_endian)
{
  raw = false; big_endian = false;
  FILE *fp = fopen(cfg_path.c_str(), "r");
  if (!fp) return;                     // the library reports the missing file
  char line[4096];
  int depth = 0, module = 0;
  while (fgets(line, sizeof line, fp)) {
    char key[64] = "", val[64] = "";
    const int n = sscanf(line, " %63s %63s", key, val);
    if (n < 1) continue;
    if (!strcmp(key, "{")) { depth++; continue; }
    if (!strcmp(key, "}")) { depth--; if (depth == 0 && module == 1) break; continue; }      // first module only: the base module
    if (depth == 0 && !strcmp(key, "module")) { module++; continue; }
    if (depth == 1 && module == 1 && n == 2) {
      if (!strcmp(key, "raw")) raw = atoi(val) != 0;
      if (!strcmp(key, "endian")) big_endian = !strcmp(val, "big");
    }
  }
  fclose(fp);
}

class FeatureGenerator {
public:
  explicit FeatureGenerator(Engine &e) : m_e(e), m_frames(0), m_dim(0), m_eof(false) {}
  void load_configuration(const std::string &path) {
    check(m_e.ctx(), akugpu_frontend_load_config(m_e.ctx(), path.c_str()));
    m_dim = akugpu_frontend_dim(m_e.ctx());
    config_audio_format(path, m_cfg_raw, m_cfg_big_endian);
  }
  bool config_raw() const { return m_cfg_raw; }                   // `raw 1` / `endian big` of the audiofile module
  bool config_big_endian() const { return m_cfg_big_endian; }
  // A `pre` base module reads stored features (int32 dim + float32 rows, feacat -H --raw-output) instead of audio.
  void open_pre(const std::string &filename) {
    FILE *fp = filename == "-" ? stdin : fopen(filename.c_str(), "rb");
    if (!fp) throw std::string("could not open file ") + filename;
    // PreModule::set_file (aku/FeatureModules.cc:602-627): a `legacy_file 1` configuration stores the dimension in ONE
    // byte, otherwise it is an int; either way it must equal the configured `dim` of the module
    int dim = 0;
    bool got;
    if (akugpu_frontend_pre_legacy(m_e.ctx()) == 1) { char d = 0; got = fread(&d, 1, 1, fp) == 1; dim = d; }
    else got = fread(&dim, sizeof(int), 1, fp) == 1;
    if (!got) { if (fp != stdin) fclose(fp); throw std::string("PreModule: Could not read the file."); }
    if (dim != akugpu_frontend_base_dim(m_e.ctx())) { if (fp != stdin) fclose(fp); throw std::string("PreModule: The file has invalid dimension"); }
    m_rows.clear();
    std::vector<float> buf(4096);
    size_t k;
    while ((k = fread(buf.data(), sizeof(float), buf.size(), fp)) > 0) m_rows.insert(m_rows.end(), buf.begin(), buf.begin() + k);
    if (fp != stdin) fclose(fp);
    m_rows.resize(m_rows.size() / (size_t)dim * (size_t)dim);      // a trailing partial row is never read (PreModule::generate :733-745)
    m_pre_dim = dim;
    int64_t ro[2] = {0, (int64_t)(m_rows.size() / dim)}, fo[2] = {0, 0};
    check(m_e.ctx(), akugpu_features_pre(m_e.ctx(), NULL, ro, 1, NULL, 1, fo));
    m_frames = (int)fo[1];
    m_feats.resize((size_t)m_frames * m_dim);
    check(m_e.ctx(), akugpu_features_pre(m_e.ctx(), m_rows.data(), ro, 1, m_feats.data(), 1, fo));
    m_eof = false;
    m_pcm.clear();
  }
  void open(const std::string &filename) {
    if (akugpu_frontend_base_is_pre(m_e.ctx()) == 1) { open_pre(filename); return; }
    m_rows.clear();
    int rate = 0;
    read_audio(filename, sample_rate(), m_cfg_raw, m_pcm, rate, m_cfg_big_endian);
    if (rate != sample_rate()) {     // aku/FeatureModules.cc:254-261
      char msg[256];
      snprintf(msg, sizeof msg, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree.", rate, sample_rate());
      throw std::string(msg);
    }
    open_pcm();
  }
  // FeatureGenerator::open(FILE*, dont_fclose, stream) / open_fd (aku/FeatureGenerator.cc:55-83): the stream is read
  // to its end here (the library computes whole utterances); open_fd ignores its raw flag like the reference's does.
  void open(FILE *file, bool dont_fclose = false, bool /*stream*/ = false) {
    std::vector<unsigned char> b;
    unsigned char buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, file)) > 0) b.insert(b.end(), buf, buf + k);
    if (!dont_fclose) fclose(file);
    m_rows.clear();
    int rate = 0;
    parse_audio(b, "<stream>", sample_rate(), m_cfg_raw, m_pcm, rate, m_cfg_big_endian);
    if (rate != sample_rate()) {
      char msg[256];
      snprintf(msg, sizeof msg, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree.", rate, sample_rate());
      throw std::string(msg);
    }
    open_pcm();
  }
  void open_fd(int fd, bool /*raw_audio*/ = false) {
    FILE *file = fdopen(fd, "rb");
    if (file == NULL) throw std::string("could not open fd ") + ": " + strerror(errno);
    open(file, false, false);
  }
  void open_pcm(const std::vector<int16_t> &pcm) { m_pcm = pcm; open_pcm(); }
  void close() { m_pcm.clear(); m_feats.clear(); m_frames = 0; }
  // Feature vector of `frame` (doubles, like aku::FeatureVec); valid until the next generate().
  const double *generate(int frame) {
    m_eof = frame >= m_frames;
    if (frame >= 0 && frame < m_frames) return &m_feats[(size_t)frame * m_dim];
    m_tmp.resize(m_dim);
    int dim = 0;
    if (!m_rows.empty())
      check(m_e.ctx(), akugpu_features_pre_range(m_e.ctx(), m_rows.data(), (int64_t)(m_rows.size() / m_pre_dim), frame, frame + 1,
                                                 NULL, m_tmp.data(), 1, &dim));
    else
      check(m_e.ctx(), akugpu_features_range(m_e.ctx(), m_pcm.data(), (int64_t)m_pcm.size(), frame, frame + 1, NULL,
                                             m_tmp.data(), 1, &dim));
    return m_tmp.data();
  }
  // Frames [start, end) in one GPU call, frames outside the file as the reference's border handling gives them
  // (feacat --start-frame / --end-frame, aku/feacat.cc:96-121).  Re-run after SpeakerConfig changes parameters.
  void generate_range(int start, int end, std::vector<double> &out) {
    int dim = 0;
    out.resize((size_t)std::max(0, end - start) * m_dim);
    if (end <= start) return;
    if (!m_rows.empty())
      check(m_e.ctx(), akugpu_features_pre_range(m_e.ctx(), m_rows.data(), (int64_t)(m_rows.size() / m_pre_dim), start, end,
                                                 NULL, out.data(), 1, &dim));
    else
      check(m_e.ctx(), akugpu_features_range(m_e.ctx(), m_pcm.data(), (int64_t)m_pcm.size(), start, end, NULL,
                                             out.data(), 1, &dim));
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
  std::vector<float> m_rows;       // stored features of a `pre` base module
  int m_pre_dim = 0;
  std::vector<double> m_feats, m_tmp;
  int m_frames, m_dim;
  bool m_eof;
  bool m_cfg_raw = false, m_cfg_big_endian = false;
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
    score(feats, n_frames, m_lik.data());
    m_cur = -1;
  }
  // The in-process decoder feed of decoder/decode-stream.cc:191-207: (float) log(max(likelihood, tiny)) for every state
  // of the frames [first, first + n) of `feats` (n = 1: the per-frame loop), ready for Toolbox::set_one_frame.
  void log_probs(const double *feats, int64_t n_frames, std::vector<float> &out, double tiny = 1e-30) {
    out.resize((size_t)n_frames * m_S);
    check(m_e.ctx(), akugpu_gmm_logprobs(m_e.ctx(), feats, 1, n_frames, m_prec, tiny, out.data()));
  }
  void reset_cache() { m_cur = -1; m_one_valid = false; }
  void precompute_likelihoods(int frame) { m_cur = frame; }
  // Linear likelihood floored at 1e-50, as aku::HmmSet::state_likelihood returns (aku/HmmSet.cc:470-481).
  double state_likelihood(int state) const { return m_lik[(size_t)m_cur * m_S + state]; }
  double state_likelihood(int state, int frame) const { return m_lik[(size_t)frame * m_S + state]; }
  // The reference's own signatures for callers that hand in one feature vector at a time (aku/HmmSet.hh:309-321):
  // precompute_likelihoods(f) scores every state for `fea` in one GPU call (F = 1); state_likelihood(s, f) returns the
  // cached value and, like pdf_likelihood (aku/HmmSet.cc:470-481), fills the cache first when reset_cache() emptied it
  // -- the cache is keyed by reset_cache(), not by the vector, exactly as there.
  void precompute_likelihoods(const double *fea) {
    m_one.resize(m_S);
    score(fea, 1, m_one.data());
    m_one_valid = true;
  }
  double state_likelihood(int state, const double *fea) {
    if (!m_one_valid) precompute_likelihoods(fea);
    return m_one[state];
  }
private:
  void score(const double *feats, int64_t n_frames, double *lik) {
    if (m_prec == AKUGPU_F64) {
      check(m_e.ctx(), akugpu_gmm_score(m_e.ctx(), feats, 1, n_frames, AKUGPU_F64, lik));
    } else {
      std::vector<float> ll((size_t)n_frames * m_S);
      check(m_e.ctx(), akugpu_gmm_score(m_e.ctx(), feats, 1, n_frames, AKUGPU_F32, ll.data()));
      for (size_t i = 0; i < ll.size(); i++) { double v = exp((double)ll[i]); lik[i] = v < 1e-50 ? 1e-50 : v; }
    }
  }
  Engine &m_e;
  int m_prec, m_S, m_cur;
  bool m_one_valid = false;
  std::vector<double> m_lik, m_one;
};

// Speaker / utterance adaptation parameters for FEATURE modules (aku::SpeakerConfig, aku/SpeakerConfig.cc:20-381;
// phone_probs -S file.spkc, aku/phone_probs.cc:93-94,192-197).  File format:
//     speaker <id | default>            (or: utterance <id | default>)
//     {
//       [feature] <module name>
//       {
//         key value ...                 e.g. matrix / bias of a lin_transform (CMLLR), mean / scale of a normalization
//       }
//     }
// set_speaker() / set_utterance() push the stored parameters through FeatureModule::set_parameters
// (akugpu_frontend_set_parameters).  A speaker's `model cmllr` entry (model-level constrained MLLR, aku/ModelModules.hh)
// is applied through akugpu_model_set_cmllr when it is the global transform (unitmode UNIT_NO) and through
// akugpu_model_set_cmllr_units for the regression-class modes (UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN).  As in the reference, a speaker without the entry keeps the previous speaker's transform.
// The stream decoder's session (decoder/decode-stream.cc:150-207: one acoustic model for the life of the stream, one
// scoring call per frame -> Toolbox::set_one_frame).  While an object lives, the model sits in the shared memory of the
// SMs (akugpu_stream_open) and log_probs() is a message to the resident kernel: no launch, no copy -- the returned rows
// are the context's pinned memory, valid until its next call.  Any other call on the engine ends the kernel; the next
// log_probs() starts it again.
class StreamSession {
public:
  explicit StreamSession(Engine &e, double tiny = 1e-30, double idle_ms = 100.0) : m_e(e), m_tiny(tiny), m_S(akugpu_model_num_states(e.ctx())) {
    check(m_e.ctx(), akugpu_stream_open(m_e.ctx(), idle_ms));
  }
  ~StreamSession() { akugpu_stream_close(m_e.ctx()); }
  int num_states() const { return m_S; }
  // (float) log(max(likelihood, tiny)) of every state for n_frames (1..16) rows of host float features
  const float *log_probs(const float *feats, int n_frames = 1) {
    const float *rows = 0;
    check(m_e.ctx(), akugpu_stream_logprobs(m_e.ctx(), feats, n_frames, m_tiny, &rows));
    return rows;
  }
  // the reference's vector form: what decode-stream.cc hands to Toolbox::set_one_frame
  void log_probs(const std::vector<float> &fea, std::vector<float> &out) {
    const float *rows = log_probs(fea.data(), 1);
    out.assign(rows, rows + m_S);
  }
private:
  StreamSession(const StreamSession &);
  StreamSession &operator=(const StreamSession &);
  Engine &m_e;
  double m_tiny;
  int m_S;
};

class SpeakerConfig {
public:
  explicit SpeakerConfig(Engine &e) : m_e(e), m_default_speaker_set(false), m_default_utterance_set(false) {}
  void read_speaker_file(const std::string &path) {
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) throw std::string("could not open speaker configuration ") + path;
    std::vector<std::string> lines;
    char buf[1 << 16];
    std::string cur;
    while (fgets(buf, sizeof buf, fp)) {
      cur += buf;
      if (!cur.empty() && cur[cur.size() - 1] == '\n') { lines.push_back(clean(cur)); cur.clear(); }
    }
    if (!cur.empty()) lines.push_back(clean(cur));
    fclose(fp);
    size_t i = 0;
    int lineno = 0;
    auto next = [&](std::string &out) -> bool {       // next non-empty line
      while (i < lines.size()) { lineno++; out = lines[i++]; if (!out.empty()) return true; }
      return false;
    };
    std::string line;
    while (next(line)) {
      std::vector<std::string> f = split(line, 0);
      if (f.size() != 2 || (f[0] != "speaker" && f[0] != "utterance")) throw fmt("SpeakerConfig: Syntax error on line %d: ", lineno) + line;
      const bool is_speaker = f[0] == "speaker", is_default = f[1] == "default";
      ModuleMap *dst;
      if (is_speaker) {
        if (is_default && m_default_speaker_set) throw fmt("SpeakerConfig: Default speaker configuration already defined, redefinition on line %d: ", lineno) + line;
        if (is_default) { m_default_speaker_set = true; dst = &m_default_speaker; } else dst = &m_speakers[f[1]];
      } else {
        if (is_default && m_default_utterance_set) throw fmt("SpeakerConfig: Default utterance configuration already defined, redefinition on line %d: ", lineno) + line;
        if (is_default) { m_default_utterance_set = true; dst = &m_default_utterance; } else dst = &m_utterances[f[1]];
      }
      if (!next(line)) break;
      if (line != "{") throw std::string("'{' expected in speaker config file: ") + line;
      while (next(line)) {
        if (line == "}") break;
        std::vector<std::string> parts = split(line, 2);
        std::string ns = "feature", name = line;
        if (parts.size() == 2) {
          if (parts[0] != "model" && parts[0] != "feature") throw fmt("SpeakerConfig: Unknown module namespace at line %d", lineno);
          ns = parts[0]; name = parts[1];
        }
        if (ns == "model") {
          // ModelTransformer::get_new_module knows one module (aku/ModelModules.cc:12-18); it is loaded per speaker
          if (name != "cmllr") throw fmt("SpeakerConfig: error on line %d: ", lineno) + "unknown model module requested: " + name;
          if (!is_speaker)
            throw fmt("SpeakerConfig: error on line %d: ", lineno) + "model modules are loaded per speaker; an utterance-level entry is not supported";
        }
        // module parameters: a ModuleConfig block (aku/ModuleConfig.cc:166-203)
        if (!next(line)) throw std::string("SpeakerConfig: Failed reading module parameters around line ") + fmt("%d: ", lineno) + "unexpected end of module config file";
        if (line != "{") throw std::string("SpeakerConfig: Failed reading module parameters around line ") + fmt("%d: ", lineno) + "'{' expected in module config file: " + line;
        std::string text;
        while (true) {
          if (!next(line)) throw std::string("SpeakerConfig: Failed reading module parameters around line ") + fmt("%d: ", lineno) + "unexpected end of module config file";
          if (line == "}") break;
          text += line + "\n";
        }
        (*dst)[ns == "model" ? "model " + name : name] = text;
      }
    }
  }
  void set_speaker(const std::string &speaker_id) {
    if (!m_cur_utterance.empty()) set_utterance("");        // the utterance level is reset with the speaker
    if (speaker_id.empty()) {
      if (!m_default_speaker_set) throw std::string("SpeakerConfig: No speaker defined, needs a default speaker.");
      apply(m_default_speaker);
    } else {
      std::map<std::string, ModuleMap>::const_iterator it = m_speakers.find(speaker_id);
      if (it == m_speakers.end()) {
        if (!m_default_speaker_set) throw std::string("SpeakerConfig: Unknown speaker ") + speaker_id + ", and default speaker settings are missing.";
        it = m_speakers.insert(std::make_pair(speaker_id, m_default_speaker)).first;
      }
      apply(it->second);
    }
    m_cur_speaker = speaker_id;
  }
  void set_utterance(const std::string &utterance_id) {
    if (utterance_id.empty()) {
      if (!m_default_utterance_set) throw std::string("SpeakerConfig: Default utterance is required.");
      apply(m_default_utterance);
    } else {
      std::map<std::string, ModuleMap>::const_iterator it = m_utterances.find(utterance_id);
      if (it == m_utterances.end()) {
        if (!m_default_utterance_set) throw std::string("SpeakerConfig: Unknown utterance ") + utterance_id + ", and default utterance settings are missing.";
        it = m_utterances.insert(std::make_pair(utterance_id, m_default_utterance)).first;
      }
      apply(it->second);
    }
    m_cur_utterance = utterance_id;
  }
  const std::string &get_cur_speaker() const { return m_cur_speaker; }
  const std::string &get_cur_utterance() const { return m_cur_utterance; }
private:
  typedef std::map<std::string, std::string> ModuleMap;     // module name -> `key value` lines
  void apply(const ModuleMap &m) {
    for (ModuleMap::const_iterator it = m.begin(); it != m.end(); ++it) {
      if (it->first == "model cmllr") apply_cmllr(it->second);
      else check(m_e.ctx(), akugpu_frontend_set_parameters(m_e.ctx(), it->first.c_str(), it->second.c_str()));
    }
  }
  // ConstrainedMllr::set_parameters (aku/ModelModules.cc:62-95): `unitmode` and entries `w<i> [units...] <dim*(dim+1)
  // numbers>`, row-major [dim x (dim+1)], column 0 = bias; no `w` entry = no transform.
  void apply_cmllr(const std::string &text) {
    const int dim = akugpu_model_dim(m_e.ctx());
    if (dim <= 0) throw std::string("cmllr: a model must be loaded before the speaker configuration is applied");
    std::map<std::string, std::vector<std::string> > params;
    size_t pos = 0;
    while (pos < text.size()) {
      size_t e = text.find('\n', pos);
      if (e == std::string::npos) e = text.size();
      std::vector<std::string> f = split(text.substr(pos, e - pos), 0);
      if (!f.empty()) params[f[0]] = std::vector<std::string>(f.begin() + 1, f.end());
      pos = e + 1;
    }
    std::string um = "UNIT_NO";
    if (params.count("unitmode") && !params["unitmode"].empty()) um = params["unitmode"][0];
    const size_t n = (size_t)dim * (dim + 1);
    const size_t need = um == "UNIT_NO" ? n : n + 1;         // regression-class entries list their units in front of the matrix
    std::map<std::vector<std::string>, std::vector<double> > found;
    for (int i = 1;; i++) {
      std::map<std::string, std::vector<std::string> >::const_iterator it = params.find(fmt("w%d", i));
      if (it == params.end()) break;
      const std::vector<std::string> &parts = it->second;
      if (parts.size() < need) throw std::string("ERROR: not enough elements for matrix ") + fmt("w%d", i);
      std::vector<double> w(n);
      for (size_t k = 0; k < n; k++) {
        const std::string &t = parts[parts.size() - n + k];
        char *end = NULL;
        w[k] = (double)(float)strtod(t.c_str(), &end);       // str::str2float returns through a float (aku/str.cc:261-282)
        if (end == t.c_str() || *end) throw std::string("invalid value: ") + t;
      }
      found[std::vector<std::string>(parts.begin(), parts.end() - n)] = w;
    }
    if (um == "UNIT_NO") {
      if (found.size() > 1) throw std::string("ERROR: speaker can only contain one transform when UNIT_NO (global transform) is set");
      check(m_e.ctx(), akugpu_model_set_cmllr(m_e.ctx(), found.empty() ? NULL : found.begin()->second.data()));
      return;
    }
    // UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN: the library resolves the units to Gaussians (it has the model's phones and
    // mixtures) and applies the transforms in the reference's std::map order
    std::vector<std::string> units;
    std::vector<const char *> unit_ptrs;
    std::vector<double> W;
    for (std::map<std::vector<std::string>, std::vector<double> >::const_iterator it = found.begin(); it != found.end(); ++it) {
      std::string u;
      for (size_t k = 0; k < it->first.size(); k++) u += (k ? " " : "") + it->first[k];
      units.push_back(u);
      W.insert(W.end(), it->second.begin(), it->second.end());
    }
    for (size_t k = 0; k < units.size(); k++) unit_ptrs.push_back(units[k].c_str());
    check(m_e.ctx(), akugpu_model_set_cmllr_units(m_e.ctx(), um.c_str(), (int)units.size(), unit_ptrs.empty() ? NULL : unit_ptrs.data(),
                                                  W.empty() ? NULL : W.data()));
  }
  static std::string clean(const std::string &s) {
    size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
    return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
  }
  static std::vector<std::string> split(const std::string &s, int max_fields) {   // whitespace, at most max_fields (0 = all)
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
      while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
      if (i >= s.size()) break;
      if (max_fields && (int)out.size() == max_fields - 1) { out.push_back(s.substr(i)); break; }
      size_t j = i;
      while (j < s.size() && s[j] != ' ' && s[j] != '\t') j++;
      out.push_back(s.substr(i, j - i));
      i = j;
    }
    return out;
  }
  static std::string fmt(const char *f, int v) { char b[256]; snprintf(b, sizeof b, f, v); return b; }
  Engine &m_e;
  std::map<std::string, ModuleMap> m_speakers, m_utterances;
  ModuleMap m_default_speaker, m_default_utterance;
  bool m_default_speaker_set, m_default_utterance_set;
  std::string m_cur_speaker, m_cur_utterance;
};

// aku::Recipe (aku/Recipe.hh:15-118, aku/Recipe.cc:24-147): one utterance per line of key=value fields.  A key missing on
// a later line keeps the value of the line before (the reference never clears its map), also across batch borders;
// with num_batches > 1 the non-empty, non-comment lines are cut into contiguous parts, the first (lines % batches) of
// them one line longer (cluster_speakers = false, as phone_probs calls it, aku/phone_probs.cc:137-139).
class Recipe {
public:
  struct Info {
    std::string audio_path, lna_path, speaker_id, utterance_id;
    double start_time, end_time;
    Info() : start_time(0), end_time(0) {}
    bool operator<(const Info &i) const { return speaker_id < i.speaker_id; }      // aku/Recipe.hh:89-91
  };
  std::vector<Info> infos;
  void clear() { infos.clear(); }
  void read(const std::string &path, int num_batches = 0, int batch_index = 0) {
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) throw std::string("could not open file ") + path + ": " + strerror(errno);
    std::vector<std::string> lines;
    std::string cur;
    char buf[4096];
    while (fgets(buf, sizeof buf, fp)) {
      cur += buf;
      if (!cur.empty() && cur[cur.size() - 1] == '\n') { push_line(lines, cur); cur.clear(); }
    }
    if (!cur.empty()) push_line(lines, cur);
    fclose(fp);
    if (num_batches > 1 && (batch_index < 1 || batch_index > num_batches)) throw std::string("Invalid batch index");
    const size_t n = lines.size();
    size_t first = 0, last = n;
    if (num_batches > 1) {
      const size_t per = n / num_batches, rem = n % num_batches, b = (size_t)batch_index - 1;
      first = b * per + std::min(b, rem);
      last = first + per + (b < rem ? 1 : 0);
    }
    std::map<std::string, std::string> kv;
    for (size_t i = 0; i < std::min(last, n); i++) {
      size_t p = 0;
      const std::string &line = lines[i];
      while (p < line.size()) {
        while (p < line.size() && (line[p] == ' ' || line[p] == '\t')) p++;
        size_t q = p;
        while (q < line.size() && line[q] != ' ' && line[q] != '\t') q++;
        if (q > p) {
          const std::string f = line.substr(p, q - p);
          const size_t e = f.find('=');
          // str::split(field, "=", false): exactly two parts, i.e. one '=' that is not the last character
          if (e == std::string::npos || e + 1 == f.size() || f.find('=', e + 1) != std::string::npos)
            throw std::string("Invalid recipe line: ") + line;
          kv[f.substr(0, e)] = f.substr(e + 1);
        }
        p = q;
      }
      if (i < first) continue;
      Info info;
      info.audio_path = kv["audio"]; info.lna_path = kv["lna"]; info.speaker_id = kv["speaker"]; info.utterance_id = kv["utterance"];
      if (kv.count("start-time")) info.start_time = atof(kv["start-time"].c_str());
      if (kv.count("end-time")) info.end_time = atof(kv["end-time"].c_str());
      infos.push_back(info);
    }
  }
  void sort_infos() { std::stable_sort(infos.begin(), infos.end()); }               // aku/Recipe.hh:115-117
private:
  static void push_line(std::vector<std::string> &lines, const std::string &raw) {   // str::clean(&line, "\n\t ")
    size_t b = raw.find_first_not_of("\n\t \r"), e = raw.find_last_not_of("\n\t \r");
    if (b == std::string::npos) return;
    if (raw[b] == '#') return;
    lines.push_back(raw.substr(b, e - b + 1));
  }
};

// aku::PPToolbox (aku/PhoneProbsToolbox.hh:13-31, the class behind aku/swig/PPToolbox.i): one utterance in, one LNA
// stream out (5-byte header, 2-byte normalised codes -- lnabytes is fixed to 2 there, aku/PhoneProbsToolbox.cc:57,138),
// whole utterance per GPU call instead of the per-frame loop (:83-131,156-207).
class PPToolbox {
public:
  explicit PPToolbox(int device = 0, int precision = AKUGPU_F32) : m_e(device), m_gen(m_e), m_model(m_e), m_prec(precision) {}
  void read_models(const std::string &base) { m_model.read_all(base); }
  void read_configuration(const std::string &cfgname) { m_gen.load_configuration(cfgname); }
  void set_clustering(const std::string &clfile_name, double eval_minc, double eval_ming) {
    m_model.read_clustering(clfile_name);
    m_model.set_clustering_min_evals(eval_minc, eval_ming);
  }
  // Audio from an open descriptor (read to its end; raw_flag: headerless PCM16), LNA stream to out_fd.  Neither
  // descriptor is closed.
  void generate_to_fd(int in_fd, int out_fd, bool raw_flag) {
    check_dims();
    std::vector<unsigned char> b;
    unsigned char buf[65536];
    for (;;) {
      ssize_t k = read(in_fd, buf, sizeof buf);
      if (k < 0) { if (errno == EINTR) continue; throw std::string("could not read fd: ") + strerror(errno); }
      if (k == 0) break;
      b.insert(b.end(), buf, buf + k);
    }
    std::vector<int16_t> pcm;
    int rate = 0;
    parse_audio(b, "<fd>", m_gen.sample_rate(), raw_flag || m_gen.config_raw(), pcm, rate, m_gen.config_big_endian());
    emit(pcm, rate, out_fd);
  }
  void generate_from_file_to_fd(const std::string &input_name, int out_fd, bool /*raw_flag: unused by the reference too*/) {
    check_dims();
    std::vector<int16_t> pcm;
    int rate = 0;
    read_audio(input_name, m_gen.sample_rate(), m_gen.config_raw(), pcm, rate, m_gen.config_big_endian());
    emit(pcm, rate, out_fd);
  }
  void generate(const std::string &input_name, const std::string &output_name, bool raw_flag) {
    int out = open(output_name.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0664);
    if (out < 0) throw std::string("could not open ") + output_name + ": " + strerror(errno);
    try { generate_from_file_to_fd(input_name, out, raw_flag); }
    catch (...) { close(out); throw; }
    close(out);
  }
private:
  void check_dims() {
    if (m_model.dim() != m_gen.dim()) {     // aku/PhoneProbsToolbox.cc:65-70
      char msg[256];
      snprintf(msg, sizeof msg, "Gaussian dimension is %d but feature dimension is %d.", m_model.dim(), m_gen.dim());
      throw std::string(msg);
    }
  }
  void emit(const std::vector<int16_t> &pcm, int rate, int out_fd) {
    if (rate != m_gen.sample_rate()) {      // aku/FeatureModules.cc:254-261
      char msg[256];
      snprintf(msg, sizeof msg, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree.", rate, m_gen.sample_rate());
      throw std::string(msg);
    }
    const int S = m_model.num_states();
    int64_t uo[2] = {0, (int64_t)pcm.size()}, fo[2] = {0, 0};
    check(m_e.ctx(), akugpu_features(m_e.ctx(), NULL, uo, 1, NULL, 0, fo));     // frame count only
    std::vector<uint8_t> rec(5 + (size_t)fo[1] * S * 2);
    check(m_e.ctx(), akugpu_lna_header(S, 2, rec.data()));
    if (fo[1] > 0)
      check(m_e.ctx(), akugpu_phone_probs(m_e.ctx(), pcm.data(), uo, 1, m_prec, 2, 1, rec.data() + 5, fo, NULL));
    size_t done = 0;
    while (done < rec.size()) {
      ssize_t k = write(out_fd, rec.data() + done, rec.size() - done);
      if (k < 0) { if (errno == EINTR) continue; throw std::string("Write error"); }
      done += (size_t)k;
    }
  }
  Engine m_e;
  FeatureGenerator m_gen;
  HmmSet m_model;
  int m_prec;
};

}  // namespace akugpu
#endif
