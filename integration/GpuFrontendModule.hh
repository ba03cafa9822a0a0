// integration/GpuFrontendModule.hh -- the binding a maintainer of AaltoASR adds to aku/ to run the feature chain on
// the GPU library from INSIDE aku::FeatureGenerator.
//
// It is a base module (aku/BaseFeaModule.hh:10-27) of type "gpu_frontend": its configuration names a feature
// configuration for libakugpu.so (`config <path>`); opening a file computes every frame of the utterance in one
// library call, and FeatureModule::at() (aku/FeatureModules.cc:102-158) is served from that matrix.  Reference modules
// placed after it in the same FeatureGenerator (deltas, transforms, ...) keep working unchanged, and so does
// everything that calls FeatureGenerator::generate().  One line registers it in the type chain of
// aku/FeatureGenerator.cc:145-178:
//
//     else if (type == GpuFrontendModule::type_str()) module = new GpuFrontendModule();
//
// This header compiles against the reference's own headers (aku/BaseFeaModule.hh) and include/akugpu.h;
// tests/test_abi.py builds it with the reference library from oracle/_ref and drives it, with reference modules
// downstream, against a fake of the C ABI.  Nothing here is used by the library itself.
#ifndef GPUFRONTENDMODULE_HH
#define GPUFRONTENDMODULE_HH

#include <stdio.h>
#include <map>
#include <string>
#include <vector>
#include "BaseFeaModule.hh"
#include "ModuleConfig.hh"
#include "akugpu.hh"        // aaltoasr_b200/csrc/host: read_audio, check, the C ABI
#include "GpuHmmSetHook.hh" // publish_utterance: lets the HmmSet hook score the whole utterance in one call

namespace aku {

class GpuFrontendModule : public BaseFeaModule {
public:
  explicit GpuFrontendModule(int device = 0) : m_engine(device), m_frames(0) { m_type_str = type_str(); }
  static const char *type_str() { return "gpu_frontend"; }

  // FeatureGenerator::open(filename) (aku/FeatureGenerator.cc:31-52): the whole utterance is computed here.
  virtual void set_fname(const char *fname) {
    int rate = 0;
    akugpu::read_audio(fname, sample_rate(), m_raw, m_pcm, rate, m_big_endian);
    open_pcm(rate);
  }
  // FeatureGenerator::open(FILE*) / open_fd (aku/FeatureGenerator.cc:55-83): the stream is read to its end.
  virtual void set_file(FILE *fp, bool /*stream*/ = false) {
    std::vector<unsigned char> bytes;
    unsigned char buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, fp)) > 0) bytes.insert(bytes.end(), buf, buf + k);
    int rate = 0;
    akugpu::parse_audio(bytes, "<stream>", sample_rate(), m_raw, m_pcm, rate, m_big_endian);
    open_pcm(rate);
  }
  virtual void discard_file(void) { m_pcm.clear(); m_feats.clear(); m_frames = 0; akugpu_hook::publish_utterance(NULL, 0, 0); }
  virtual ~GpuFrontendModule() { if (akugpu_hook::current_utterance().feats == m_feats.data()) akugpu_hook::publish_utterance(NULL, 0, 0); }
  virtual bool eof(int frame) { return frame >= m_frames; }
  virtual int sample_rate(void) { return akugpu_frontend_sample_rate(m_engine.ctx()); }
  virtual float frame_rate(void) { return akugpu_frontend_frame_rate(m_engine.ctx()); }
  virtual int last_frame(void) { return m_frames - 1; }

  // Speaker adaptation (SpeakerConfig -> FeatureModule::set_parameters, aku/FeatureModule.hh:107): a key
  // "<module>.<parameter>" of this module's block addresses <parameter> of <module> in the GPU chain, e.g.
  //     gpu { cmllr.matrix 1 0 ...   cmllr.bias 0 0 ... }      for a lin_transform named cmllr in the GPU configuration.
  // ModuleConfig has no key enumeration; write() lists every pair (aku/ModuleConfig.cc:205-223).
  virtual void set_parameters(const ModuleConfig &params) {
    FILE *tmp = tmpfile();
    if (!tmp) throw std::string("GpuFrontendModule: tmpfile() failed");
    params.write(tmp, 0);
    rewind(tmp);
    std::map<std::string, std::string> per_module;
    std::string line;
    int ch;
    while ((ch = fgetc(tmp)) != EOF) {
      if (ch != '\n') { line += (char)ch; continue; }
      size_t b = line.find_first_not_of(" \t");
      if (b != std::string::npos && line[b] != '{' && line[b] != '}') {
        size_t dot = line.find('.', b), sp = line.find(' ', b);
        if (dot == std::string::npos || (sp != std::string::npos && dot > sp)) {
          fclose(tmp);
          throw std::string("GpuFrontendModule: parameter keys are <module>.<parameter>: ") + line.substr(b);
        }
        per_module[line.substr(b, dot - b)] += line.substr(dot + 1) + "\n";
      }
      line.clear();
    }
    fclose(tmp);
    for (std::map<std::string, std::string>::const_iterator it = per_module.begin(); it != per_module.end(); ++it)
      akugpu::check(m_engine.ctx(), akugpu_frontend_set_parameters(m_engine.ctx(), it->first.c_str(), it->second.c_str()));
    if (!m_pcm.empty()) compute();
  }

private:
  virtual void set_module_config(const ModuleConfig &config) {
    if (!config.get("config", m_config_path)) throw std::string("GpuFrontendModule: Must set config (a feature configuration file)");
    akugpu::check(m_engine.ctx(), akugpu_frontend_load_config(m_engine.ctx(), m_config_path.c_str()));
    m_dim = akugpu_frontend_dim(m_engine.ctx());
    akugpu::config_audio_format(m_config_path, m_raw, m_big_endian);      // `raw` / `endian` of the GPU chain's audiofile module
    m_own_offset_left = 0;
    m_own_offset_right = 0;
  }
  virtual void get_module_config(ModuleConfig &config) { config.set("config", m_config_path); }
  virtual void reset_module() {}
  // FeatureModule::at() asks for the frames missing from its ring buffer.  Frames inside the file come from the
  // matrix; frames outside it (the reference replicates the first / last window at the INPUT of the chain, so delta-type
  // outputs there are not copies of the border frame) are computed exactly by the library's frame-range call.
  virtual void generate(int frame) {
    FeatureVec target = m_buffer[frame];
    if (frame >= 0 && frame < m_frames) {
      for (int i = 0; i < m_dim; i++) target[i] = m_feats[(size_t)frame * m_dim + i];
      return;
    }
    std::vector<double> row(m_dim);
    int dim = 0;
    akugpu::check(m_engine.ctx(), akugpu_features_range(m_engine.ctx(), m_pcm.data(), (int64_t)m_pcm.size(), frame, frame + 1, NULL,
                                                        row.data(), 1, &dim));
    for (int i = 0; i < m_dim; i++) target[i] = row[i];
  }
  void open_pcm(int rate) {
    if (rate != sample_rate()) {       // aku/FeatureModules.cc:254-261
      char msg[256];
      snprintf(msg, sizeof msg, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree.", rate, sample_rate());
      throw std::string(msg);
    }
    compute();
  }
  void compute() {
    int64_t uo[2] = {0, (int64_t)m_pcm.size()}, fo[2] = {0, 0};
    akugpu::check(m_engine.ctx(), akugpu_features(m_engine.ctx(), NULL, uo, 1, NULL, 1, fo));
    m_frames = (int)fo[1];
    m_feats.resize((size_t)m_frames * m_dim);
    if (m_frames > 0) akugpu::check(m_engine.ctx(), akugpu_features(m_engine.ctx(), m_pcm.data(), uo, 1, m_feats.data(), 1, fo));
    akugpu_hook::publish_utterance(m_feats.data(), m_frames, m_dim);
    reset();                           // the ring buffer of at() holds frames of the previous file / parameters
  }

  akugpu::Engine m_engine;
  std::string m_config_path;
  std::vector<int16_t> m_pcm;
  std::vector<double> m_feats;
  int m_frames;
  bool m_raw = false, m_big_endian = false;
};

}  // namespace aku

#endif
