// oracle/ref_capi.cc -- C-callable view of the REFERENCE's own classes
// (aku::FeatureGenerator, aku::HmmSet), linked against the reference objects
// that oracle/build_ref.sh compiles from /root/reference/aku in place.
//
// TEST INFRASTRUCTURE ONLY.  It gives pytest (ctypes) double-precision access
// to what the reference computes, so that the numpy/C restatements in oracle/
// and the CUDA product can be checked against the real thing:
//   * features:    FeatureGenerator::generate(f)          (aku/FeatureGenerator.cc:267-273)
//   * likelihoods: HmmSet::precompute_likelihoods + state_likelihood
//                                                          (aku/HmmSet.cc:485-501, aku/HmmSet.hh:309)
// The LNA byte stream itself is produced by the literal aku/phone_probs.cc,
// built as oracle/_ref/ref_phone_probs.
#include <stdio.h>
#include <algorithm>
#include <string.h>
#include <string>
#include <exception>
#include "io.hh"
#include "FeatureGenerator.hh"
#include "FeatureModules.hh"
#include "HmmSet.hh"
#include "SpeakerConfig.hh"
#include "Recipe.hh"
#include "LnaReaderCircular.hh"   // decoder/src: the consumer of the LNA stream

using namespace aku;

static std::string g_err;
static int fail(const std::string &m) { g_err = m; return -1; }

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// Runs the reference feature pipeline on one audio file.
// Writes frames [start, end) (end<0: until the reference reports EOF) as doubles
// into out[max_frames*dim]; returns the number of frames written, dim in *dim_out.
long ref_features(const char *cfg_path, const char *audio_path, int start, int end,
                  double *out, long max_frames, int *dim_out, int *last_frame_out, float *frame_rate_out)
{
  try {
    FeatureGenerator gen;
    gen.load_configuration(io::Stream(cfg_path));
    gen.open(audio_path);
    int dim = gen.dim();
    if (dim_out) *dim_out = dim;
    if (frame_rate_out) *frame_rate_out = gen.frame_rate();
    long n = 0;
    for (int f = start; end < 0 || f < end; f++) {
      const FeatureVec fea = gen.generate(f);
      if (end < 0 && gen.eof()) break;
      if (n >= max_frames) break;
      for (int i = 0; i < dim; i++) out[n * dim + i] = fea[i];
      n++;
    }
    if (last_frame_out) *last_frame_out = gen.last_frame();
    gen.close();
    return n;
  } catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

// The same with a speaker configuration file applied first (aku::SpeakerConfig::read_speaker_file + set_speaker,
// aku/SpeakerConfig.cc:20-147,239-286; what phone_probs -S does per utterance): run-time module parameters such as a
// VTLN warp factor or a CMLLR lin_transform.
long ref_features_spk(const char *cfg_path, const char *audio_path, const char *spkc_path, const char *speaker,
                      int start, int end, double *out, long max_frames, int *dim_out)
{
  try {
    FeatureGenerator gen;
    gen.load_configuration(io::Stream(cfg_path));
    SpeakerConfig sc(gen);
    sc.read_speaker_file(io::Stream(spkc_path));
    gen.open(audio_path);
    sc.set_speaker(speaker);
    int dim = gen.dim();
    if (dim_out) *dim_out = dim;
    long n = 0;
    for (int f = start; end < 0 || f < end; f++) {
      const FeatureVec fea = gen.generate(f);
      if (end < 0 && gen.eof()) break;
      if (n >= max_frames) break;
      for (int i = 0; i < dim; i++) out[n * dim + i] = fea[i];
      n++;
    }
    gen.close();
    return n;
  } catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

// Output of an intermediate module (by name) for frames [start,end).
long ref_module_output(const char *cfg_path, const char *audio_path, const char *module_name,
                       int start, int end, double *out, long max_frames, int *dim_out)
{
  try {
    FeatureGenerator gen;
    gen.load_configuration(io::Stream(cfg_path));
    gen.open(audio_path);
    FeatureModule *m = gen.module(module_name);
    int dim = m->dim();
    if (dim_out) *dim_out = dim;
    long n = 0;
    for (int f = start; f < end && n < max_frames; f++) {
      const FeatureVec fea = m->at(f);
      for (int i = 0; i < dim; i++) out[n * dim + i] = fea[i];
      n++;
    }
    gen.close();
    return n;
  } catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

struct RefModel {
  HmmSet model;
  FeatureGenerator gen;          // no configuration: only `model` entries of a speaker file are used through it
  SpeakerConfig *spk;
  RefModel() : spk(NULL) {}
  ~RefModel() { delete spk; }    // resets the model transformations before the model goes
};

void *ref_model_open(const char *base)
{
  try {
    RefModel *m = new RefModel;
    m->model.read_all(base);
    return m;
  } catch (std::string &s) { fail(s); return NULL; }
  catch (std::exception &e) { fail(e.what()); return NULL; }
}
void ref_model_close(void *h) { delete (RefModel *)h; }
int ref_model_num_states(void *h) { return ((RefModel *)h)->model.num_states(); }
int ref_model_dim(void *h) { return ((RefModel *)h)->model.dim(); }
int ref_model_num_gaussians(void *h) { return ((RefModel *)h)->model.get_pool()->size(); }

// Gaussian clustering approximation: HmmSet::read_clustering + set_clustering_min_evals (aku/HmmSet.cc:1354-1366).
int ref_model_read_clustering(void *h, const char *path)
{
  try { ((RefModel *)h)->model.read_clustering(path); return 0; }
  catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}
int ref_model_set_clustering_min_evals(void *h, double min_clusters, double min_gaussians)
{
  try { ((RefModel *)h)->model.set_clustering_min_evals(min_clusters, min_gaussians); return 0; }
  catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

// aku::SpeakerConfig(gen, &model): read_speaker_file once, then set_speaker -- loads the speaker's model
// transformations (`model cmllr` entries; ModelTransformer, aku/SpeakerConfig.cc:236-285) into the HmmSet.
int ref_model_set_speaker(void *h, const char *spkc_path, const char *speaker)
{
  try {
    RefModel *m = (RefModel *)h;
    if (!m->spk) {
      m->spk = new SpeakerConfig(m->gen, &m->model);
      m->spk->read_speaker_file(io::Stream(spkc_path));
    }
    m->spk->set_speaker(speaker);
    return 0;
  } catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

// Linear state likelihoods (double, floored at 1e-50 by the reference) for F frames.
int ref_state_likelihoods(void *h, const double *feats, long F, int D, double *out /*[F*S]*/)
{
  try {
    HmmSet &model = ((RefModel *)h)->model;
    if (D != model.dim()) return fail("feature dim != model dim");
    int S = model.num_states();
    FeatureVec fv;
    Vector v(D);
    for (long f = 0; f < F; f++) {
      for (int i = 0; i < D; i++) v(i) = feats[f * D + i];
      FeatureVec fea(&v, D);
      model.reset_cache();
      model.precompute_likelihoods(fea);
      for (int s = 0; s < S; s++) out[f * S + s] = model.state_likelihood(s, fea);
    }
    return 0;
  } catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

// Per-Gaussian log-likelihoods (double) straight from the pool.
int ref_gaussian_loglik(void *h, const double *feats, long F, int D, double *out /*[F*G]*/)
{
  try {
    HmmSet &model = ((RefModel *)h)->model;
    PDFPool *pool = model.get_pool();
    int G = pool->size();
    Vector v(D);
    for (long f = 0; f < F; f++) {
      for (int i = 0; i < D; i++) v(i) = feats[f * D + i];
      for (int g = 0; g < G; g++) out[f * G + g] = pool->get_pdf(g)->compute_log_likelihood(v);
    }
    return 0;
  } catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

// aku::Recipe::read(file, num_batches, batch_index, cluster_speakers = false) as phone_probs calls it
// (aku/phone_probs.cc:137-142), optionally followed by sort_infos().  Writes one line per Info:
// audio|lna|speaker|utterance|start|end.  Returns the number of infos, -1 on error, -2 if `cap` is too small.
long ref_recipe_read(const char *path, int num_batches, int batch_index, int sort, char *out, long cap)
{
  try {
    Recipe recipe;
    recipe.read(io::Stream(path), num_batches, batch_index, false);
    if (sort) recipe.sort_infos();
    std::string text;
    for (size_t i = 0; i < recipe.infos.size(); i++) {
      const Recipe::Info &r = recipe.infos[i];
      char t[64];
      snprintf(t, sizeof t, "|%g|%g\n", (double)r.start_time, (double)r.end_time);
      text += r.audio_path + "|" + r.lna_path + "|" + r.speaker_id + "|" + r.utterance_id + t;
    }
    if ((long)text.size() + 1 > cap) return -2;
    memcpy(out, text.c_str(), text.size() + 1);
    return (long)recipe.infos.size();
  } catch (std::string &s) { return fail(s); }
  catch (std::exception &e) { return fail(e.what()); }
}

// Reads an LNA file with the DECODER's reader (decoder/src/LnaReaderCircular.cc:46-101,130-209): go_to(frame) for
// frames 0.. until it reports the end, log_prob(model) for every model.  Returns the number of frames; *n_models_out
// = number of models in the header.  out may be NULL (count only); with `backwards` the frames still inside the
// reader's circular buffer are revisited in descending order afterwards (the decoder steps back like that), values
// must agree.
long ref_lna_read(const char *path, int buf_frames, float *out, long max_frames, int *n_models_out, int backwards)
{
  LnaReaderCircular r;
  r.open_file(path, buf_frames);
  const int S = r.num_models();
  if (n_models_out) *n_models_out = S;
  long n = 0;
  while (r.go_to((int)n)) {
    if (out && n < max_frames)
      for (int s = 0; s < S; s++) out[n * S + s] = r.log_prob(s);
    n++;
  }
  if (backwards && out)
    for (long f = std::min(n, max_frames) - 1; f >= 0 && f > n - buf_frames; f -= 3) {
      if (!r.go_to((int)f)) { r.close(); return fail("go_to failed on a backward seek"); }
      for (int s = 0; s < S; s++)
        if (out[f * S + s] != r.log_prob(s)) { r.close(); return fail("backward seek returned different values"); }
    }
  r.close();
  return n;
}

}  // extern "C"
