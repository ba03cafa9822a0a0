// integration/GpuHmmSetHook.hh -- lets the reference's own aku::HmmSet (a concrete class, aku/HmmSet.hh:94) score on the
// GPU library without changing any caller: HmmSet::precompute_likelihoods(const FeatureVec&) asks this hook for the
// likelihoods of all emission pdfs of the frame and only runs its CPU loop (aku/HmmSet.cc:485-501) when no GPU model is
// attached.  Three edits in aku/HmmSet.cc:
//
//     void HmmSet::read_all(const std::string &base) {
//       read_mc(base + ".mc"); read_ph(base + ".ph"); read_gk(base + ".gk");
//       akugpu_hook::attach(this, base);                                   // + the same files into the GPU library
//     }
//     void HmmSet::precompute_likelihoods(const FeatureVec &f) {
//       reset_cache();
//       if (akugpu_hook::score(this, *f.get_vector(), m_pdf_likelihoods, m_valid_pdf_likelihoods)) return;   // +
//       ...                                                               // the CPU loop, unchanged
//     double HmmSet::pdf_likelihood(const int p, const FeatureVec &feature) {
//       if (m_pdf_likelihoods[p] > 0) return m_pdf_likelihoods[p];
//       if (akugpu_hook::score(this, *feature.get_vector(), m_pdf_likelihoods, m_valid_pdf_likelihoods))   // + lazy callers
//         return m_pdf_likelihoods[p];                                    //   (Viterbi.cc:249,369, HmmNetBaumWelch.cc:1930):
//       ...                                                               //   the first miss after reset_cache() scores the
//                                                                         //   frame's every state, the rest are cache hits
//
// phone_probs, the aligner and every other caller of precompute_likelihoods / state_likelihood run unmodified.  When the
// feature vector handed in is a frame of the utterance that integration/GpuFrontendModule.hh has just computed (the
// module publishes its matrix below), the WHOLE utterance is scored in one library call at the first request and every
// later frame is served from that result -- the throughput path, reached from the reference's own per-frame loop.
// Any other vector (e.g. after a CPU-side module downstream of the GPU module) takes the single-frame path: one
// library call per frame, ~40 us.  Model-level transformations installed through
// SpeakerConfig wrap the CPU Gaussians only; with the hook attached they must be given to the library
// (akugpu_model_set_cmllr).  The hook is off unless the environment has AKUGPU_HOOK=1 (the reference's tools link it in
// but stay CPU-only by default).
#ifndef GPUHMMSETHOOK_HH
#define GPUHMMSETHOOK_HH

#include <stdlib.h>
#include <map>
#include <string>
#include <vector>
#include "akugpu.hh"        // aaltoasr_b200/csrc/host

namespace aku { class HmmSet; }

namespace akugpu_hook {

// The utterance most recently computed by a GpuFrontendModule: [frames x dim] doubles, owned by the module.
struct Utterance {
  const double *feats;
  int frames, dim;
  long serial;             // bumped whenever the matrix changes
};
inline Utterance &current_utterance()
{
  static Utterance u = {NULL, 0, 0, 0};
  return u;
}
inline void publish_utterance(const double *feats, int frames, int dim)
{
  Utterance &u = current_utterance();
  u.feats = feats; u.frames = frames; u.dim = dim; u.serial++;
}

struct Attached {
  akugpu::Engine engine;
  int num_states, dim;
  std::vector<double> feature;
  std::vector<double> lik;     // [frames x states] of the utterance with serial `scored`
  long scored;
  int next;                    // the frame expected next (callers walk forward)
  Attached() : engine(getenv("AKUGPU_DEVICE") ? atoi(getenv("AKUGPU_DEVICE")) : 0), num_states(0), dim(0), scored(-1), next(0) {}
};

inline std::map<const aku::HmmSet *, Attached *> &registry()
{
  static std::map<const aku::HmmSet *, Attached *> r;
  return r;
}

inline void detach(const aku::HmmSet *model)
{
  std::map<const aku::HmmSet *, Attached *>::iterator it = registry().find(model);
  if (it == registry().end()) return;
  delete it->second;
  registry().erase(it);
}

inline void attach(const aku::HmmSet *model, const std::string &base)
{
  const char *on = getenv("AKUGPU_HOOK");
  if (!on || atoi(on) == 0) return;
  detach(model);
  Attached *a = new Attached;
  try {
    akugpu::check(a->engine.ctx(), akugpu_model_read(a->engine.ctx(), base.c_str()));
  } catch (...) { delete a; throw; }
  a->num_states = akugpu_model_num_states(a->engine.ctx());
  a->dim = akugpu_model_dim(a->engine.ctx());
  registry()[model] = a;
}

// Likelihoods (linear, floored at 1e-50 like aku/HmmSet.cc:494-497) of every emission pdf for one feature vector.
// Vec is the reference's Vector type (operator()(int), size()); valid receives 0..n-1 as the CPU loop leaves it.
template <class Vec>
inline bool score(const aku::HmmSet *model, const Vec &f, std::vector<double> &pdf_likelihoods, std::vector<int> &valid)
{
  std::map<const aku::HmmSet *, Attached *>::iterator it = registry().find(model);
  if (it == registry().end()) return false;
  Attached &a = *it->second;
  if ((int)f.size() != a.dim) throw std::string("akugpu_hook: feature dimension and model dimension don't agree");
  a.feature.resize(a.dim);
  for (int i = 0; i < a.dim; i++) a.feature[i] = f(i);
  if ((int)pdf_likelihoods.size() < a.num_states) throw std::string("akugpu_hook: the GPU model has more states than the HmmSet");
  // a frame of the published utterance?  (exact comparison: the module's rows reach the caller as copies)
  const Utterance &u = current_utterance();
  int idx = -1;
  if (u.feats && u.dim == a.dim && u.frames > 0) {
    for (int k = 0; k < u.frames && idx < 0; k++) {
      const int cand = (a.next + k) % u.frames;          // the expected frame first, then the others
      const double *row = u.feats + (size_t)cand * a.dim;
      int i = 0;
      while (i < a.dim && row[i] == a.feature[i]) i++;
      if (i == a.dim) idx = cand;
    }
  }
  if (idx >= 0) {
    if (a.scored != u.serial) {
      a.lik.resize((size_t)u.frames * a.num_states);
      akugpu::check(a.engine.ctx(), akugpu_gmm_score(a.engine.ctx(), u.feats, 1, u.frames, AKUGPU_F64, a.lik.data()));
      a.scored = u.serial;
    }
    for (int i = 0; i < a.num_states; i++) pdf_likelihoods[i] = a.lik[(size_t)idx * a.num_states + i];
    a.next = idx + 1;
  } else {
    akugpu::check(a.engine.ctx(), akugpu_gmm_score(a.engine.ctx(), a.feature.data(), 1, 1, AKUGPU_F64, pdf_likelihoods.data()));
  }
  valid.clear();
  for (int i = 0; i < a.num_states; i++) valid.push_back(i);
  return true;
}

}  // namespace akugpu_hook

#endif
