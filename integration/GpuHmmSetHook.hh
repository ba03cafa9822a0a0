// integration/GpuHmmSetHook.hh -- lets the reference's own aku::HmmSet (a concrete class, aku/HmmSet.hh:94) score on the
// GPU library without changing any caller: HmmSet::precompute_likelihoods(const FeatureVec&) asks this hook for the
// likelihoods of all emission pdfs of the frame and only runs its CPU loop (aku/HmmSet.cc:485-501) when no GPU model is
// attached.  Two edits in aku/HmmSet.cc:
//
//     void HmmSet::read_all(const std::string &base) {
//       read_mc(base + ".mc"); read_ph(base + ".ph"); read_gk(base + ".gk");
//       akugpu_hook::attach(this, base);                                   // + the same files into the GPU library
//     }
//     void HmmSet::precompute_likelihoods(const FeatureVec &f) {
//       reset_cache();
//       if (akugpu_hook::score(this, *f.get_vector(), m_pdf_likelihoods, m_valid_pdf_likelihoods)) return;   // +
//       ...                                                               // the CPU loop, unchanged
//
// This is the single-frame path (one library call per frame, ~40 us): phone_probs, the aligner and every other caller
// of precompute_likelihoods / state_likelihood run unmodified.  Whole-utterance callers should use akugpu_phone_probs /
// akugpu_gmm_score directly (that is where the throughput is).  Model-level transformations installed through
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

struct Attached {
  akugpu::Engine engine;
  int num_states, dim;
  std::vector<double> feature;
  Attached() : engine(getenv("AKUGPU_DEVICE") ? atoi(getenv("AKUGPU_DEVICE")) : 0), num_states(0), dim(0) {}
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
  akugpu::check(a.engine.ctx(), akugpu_gmm_score(a.engine.ctx(), a.feature.data(), 1, 1, AKUGPU_F64, pdf_likelihoods.data()));
  valid.clear();
  for (int i = 0; i < a.num_states; i++) valid.push_back(i);
  return true;
}

}  // namespace akugpu_hook

#endif
