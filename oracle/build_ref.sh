#!/bin/bash
# Builds the reference's OWN sources (read in place from $AKU_REF, default
# /root/reference) into oracle/_ref/: the unmodified aku library + its literal
# phone_probs and feacat tools, against the header shims in oracle/shim/ for the
# two third-party dependencies that are absent here (LapackPP 2.5.4, libsndfile,
# plus two Boost headers) and the reference's vendored KissFFT.
# TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
# cpu_baseline / --impl reference legs may execute what this produces.
# No reference source is copied into the repository; outputs are git-ignored.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
R="${AKU_REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$R/aku" ]; then
  echo "build_ref.sh: $R/aku not found; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
CXXFLAGS="-O2 -std=gnu++11 -DKISS_FFT -DDLLIMPORT= -fPIC -fpermissive -w -I$HERE/shim -I$R/aku -I$R/vendor/kiss_fft"
LIBSRC="FeatureGenerator FeatureModules AudioReader ModuleConfig HmmSet PhnReader ModelModules SpeakerConfig Recipe conf io str endian Distributions LinearAlgebra HmmNetBaumWelch Lattice Viterbi PhonePool MllrTrainer ziggurat mtw LmbfgsOptimize RegClassTree SegErrorEvaluator util PhoneProbsToolbox"
pids=()
for f in $LIBSRC phone_probs feacat align; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$R/aku/$f.cc" -nt "$OUT/obj/$f.o" ] || [ "$HERE/shim/lapackpp.h" -nt "$OUT/obj/$f.o" ] || [ "$HERE/shim/sndfile.h" -nt "$OUT/obj/$f.o" ]; then
    g++ $CXXFLAGS -c "$R/aku/$f.cc" -o "$OUT/obj/$f.o" &
    pids+=($!)
  fi
done
for f in kiss_fft kiss_fftr; do
  gcc -O2 -fPIC -w -I"$R/vendor/kiss_fft" -c "$R/vendor/kiss_fft/$f.c" -o "$OUT/obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
LIBOBJ=""
for f in $LIBSRC kiss_fft kiss_fftr; do LIBOBJ="$LIBOBJ $OUT/obj/$f.o"; done
rm -f "$OUT/libaku_ref.a"
ar rcs "$OUT/libaku_ref.a" $LIBOBJ
g++ -O2 -o "$OUT/ref_phone_probs" "$OUT/obj/phone_probs.o" "$OUT/libaku_ref.a" -lm
g++ -O2 -o "$OUT/ref_feacat" "$OUT/obj/feacat.o" "$OUT/libaku_ref.a" -lm
g++ -O2 -o "$OUT/ref_align" "$OUT/obj/align.o" "$OUT/libaku_ref.a" -lm      # the Viterbi aligner: a lazy caller of state_likelihood
# The consumer of the LNA stream: the decoder's own reader (decoder/src/LnaReaderCircular.cc, self-contained).
g++ -O2 -fPIC -w -I"$R/decoder/src" -c "$R/decoder/src/LnaReaderCircular.cc" -o "$OUT/obj/LnaReaderCircular.o"
# Thin C-callable view of the reference classes for pytest (oracle/ref_capi.cc is ours).
g++ $CXXFLAGS -I"$R/decoder/src" -shared -o "$OUT/libref_capi.so" "$HERE/ref_capi.cc" "$OUT/obj/LnaReaderCircular.o" "$OUT/libaku_ref.a" -lm
# The reference's LITERAL tools on the GPU library: scratch copies of aku/FeatureGenerator.cc (+ the one-line registration
# of integration/GpuFrontendModule.hh) and aku/HmmSet.cc (+ the three hook lines of integration/GpuHmmSetHook.hh), written
# to oracle/_ref/obj only, linked with the unmodified phone_probs.o / feacat.o / align.o and the in-tree libakugpu.so.  They run
# as the CPU tools unless a configuration uses `type gpu_frontend` / the environment has AKUGPU_HOOK=1.
LIBAKU="$HERE/../aaltoasr_b200/libakugpu.so"
if [ -f "$LIBAKU" ]; then
  python3 - "$R" "$OUT/obj" <<'PY'
import sys
R, obj = sys.argv[1], sys.argv[2]
fg = open(R + "/aku/FeatureGenerator.cc").read()
marker = "    else\n      throw std::string(\"Unknown module type '\")"
assert fg.count(marker) == 1
fg = fg.replace('#include "FeatureModules.hh"\n', '#include "FeatureModules.hh"\n#include "GpuFrontendModule.hh"\n', 1)
fg = fg.replace(marker, "    else if (type == GpuFrontendModule::type_str())\n      module = new GpuFrontendModule();\n" + marker)
hs = open(R + "/aku/HmmSet.cc").read()
a = '  read_gk(base + ".gk");\n}\n'
b = "  // Precompute base distribution likelihoods\n  m_pool.precompute_likelihoods(*f.get_vector());\n"
c = "    return m_pdf_likelihoods[p];\n\n  m_pdf_likelihoods[p] = m_emission_pdfs[p]->compute_likelihood(*feature.get_vector());\n"
assert hs.count(a) == 1 and hs.count(b) == 1 and hs.count(c) == 1
hs = hs.replace('#include "HmmSet.hh"\n', '#include "HmmSet.hh"\n#include "GpuHmmSetHook.hh"\n', 1)
# lazy callers (Viterbi, HmmNetBaumWelch: reset_cache() + state_likelihood(s, f) for the active states only): the first
# miss after a reset scores every state of the frame on the GPU, the rest of the frame's requests are cache hits
hs = hs.replace(c, "    return m_pdf_likelihoods[p];\n  if (akugpu_hook::score(this, *feature.get_vector(), m_pdf_likelihoods, m_valid_pdf_likelihoods))\n    return m_pdf_likelihoods[p];\n\n  m_pdf_likelihoods[p] = m_emission_pdfs[p]->compute_likelihood(*feature.get_vector());\n")
hs = hs.replace(a, '  read_gk(base + ".gk");\n  akugpu_hook::attach(this, base);\n}\n')
hs = hs.replace(b, "  if (akugpu_hook::score(this, *f.get_vector(), m_pdf_likelihoods, m_valid_pdf_likelihoods)) return;\n" + b)
open(obj + "/FeatureGenerator_registered.cc", "w").write(fg)
open(obj + "/HmmSet_hooked.cc", "w").write(hs)
PY
  GPUFLAGS="$CXXFLAGS -I$HERE/../integration -I$HERE/../aaltoasr_b200/csrc/host"
  g++ $GPUFLAGS -c "$OUT/obj/FeatureGenerator_registered.cc" -o "$OUT/obj/FeatureGenerator_registered.o" &
  g++ $GPUFLAGS -c "$OUT/obj/HmmSet_hooked.cc" -o "$OUT/obj/HmmSet_hooked.o" &
  wait
  rm -f "$OUT/obj/FeatureGenerator_registered.cc" "$OUT/obj/HmmSet_hooked.cc"
  for t in phone_probs feacat align; do
    g++ -O2 -o "$OUT/ref_${t}_gpu" "$OUT/obj/$t.o" "$OUT/obj/FeatureGenerator_registered.o" "$OUT/obj/HmmSet_hooked.o" \
        "$OUT/libaku_ref.a" -L"$HERE/../aaltoasr_b200" -lakugpu -Wl,-rpath,'$ORIGIN/../../aaltoasr_b200' -lm
  done
fi
echo "oracle/_ref built: ref_phone_probs ref_feacat ref_align libref_capi.so$([ -f "$OUT/ref_phone_probs_gpu" ] && echo ' ref_phone_probs_gpu ref_feacat_gpu ref_align_gpu')"
