// Builds integration/GpuFrontendModule.hh against the REFERENCE's headers and library (oracle/_ref/libaku_ref.a) and a
// fake of the C ABI (tests/cpp/stub_akugpu.cc: feature(f, d) = clamp(f, 0, n-1) + 0.25 d), and drives it the way
// aku::FeatureGenerator drives a base module -- with one of the reference's own modules (DeltaModule) downstream.
// TEST INFRASTRUCTURE (CPU): tests/test_abi.py compiles and runs it when /root/reference is present.
#include "GpuFrontendModule.hh"
#include "FeatureModules.hh"

using namespace aku;

int main(int argc, char **argv)
{
  if (argc < 3) return 2;
  try {
    GpuFrontendModule base;
    ModuleConfig cfg;
    cfg.set("config", std::string(argv[1]));
    base.set_name("gpu");
    base.set_config(cfg);
    DeltaModule delta;                         // the reference's module, fed by the GPU base module
    delta.add_source(&base);
    ModuleConfig dcfg;
    dcfg.set("width", 2);
    delta.set_name("d");
    delta.set_config(dcfg);
    printf("type %s dim %d rate %d fr %g\n", static_cast<FeatureModule &>(base).type_str().c_str(), base.dim(), base.sample_rate(), base.frame_rate());
    ModuleConfig back;
    base.get_config(back);
    std::string path;
    back.get("config", path);
    printf("config %s\n", path == argv[1] ? "roundtrip" : "lost");
    for (int file = 2; file < argc; file++) {
      base.reset(); delta.reset();             // FeatureGenerator::open resets every module, then names the file
      if (std::string(argv[file]) == "-") base.set_file(stdin); else base.set_fname(argv[file]);
      printf("file last_frame %d eof(last) %d eof(last+1) %d\n", base.last_frame(), (int)base.eof(base.last_frame()),
             (int)base.eof(base.last_frame() + 1));
      const int probe[] = {0, 1, 5, base.last_frame(), base.last_frame() + 3, -2, 3};     // forward, past the end, backward
      for (size_t i = 0; i < sizeof probe / sizeof probe[0]; i++) {
        const FeatureVec b = base.at(probe[i]);
        printf("base %d: %g %g %g\n", probe[i], b[0], b[1], b[2]);
      }
      for (int f = -1; f <= 2; f++) {
        const FeatureVec d = delta.at(f);
        printf("delta %d: %g %g %g\n", f, d[0], d[1], d[2]);
      }
      if (file == 2) {                         // speaker parameters reach the GPU chain and the open file is recomputed
        ModuleConfig p;
        p.set("shift.value", 2.0f);
        base.set_parameters(p);
        const FeatureVec b = base.at(4);
        printf("shifted 4: %g %g %g\n", b[0], b[1], b[2]);
        ModuleConfig q;
        q.set("shift.value", 0.0f);
        base.set_parameters(q);
        ModuleConfig bad;
        bad.set("value", 1.0f);
        try { base.set_parameters(bad); } catch (std::string &s) { printf("refused: %s\n", s.c_str()); }
      }
      base.discard_file();
    }
  } catch (std::string &s) {
    printf("exception: %s\n", s.c_str());
    return 1;
  }
  return 0;
}
