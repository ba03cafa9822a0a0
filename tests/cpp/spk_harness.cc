// CPU harness for the header-only C++ adapters (aaltoasr_b200/csrc/host/akugpu.hh): akugpu::SpeakerConfig driven
// against STUBS of the few C-ABI calls it makes, which print what they receive; and akugpu::read_audio (mode `audio`).  tests/test_abi.py compiles this with
// g++ and compares the calls with the Python mirror's parse of the same speaker file.  No GPU, no libakugpu.so.
#include "../../aaltoasr_b200/csrc/host/akugpu.hh"

static int g_dim = 0;
extern "C" {
struct akugpu_ctx { int dummy; };
akugpu_ctx *akugpu_create(int) { static akugpu_ctx c; return &c; }
void akugpu_destroy(akugpu_ctx *) {}
const char *akugpu_last_error(akugpu_ctx *) { return "stub"; }
int akugpu_model_dim(akugpu_ctx *) { return g_dim; }
int akugpu_frontend_set_parameters(akugpu_ctx *, const char *module, const char *text)
{
  printf("feature %s\n%s.\n", module, text);
  return 0;
}
int akugpu_model_set_cmllr(akugpu_ctx *, const double *W)
{
  if (!W) { printf("cmllr none\n"); return 0; }
  printf("cmllr");
  for (int i = 0; i < g_dim * (g_dim + 1); i++) printf(" %.17g", W[i]);
  printf("\n");
  return 0;
}
}

int main(int argc, char **argv)
{
  if (argc < 3) return 2;
  if (std::string(argv[1]) == "audio") {       // spk_harness audio PATH CONFIG_RATE RAW(0|1): akugpu::read_audio
    try {
      std::vector<int16_t> pcm;
      int rate = 0;
      akugpu::read_audio(argv[2], atoi(argv[3]), atoi(argv[4]) != 0, pcm, rate);
      long long sum = 0;
      for (size_t i = 0; i < pcm.size(); i++) sum += (long long)pcm[i] * (long long)(i % 97 + 1);
      printf("%zu %d %lld\n", pcm.size(), rate, sum);
    } catch (std::string &s) {
      printf("exception: %s\n", s.c_str());
      return 1;
    }
    return 0;
  }
  g_dim = atoi(argv[2]);
  try {
    akugpu::Engine eng(0);
    akugpu::SpeakerConfig sc(eng);
    sc.read_speaker_file(argv[1]);
    for (int i = 3; i < argc; i++) {
      printf("speaker %s\n", argv[i]);
      sc.set_speaker(argv[i]);
    }
  } catch (std::string &s) {
    printf("exception: %s\n", s.c_str());
    return 1;
  }
  return 0;
}
