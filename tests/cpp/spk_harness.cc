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
// feature stub for the FeatureGenerator adapter: frames = samples / 128, feature(f, d) = clamp(f) + 0.25 d
int akugpu_frontend_load_config(akugpu_ctx *, const char *) { return 0; }
int akugpu_frontend_dim(akugpu_ctx *) { return 3; }
int akugpu_frontend_sample_rate(akugpu_ctx *) { return 16000; }
int akugpu_frontend_base_is_pre(akugpu_ctx *) { return 0; }
int akugpu_frontend_base_dim(akugpu_ctx *) { return 0; }
int akugpu_frontend_pre_legacy(akugpu_ctx *) { return 0; }
int akugpu_features(akugpu_ctx *, const int16_t *, const int64_t *uo, int n_utts, void *out, int f64, int64_t *fo)
{
  fo[0] = 0;
  for (int u = 0; u < n_utts; u++) fo[u + 1] = fo[u] + (uo[u + 1] - uo[u]) / 128;
  if (out && f64)
    for (int64_t f = 0; f < fo[n_utts]; f++) for (int d = 0; d < 3; d++) ((double *)out)[f * 3 + d] = f + 0.25 * d;
  return 0;
}
int akugpu_features_range(akugpu_ctx *, const int16_t *, int64_t n_samples, int start, int end, const char *, void *out, int,
                          int *dim_out)
{
  if (dim_out) *dim_out = 3;
  const int64_t n = n_samples / 128;
  for (int f = start; out && f < end; f++)
    for (int d = 0; d < 3; d++) ((double *)out)[(size_t)(f - start) * 3 + d] = (f < 0 ? 0 : (f >= n ? n - 1 : f)) + 0.25 * d;
  return 0;
}
int akugpu_model_set_cmllr_units(akugpu_ctx *, const char *unitmode, int n, const char *const *units, const double *W)
{
  printf("cmllr_units %s", unitmode);
  for (int t = 0; t < n; t++) printf(" [%s] %g %g", units[t], W[(size_t)t * 6], W[(size_t)t * 6 + 5]);
  printf("\n");
  return 0;
}
int akugpu_features_pre(akugpu_ctx *, const float *, const int64_t *, int, void *, int, int64_t *) { return -1; }
int akugpu_features_pre_range(akugpu_ctx *, const float *, int64_t, int, int, const char *, void *, int, int *) { return -1; }

// scoring stub: S = 3 states; linear likelihood (F64) or log-likelihood (F32) of state s for a frame = f(frame sum, s)
static int g_score_calls = 0;
// resident scorer, stubbed: log((1 + x0 + x1) (s + 1) / 1000) floored at log(tiny), rows in a static buffer
static int g_session_open = 0, g_session_calls = 0;
static float g_session_rows[16 * 3];
int akugpu_stream_open(akugpu_ctx *, double idle_ms) { g_session_open = idle_ms > 0 ? 1 : -1; return 0; }
int akugpu_stream_close(akugpu_ctx *) { g_session_open = 0; return 0; }
int akugpu_stream_logprobs(akugpu_ctx *, const float *feats, int n_frames, double tiny, const float **rows)
{
  if (!g_session_open || n_frames < 1 || n_frames > 16) return -1;
  g_session_calls++;
  for (int f = 0; f < n_frames; f++)
    for (int s = 0; s < 3; s++) {
      double lik = (1 + feats[f * g_dim] + feats[f * g_dim + 1]) * (s + 1) / 1000.0;
      g_session_rows[f * 3 + s] = (float)log(lik < tiny ? tiny : lik);
    }
  *rows = g_session_rows;
  return 0;
}
int akugpu_model_num_states(akugpu_ctx *) { return 3; }
int akugpu_model_read(akugpu_ctx *, const char *) { return 0; }
int akugpu_gmm_score(akugpu_ctx *, const void *feats, int feats_f64, int64_t n_frames, int precision, void *out)
{
  g_score_calls++;
  if (!feats_f64) return -1;
  const double *x = (const double *)feats;
  for (int64_t f = 0; f < n_frames; f++) {
    const double sum = x[f * g_dim] + x[f * g_dim + 1];
    for (int s = 0; s < 3; s++) {
      const double lik = (s == 2 && sum > 100) ? 1e-60 : (sum + 1) * (s + 1);
      if (precision == AKUGPU_F64) ((double *)out)[f * 3 + s] = lik < 1e-50 ? 1e-50 : lik;
      else ((float *)out)[f * 3 + s] = (float)log(lik);
    }
  }
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
  if (std::string(argv[1]) == "fgopen") {      // spk_harness fgopen PATH: akugpu::FeatureGenerator::open(FILE*) / open_fd / open
    try {
      g_dim = 3;
      akugpu::Engine eng(0);
      akugpu::FeatureGenerator gen(eng);
      gen.load_configuration("stub.cfg");
      FILE *fp = fopen(argv[2], "rb");
      gen.open(fp, true);
      printf("FILE* %d frames, f(3,1)=%g\n", gen.num_frames(), gen.generate(3)[1]);
      fclose(fp);                              // dont_fclose: still ours
      gen.open_fd(open(argv[2], O_RDONLY));
      printf("fd %d frames, eof(last)=%d", gen.num_frames(), (gen.generate(gen.last_frame()), (int)gen.eof()));
      gen.generate(gen.last_frame() + 1);
      printf(" eof(last+1)=%d\n", (int)gen.eof());
      gen.open(argv[2]);
      printf("path %d frames\n", gen.num_frames());
    } catch (std::string &s) {
      printf("exception: %s\n", s.c_str());
      return 1;
    }
    return 0;
  }
  if (std::string(argv[1]) == "recipe") {      // spk_harness recipe PATH BATCHES INDEX SORT: akugpu::Recipe
    try {
      akugpu::Recipe r;
      r.read(argv[2], atoi(argv[3]), atoi(argv[4]));
      if (atoi(argv[5])) r.sort_infos();
      for (size_t i = 0; i < r.infos.size(); i++)
        printf("%s|%s|%s|%s|%g|%g\n", r.infos[i].audio_path.c_str(), r.infos[i].lna_path.c_str(), r.infos[i].speaker_id.c_str(),
               r.infos[i].utterance_id.c_str(), r.infos[i].start_time, r.infos[i].end_time);
    } catch (std::string &s) {
      printf("exception: %s\n", s.c_str());
      return 1;
    }
    return 0;
  }
  if (std::string(argv[1]) == "session") {     // spk_harness session: akugpu::StreamSession over the stubbed entry points
    g_dim = 2;
    akugpu::Engine eng(0);
    {
      akugpu::StreamSession ses(eng, 1e-30, 50.0);
      const float x[4] = {1, 2, 5, 6};
      const float *r = ses.log_probs(x, 2);
      printf("%d states, open=%d, rows %.6g %.6g %.6g | %.6g\n", ses.num_states(), g_session_open, r[0], r[1], r[2], r[3]);
      std::vector<float> fea(x + 2, x + 4), out;
      ses.log_probs(fea, out);
      printf("%d values, %.6g, calls=%d\n", (int)out.size(), out[2], g_session_calls);
    }
    printf("open=%d\n", g_session_open);
    return 0;
  }
  if (std::string(argv[1]) == "hmm") {         // spk_harness hmm f32|f64: the akugpu::HmmSet cache protocol
    g_dim = 2;
    akugpu::Engine eng(0);
    akugpu::HmmSet model(eng, std::string(argv[2]) == "f64" ? AKUGPU_F64 : AKUGPU_F32);
    model.read_all("stub");
    const double a[2] = {1, 2}, b[2] = {5, 6}, c[2] = {100, 1}, utt[6] = {1, 2, 5, 6, 100, 1};
    double v0, v1;                                 // (values first: argument evaluation order is unspecified)
    model.reset_cache();
    model.precompute_likelihoods(a);
    v0 = model.state_likelihood(0, a); v1 = model.state_likelihood(2, a);
    printf("%.6g %.6g calls=%d\n", v0, v1, g_score_calls);
    v0 = model.state_likelihood(1, b);             // cache keyed by reset_cache(), not by the vector
    printf("%.6g calls=%d\n", v0, g_score_calls);
    model.reset_cache();
    v0 = model.state_likelihood(1, b); v1 = model.state_likelihood(0, b);     // filled on demand
    printf("%.6g %.6g calls=%d\n", v0, v1, g_score_calls);
    model.reset_cache();
    v0 = model.state_likelihood(2, c);             // floored at 1e-50
    printf("%.6g calls=%d\n", v0, g_score_calls);
    model.set_utterance(utt, 3);
    model.precompute_likelihoods(1);
    v0 = model.state_likelihood(2); v1 = model.state_likelihood(0, 2);
    printf("%.6g %.6g calls=%d\n", v0, v1, g_score_calls);
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
