// A FAKE of the C ABI (include/akugpu.h) for CPU tests of the host tools' own logic: the recipe loop, output naming,
// cropping, batching by speaker and flag handling of aaltoasr_b200/csrc/host/{phone_probs_main,feacat_main}.cc.
// TEST INFRASTRUCTURE ONLY -- it computes nothing real: "features" and "LNA bytes" are closed-form functions of the
// frame / state indices that tests/test_abi.py re-evaluates, and every call is appended to $AKUGPU_STUB_LOG.
//   frames(n_samples) = n_samples / 128         dim = 3        states = 4        16 kHz, 125 frames/s
//   feature(f, d)     = clamp(f, 0, n-1) + 0.25 d + shift      (shift: set_parameters("shift", "value X"))
//   lna byte(f, s, b) = (7 f + 3 s + b + 100 normalize + 50 cmllr + (int)shift) & 255     (f = frame within the utterance)
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "../../include/akugpu.h"

struct akugpu_ctx { int dummy; };
static akugpu_ctx g_ctx;
static double g_shift = 0;
static int g_cmllr = 0, g_pre = 0, g_legacy = 0;
static std::string g_err = "stub";

static void logf(const char *fmt, ...)
{
  const char *path = getenv("AKUGPU_STUB_LOG");
  if (!path) return;
  FILE *fp = fopen(path, "a");
  if (!fp) return;
  va_list ap;
  va_start(ap, fmt);
  vfprintf(fp, fmt, ap);
  va_end(ap);
  fclose(fp);
}
static int64_t frames_of(int64_t n_samples) { return n_samples / 128; }
static void log_pcm(const int16_t *pcm, int64_t n)
{
  long long w = 0;
  for (int64_t i = 0; i < n; i++) w += (long long)pcm[i] * (long long)(i % 97 + 1);
  logf("pcm n=%lld wsum=%lld\n", (long long)n, w);
}

extern "C" {

akugpu_ctx *akugpu_create(int device) { logf("create %d\n", device); return &g_ctx; }
void akugpu_destroy(akugpu_ctx *) {}
const char *akugpu_last_error(akugpu_ctx *) { return g_err.c_str(); }

int akugpu_frontend_load_config(akugpu_ctx *, const char *cfg_path)
{
  FILE *fp = fopen(cfg_path, "r");
  if (!fp) { g_err = std::string("could not open file ") + cfg_path; return AKUGPU_E_IO; }
  fclose(fp);
  const size_t n = strlen(cfg_path);
  g_pre = n >= 7 && !strcmp(cfg_path + n - 7, "pre.cfg");
  g_legacy = n >= 14 && !strcmp(cfg_path + n - 14, "legacy_pre.cfg");
  logf("load_config %s\n", cfg_path);
  return 0;
}
int akugpu_frontend_dim(akugpu_ctx *) { return 3; }
int akugpu_frontend_sample_rate(akugpu_ctx *) { return 16000; }
float akugpu_frontend_frame_rate(akugpu_ctx *) { return 125.0f; }
int akugpu_frontend_base_is_pre(akugpu_ctx *) { return g_pre; }
int akugpu_frontend_base_dim(akugpu_ctx *) { return 3; }
int akugpu_frontend_pre_legacy(akugpu_ctx *) { return g_legacy; }
int akugpu_frontend_set_parameters(akugpu_ctx *, const char *module, const char *text)
{
  std::string t(text);
  for (size_t i = 0; i < t.size(); i++) if (t[i] == '\n') t[i] = ';';
  logf("set_parameters %s %s\n", module, t.c_str());
  if (!strcmp(module, "shift")) {
    const char *p = strstr(text, "value ");
    g_shift = p ? atof(p + 6) : 0;
    return 0;
  }
  g_err = std::string("unknown module requested: ") + module;
  return AKUGPU_E_CONFIG;
}

static void fill_features(int64_t n, int start, int end, void *out, int f64)
{
  for (int f = start; f < end; f++) {
    int64_t c = f < 0 ? 0 : (f >= n ? n - 1 : f);
    for (int d = 0; d < 3; d++) {
      const double v = (double)c + 0.25 * d + g_shift;
      if (f64) ((double *)out)[(size_t)(f - start) * 3 + d] = v;
      else ((float *)out)[(size_t)(f - start) * 3 + d] = (float)v;
    }
  }
}
int akugpu_features(akugpu_ctx *, const int16_t *pcm, const int64_t *uo, int n_utts, void *out, int f64, int64_t *fo)
{
  fo[0] = 0;
  for (int u = 0; u < n_utts; u++) fo[u + 1] = fo[u] + frames_of(uo[u + 1] - uo[u]);
  if (!out) return 0;
  if (!pcm) { g_err = "pcm is NULL"; return AKUGPU_E_ARG; }
  logf("features n_utts=%d samples=%lld\n", n_utts, (long long)uo[n_utts]);
  log_pcm(pcm, uo[n_utts]);
  for (int u = 0; u < n_utts; u++)
    fill_features(fo[u + 1] - fo[u], 0, (int)(fo[u + 1] - fo[u]), (char *)out + (size_t)fo[u] * 3 * (f64 ? 8 : 4), f64);
  return 0;
}
int akugpu_features_range(akugpu_ctx *, const int16_t *, int64_t n_samples, int start, int end, const char *, void *out,
                          int f64, int *dim_out)
{
  if (dim_out) *dim_out = 3;
  if (out) { logf("features_range %d %d\n", start, end); fill_features(frames_of(n_samples), start, end, out, f64); }
  return 0;
}
int akugpu_features_pre(akugpu_ctx *, const float *rows, const int64_t *ro, int n_utts, void *out, int f64, int64_t *fo)
{
  for (int u = 0; u <= n_utts; u++) fo[u] = ro[u];
  if (!out) return 0;
  logf("features_pre rows=%lld\n", (long long)ro[n_utts]);
  for (int64_t i = 0; i < ro[n_utts] * 3; i++) {
    if (f64) ((double *)out)[i] = rows[i] + g_shift; else ((float *)out)[i] = rows[i] + (float)g_shift;
  }
  return 0;
}
int akugpu_features_pre_range(akugpu_ctx *, const float *rows, int64_t n_rows, int start, int end, const char *, void *out,
                              int f64, int *dim_out)
{
  if (dim_out) *dim_out = 3;
  if (!out) return 0;
  logf("features_pre_range %d %d\n", start, end);
  for (int f = start; f < end; f++) {
    int64_t c = f < 0 ? 0 : (f >= n_rows ? n_rows - 1 : f);
    for (int d = 0; d < 3; d++) {
      const double v = rows[c * 3 + d] + g_shift;
      if (f64) ((double *)out)[(size_t)(f - start) * 3 + d] = v; else ((float *)out)[(size_t)(f - start) * 3 + d] = (float)v;
    }
  }
  return 0;
}

int akugpu_model_read(akugpu_ctx *, const char *base) { logf("model_read %s\n", base); return 0; }
int akugpu_model_read_files(akugpu_ctx *, const char *gk, const char *mc, const char *ph)
{
  logf("model_read_files %s %s %s\n", gk, mc, ph);
  return 0;
}
int akugpu_model_num_states(akugpu_ctx *) { return 4; }
int akugpu_model_dim(akugpu_ctx *) { return getenv("AKUGPU_STUB_MODEL_DIM") ? atoi(getenv("AKUGPU_STUB_MODEL_DIM")) : 3; }
int akugpu_model_read_clustering(akugpu_ctx *, const char *path) { logf("read_clustering %s\n", path); return 0; }
int akugpu_model_set_clustering_min_evals(akugpu_ctx *, double a, double b) { logf("min_evals %g %g\n", a, b); return 0; }
int akugpu_model_set_cmllr(akugpu_ctx *, const double *W)
{
  g_cmllr = W ? 1 : 0;
  if (W) logf("set_cmllr %g %g\n", W[0], W[1]); else logf("set_cmllr none\n");
  return 0;
}

int akugpu_model_set_cmllr_units(akugpu_ctx *, const char *unitmode, int n, const char *const *units, const double *W)
{
  g_cmllr = n > 0 ? 1 : 0;
  logf("set_cmllr_units %s n=%d", unitmode, n);
  for (int t = 0; t < n; t++) logf(" [%s] %g", units[t], W[(size_t)t * 12]);
  logf("\n");
  return 0;
}

// F64: linear likelihood of state s for a frame = (1 + x0 + x1) (s + 1) / 1000 (the test features are >= 0), floored at 1e-50;
// F32: its logarithm.  Logged once per call.
int akugpu_gmm_score(akugpu_ctx *, const void *feats, int feats_f64, int64_t n_frames, int precision, void *out)
{
  logf("gmm_score frames=%lld precision=%d\n", (long long)n_frames, precision);
  if (!feats_f64) { g_err = "stub: double features only"; return AKUGPU_E_ARG; }
  const int D = akugpu_model_dim(NULL);
  const double *x = (const double *)feats;
  for (int64_t f = 0; f < n_frames; f++)
    for (int s = 0; s < 4; s++) {
      double lik = (1 + x[f * D] + x[f * D + 1]) * (s + 1) / 1000.0;
      if (lik < 1e-50) lik = 1e-50;
      if (precision == AKUGPU_F64) ((double *)out)[f * 4 + s] = lik; else ((float *)out)[f * 4 + s] = (float)log(lik);
    }
  return 0;
}

// records of explicit feature rows: byte = (200 + 7 row + 3 s + b + 100 normalize + (int)x[row][0]) & 255
int akugpu_gmm_lna(akugpu_ctx *, const void *feats, int feats_f64, int64_t n_frames, int precision, int lnabytes, int normalize,
                   uint8_t *out)
{
  logf("gmm_lna frames=%lld precision=%d lnabytes=%d normalize=%d\n", (long long)n_frames, precision, lnabytes, normalize);
  if (!feats_f64) { g_err = "stub: double features only"; return AKUGPU_E_ARG; }
  const double *x = (const double *)feats;
  for (int64_t f = 0; f < n_frames; f++)
    for (int s = 0; s < 4; s++)
      for (int b = 0; b < lnabytes; b++)
        out[((size_t)f * 4 + s) * lnabytes + b] = (uint8_t)((200 + 7 * f + 3 * s + b + 100 * normalize + (int)x[f * 3]) & 255);
  return 0;
}

int akugpu_lna_header(int n_states, int lnabytes, uint8_t out5[5])
{
  out5[0] = (uint8_t)(n_states >> 24); out5[1] = (uint8_t)(n_states >> 16); out5[2] = (uint8_t)(n_states >> 8);
  out5[3] = (uint8_t)n_states; out5[4] = (uint8_t)lnabytes;
  return 0;
}
int akugpu_phone_probs(akugpu_ctx *, const int16_t *pcm, const int64_t *uo, int n_utts, int precision, int lnabytes,
                       int normalize, uint8_t *out, int64_t *fo, uint64_t *)
{
  fo[0] = 0;
  for (int u = 0; u < n_utts; u++) fo[u + 1] = fo[u] + frames_of(uo[u + 1] - uo[u]);
  if (n_utts && !pcm) { g_err = "pcm is NULL"; return AKUGPU_E_ARG; }
  logf("phone_probs n_utts=%d samples=%lld precision=%d lnabytes=%d normalize=%d\n", n_utts, (long long)uo[n_utts], precision,
       lnabytes, normalize);
  if (n_utts) log_pcm(pcm, uo[n_utts]);
  if (!out) return 0;
  for (int u = 0; u < n_utts; u++)
    for (int64_t f = 0; f < fo[u + 1] - fo[u]; f++)
      for (int s = 0; s < 4; s++)
        for (int b = 0; b < lnabytes; b++)
          out[((size_t)(fo[u] + f) * 4 + s) * lnabytes + b] =
              (uint8_t)((7 * f + 3 * s + b + 100 * normalize + 50 * g_cmllr + (int)g_shift) & 255);
  return 0;
}

}  // extern "C"
