// api.cu -- the extern "C" boundary declared in include/akugpu.h.
// No C++ type and no exception crosses it; every entry point converts
// akugpu::Error into a return code + akugpu_last_error().
#include "ctx.hpp"
#include "kernels.hpp"
#include <string.h>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <mutex>
#include <chrono>

using namespace akugpu;

static std::string g_create_err;

// API_BEGIN_KEEP: entry points that may be served by the resident scorer (gmm_resident.cu); every other entry point
// ends a running resident kernel first, so that the device is whole before anything else is launched or allocated.
#define API_BEGIN_KEEP                                        \
  if (!ctx) return AKUGPU_E_ARG;                              \
  try {                                                       \
    AKU_CUDA(cudaSetDevice(ctx->device));
#define API_BEGIN                                             \
  API_BEGIN_KEEP                                              \
    session_quiesce(ctx);
#define API_END                                               \
    return AKUGPU_OK;                                         \
  } catch (const Error &e) {                                  \
    ctx->err = e.msg;                                         \
    return e.code;                                            \
  } catch (const std::exception &e) {                         \
    ctx->err = e.what();                                      \
    return AKUGPU_E_ARG;                                      \
  }

namespace akugpu {

void ensure_dynamic_smem(akugpu_ctx *ctx, const void *kernel, size_t bytes)
{
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> have;
  std::lock_guard<std::mutex> lock(mu);
  size_t &h = have[std::make_pair(ctx->device, kernel)];
  if (bytes > h) {
    AKU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    h = bytes;
  }
}

StageScope::StageScope(akugpu_ctx *c, int s) : ctx(c), stage(s), l0(c->launches)
{
  if (!ctx->timer.enabled) return;
  auto get = [&]() {
    cudaEvent_t e;
    if (!ctx->timer.pool.empty()) { e = ctx->timer.pool.back(); ctx->timer.pool.pop_back(); }
    else AKU_CUDA(cudaEventCreate(&e));
    return e;
  };
  e0 = get(); e1 = get();
  st = ctx->stream;
  AKU_CUDA(cudaEventRecord(e0, st));
}
StageScope::~StageScope()
{
  if (!ctx->timer.enabled || !e0) return;
  cudaEventRecord(e1, st);
  ctx->timer.pending.push_back(std::make_pair(stage, std::make_pair(e0, e1)));
  ctx->timer.launches[stage] += ctx->launches - l0;
}

}  // namespace akugpu

// Frames per chunk of the scoring pipeline.  Auto (chunk_frames == 0): exactly one wave of the fp32
// scorer (sm_count x resident CTAs x 64 frames) so that no launch ends in a partial wave; an explicit
// value is rounded to whole waves when it is at least one wave.  Always a multiple of 128.
static int64_t pick_chunk(akugpu_ctx *ctx, int64_t F, int use_tc)
{
  if (use_tc) {   // tensor-core scorers: 1 = bf16x3 (two CTAs per SM), 2 = fp16x2 (one CTA per SM; two waves per chunk)
    const int64_t wave = use_tc == 2 ? 2 * gmm_tc16_wave_frames(ctx) : gmm_tc_wave_frames(ctx);
    int64_t chunk = ctx->chunk_frames <= 0 ? wave : std::max<int64_t>(128, (ctx->chunk_frames + 127) / 128 * 128);
    if (chunk > F) chunk = (F + 127) / 128 * 128;
    return chunk;
  }
  if (ctx->hm.n_full > 0) {   // full-covariance path keeps a [G][chunk] double matrix: bound it to ~1 GB
    int64_t chunk = std::max<int64_t>(128, ((int64_t)1 << 27) / std::max(1, ctx->hm.G) / 128 * 128);
    if (ctx->chunk_frames > 0) chunk = std::min<int64_t>(chunk, (ctx->chunk_frames + 127) / 128 * 128);
    if (chunk > F) chunk = (F + 127) / 128 * 128;
    return chunk;
  }
  const int64_t wave = gmm_wave_frames(ctx);
  int64_t chunk = ctx->chunk_frames <= 0 ? wave : ctx->chunk_frames;
  if (chunk >= wave) chunk = chunk / wave * wave;
  chunk = (chunk + 127) / 128 * 128;
  if (chunk > F) chunk = (F + 127) / 128 * 128;
  return chunk;
}

static void require_model(akugpu_ctx *ctx)
{
  if (!ctx->have_model) throw Error(AKUGPU_E_STATE, "no acoustic model loaded (akugpu_model_read / akugpu_model_load_diag)");
}
static void require_frontend(akugpu_ctx *ctx)
{
  if (!ctx->fe.configured) throw Error(AKUGPU_E_STATE, "no feature configuration loaded (akugpu_frontend_load_config)");
}

// Scores frames [0,F) of device-resident features chunk by chunk and emits LNA records.
// out may be host (pipelined D2H on a second stream), device (written in place) or NULL.
// Which tensor-core scorer serves throughput-mode calls: 2 = fp16x2 (diagonal pools), 1 = bf16x3, 0 = none.
static bool clustering_on(akugpu_ctx *ctx) { return ctx->hm.use_clustering && ctx->hm.n_clusters > 0; }

static int tc_mode(akugpu_ctx *ctx, int precision)
{
  if (precision != AKUGPU_F32) return 0;
  if (clustering_on(ctx)) return 0;      // the clustering approximation is evaluated in double (reference semantics)
  if (ctx->hm.n_tr > 0) return 0;        // regression-class CMLLR: one feature row per class, served by the double path
  if (ctx->ptc16.ready && !ctx->tc16_suspended) return 2;
  return ctx->ptc.ready ? 1 : 0;
}

// The fp16x2 scorer flags features outside the fp16 range of its scaled terms; such a call is redone with the
// bf16x3 kernel (no range limit), packed on first need.
static bool tc16_needs_redo(akugpu_ctx *ctx, int mode)
{
  if (mode != 2 || !gmm_tc16_overflowed(ctx)) return false;
  if (!ctx->ptc.ready) model_pack_tc(ctx);
  return true;
}

static void score_to_lna_impl(akugpu_ctx *ctx, const void *d_feats, int feats_f64, int64_t F, int precision, int lnabytes,
                              int normalize, uint8_t *out, uint64_t *checksum_out, int *mode_out, bool utt_chk = false)
{
  const int S = ctx->hm.S;
  if (lnabytes != 2 && lnabytes != 4) throw Error(AKUGPU_E_ARG, "lnabytes must be 2 or 4");
  if (precision != AKUGPU_F32 && precision != AKUGPU_F64) throw Error(AKUGPU_E_ARG, "precision must be AKUGPU_F32 or AKUGPU_F64");
  if (F <= 0 || S <= 0) { if (checksum_out) *checksum_out = 0; return; }
  const int use_tc = tc_mode(ctx, precision);
  *mode_out = use_tc;
  if (ctx->hm.n_full > 0 && !use_tc) precision = AKUGPU_F64;   // full-covariance pools are scored in double
  if (clustering_on(ctx) || ctx->hm.n_tr > 0) precision = AKUGPU_F64;
  const int64_t chunk = pick_chunk(ctx, F, use_tc);
  const size_t rec = (size_t)S * lnabytes;
  const bool out_dev = out && is_device_ptr(out);
  const bool out_host = out && !out_dev;
  if (out_dev && ((uintptr_t)out & 3))      // the LNA kernels store 32-bit words (16-bit when S * lnabytes is odd-sized)
    throw Error(AKUGPU_E_ARG, "a device `out` buffer must be 4-byte aligned");
  const size_t esz = precision == AKUGPU_F64 ? 8 : 4;
  // Throughput mode with more than one chunk: the LNA epilogue of chunk k (HBM bound, 33 KB of shared memory per CTA)
  // runs on a second stream next to the scorer of chunk k+1 (compute bound; its ring is one slot shorter so that
  // both fit an SM): scores and normalisers are double buffered.
  const bool overlap = precision == AKUGPU_F32 && ctx->overlap_lna && F > chunk;
  ctx->d_sll.reserve((size_t)S * chunk * esz);
  if (overlap) ctx->d_sll2.reserve((size_t)S * chunk * esz);
  if (!out_dev) { ctx->d_lna[0].reserve(chunk * rec); if (out_host) ctx->d_lna[1].reserve(chunk * rec); }
  if (checksum_out) { ctx->d_chk.reserve(8); AKU_CUDA(cudaMemsetAsync(ctx->d_chk.p, 0, 8, ctx->stream)); }
  cudaStream_t const main_stream = ctx->stream;
  struct Restore { akugpu_ctx *c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{ctx, main_stream};
  if (overlap) {   // what the LNA stream touches first must be ordered after what the main stream did before this call
    AKU_CUDA(cudaEventRecord(ctx->ev_sc[0], main_stream));
    AKU_CUDA(cudaStreamWaitEvent(ctx->lna_stream, ctx->ev_sc[0], 0));
  }
  int c = 0;
  for (int64_t c0 = 0; c0 < F; c0 += chunk, ++c) {
    const int64_t c1 = std::min(F, c0 + chunk), nf = c1 - c0;
    const int b = out_host ? (c & 1) : 0;
    const int b2 = overlap ? (c & 1) : 0;
    DevBuf &sllb = b2 ? ctx->d_sll2 : ctx->d_sll;
    DevBuf &normb = b2 ? ctx->d_norm2 : ctx->d_norm;
    uint8_t *dst = out_dev ? out + c0 * rec : ctx->d_lna[b].as<uint8_t>();
    if (overlap && c >= 2) AKU_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_ln[b2], 0));      // the LNA kernel of chunk c-2 has read this buffer
    const float2 *norm = nullptr;
    if (precision == AKUGPU_F32) {
      if (use_tc) {   // times its own stages; also yields the per-frame normaliser when it sweeps all states
        normb.reserve((size_t)chunk * sizeof(float2));
        // ill-conditioned states: FP32-pipe kernel, same chunk -- also when the call is being redone with the bf16x3
        // kernel after an fp16 range overflow (that image holds every state in the expanded form; the FP32 image
        // holds exactly the states the expanded form is not trusted with and overwrites their rows)
        const bool hybrid = ctx->ptc16.ready && ctx->ptc16.hybrid;
        float2 *const nrm = hybrid ? nullptr : normb.as<float2>();
        const bool got = use_tc == 2 ? launch_gmm_tc16(ctx, d_feats, feats_f64, c0, c1, sllb.as<float>(), chunk, nrm)
                                     : launch_gmm_tc(ctx, d_feats, feats_f64, c0, c1, sllb.as<float>(), chunk, nrm);
        if (got) norm = normb.as<float2>();
        if (hybrid) { StageScope sc(ctx, 1); launch_gmm_f32(ctx, d_feats, feats_f64, c0, c1, sllb.as<float>(), chunk); }
      } else {
        StageScope sc(ctx, 1);
        launch_gmm_f32(ctx, d_feats, feats_f64, c0, c1, sllb.as<float>(), chunk);
      }
    } else {
      StageScope sc(ctx, 1);
      if (ctx->hm.n_full > 0) launch_gmm_full_f64(ctx, d_feats, feats_f64, c0, c1, sllb.as<double>(), chunk);
      else launch_gmm_f64(ctx, d_feats, feats_f64, c0, c1, sllb.as<double>(), chunk);
    }
    if (overlap) {   // everything from here to the end of the iteration is issued on the LNA stream
      AKU_CUDA(cudaEventRecord(ctx->ev_sc[b2], main_stream));
      ctx->stream = ctx->lna_stream;
      AKU_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_sc[b2], 0));
    }
    if (out_host && c >= 2) AKU_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));       // the ring buffer has left the device
    if (precision == AKUGPU_F32) {
      if (normalize && !norm && overlap) normb.reserve((size_t)(chunk + 31) / 32 * 32 * sizeof(float2));
      StageScope sc(ctx, 2);
      launch_lna_f32(ctx, sllb.as<float>(), chunk, S, nf, lnabytes, normalize, norm, dst, overlap ? normb.as<float2>() : nullptr);
    } else {
      StageScope sc(ctx, 2);
      launch_lna_f64(ctx, sllb.as<double>(), chunk, S, nf, lnabytes, normalize, dst);
    }
    if (checksum_out) launch_checksum(ctx, dst, nf * rec, ctx->d_chk.as<unsigned long long>());
    if (utt_chk) checksum_update(ctx, dst, c0, nf);      // per-utterance sums; the records are still in L2
    if (out_host) {
      AKU_CUDA(cudaEventRecord(ctx->ev_k[b], ctx->stream));
      AKU_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_k[b], 0));
      AKU_CUDA(cudaMemcpyAsync(out + c0 * rec, dst, nf * rec, cudaMemcpyDeviceToHost, ctx->copy_out));
      AKU_CUDA(cudaEventRecord(ctx->ev_out[b], ctx->copy_out));
    }
    if (overlap) {
      AKU_CUDA(cudaEventRecord(ctx->ev_ln[b2], ctx->stream));
      ctx->stream = main_stream;
    }
  }
  if (overlap) {   // the caller's stream sees the whole call: it waits for the LNA stream's last event
    AKU_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_ln[(c - 1) & 1], 0));
    if (c >= 2) AKU_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_ln[c & 1], 0));
  }
  if (checksum_out) {
    unsigned long long h = 0;
    AKU_CUDA(cudaMemcpyAsync(&h, ctx->d_chk.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
    *checksum_out = h;
  }
  if (out_host) AKU_CUDA(cudaStreamSynchronize(ctx->copy_out));
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
}

static void score_to_lna(akugpu_ctx *ctx, const void *d_feats, int feats_f64, int64_t F, int precision, int lnabytes,
                         int normalize, uint8_t *out, uint64_t *checksum_out, const std::vector<int64_t> *fo = nullptr,
                         uint64_t *utt_checksums = nullptr)
{
  int mode = 0;
  const bool utt_chk = fo && utt_checksums;
  const int n_utts = utt_chk ? (int)fo->size() - 1 : 0;
  if (utt_chk) checksum_begin(ctx, fo->data(), n_utts, (int64_t)ctx->hm.S * lnabytes);
  score_to_lna_impl(ctx, d_feats, feats_f64, F, precision, lnabytes, normalize, out, checksum_out, &mode, utt_chk);
  if (tc16_needs_redo(ctx, mode)) {
    ctx->tc16_suspended = true;
    try {
      if (utt_chk) checksum_begin(ctx, fo->data(), n_utts, (int64_t)ctx->hm.S * lnabytes);
      score_to_lna_impl(ctx, d_feats, feats_f64, F, precision, lnabytes, normalize, out, checksum_out, &mode, utt_chk);
    }
    catch (...) { ctx->tc16_suspended = false; throw; }
    ctx->tc16_suspended = false;
  }
  if (utt_chk) checksum_end(ctx, utt_checksums);
}

// Makes `src` (host or device) available on the device; returns the device pointer.
static const void *to_device(akugpu_ctx *ctx, const void *src, size_t bytes, DevBuf &scratch)
{
  if (is_device_ptr(src)) return src;
  scratch.reserve(std::max<size_t>(bytes, 16));
  AKU_CUDA(cudaMemcpyAsync(scratch.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return scratch.p;
}

// With a global model-level CMLLR set (akugpu_model_set_cmllr) the scorers see A f + b instead of f.
static const void *adapt_feats(akugpu_ctx *ctx, const void *d_feats, int feats_f64, int64_t F)
{
  if (!ctx->hm.cmllr_on || F <= 0) return d_feats;
  const int D = ctx->hm.D;
  if (ctx->hm.n_tr > 0) {
    // regression classes: one adapted copy of the rows per transform next to the unadapted rows; the double scorer
    // picks a Gaussian's row by its class (gmm_diag_f64)
    const size_t esz = feats_f64 ? 8 : 4;
    ctx->d_adapt.reserve((size_t)ctx->hm.n_tr * F * D * esz);
    ctx->adapt_stride = F * D;
    StageScope sc(ctx, 1);
    for (int t = 0; t < ctx->hm.n_tr; t++)
      launch_affine_rows(ctx, d_feats, feats_f64, F, D, ctx->d_cmllr.as<double>() + (size_t)t * (D * D + D),
                         (char *)ctx->d_adapt.p + (size_t)t * F * D * esz);
    return d_feats;
  }
  ctx->d_adapt.reserve((size_t)F * D * (feats_f64 ? 8 : 4));
  StageScope sc(ctx, 1);
  launch_affine_rows(ctx, d_feats, feats_f64, F, D, ctx->d_cmllr.as<double>(), ctx->d_adapt.p);
  return ctx->d_adapt.p;
}

extern "C" {

akugpu_ctx *akugpu_create(int device)
{
  try {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      cudaGetLastError();
      throw Error(AKUGPU_E_CUDA, std::string("no CUDA device available (") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                     "); akugpu has no CPU fallback");
    }
    if (device < 0 || device >= n) throw Error(AKUGPU_E_ARG, fmt("device %d out of range (0..%d)", device, n - 1));
    AKU_CUDA(cudaSetDevice(device));
    akugpu_ctx *ctx = new akugpu_ctx;
    ctx->device = device;
    cudaDeviceProp prop;
    AKU_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
      delete ctx;
      throw Error(AKUGPU_E_CUDA, fmt("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                                     prop.minor));
    }
    AKU_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    AKU_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    AKU_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    AKU_CUDA(cudaStreamCreateWithFlags(&ctx->lna_stream, cudaStreamNonBlocking));
    if (getenv("AKUGPU_OVERLAP")) ctx->overlap_lna = atoi(getenv("AKUGPU_OVERLAP")) != 0;
    if (getenv("AKUGPU_CHUNK_FRAMES")) ctx->chunk_frames = atoll(getenv("AKUGPU_CHUNK_FRAMES"));
    for (int i = 0; i < 2; i++) {
      AKU_CUDA(cudaEventCreateWithFlags(&ctx->ev_sc[i], cudaEventDisableTiming));
      AKU_CUDA(cudaEventCreateWithFlags(&ctx->ev_ln[i], cudaEventDisableTiming));
      AKU_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
      AKU_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
      AKU_CUDA(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
    }
    return ctx;
  } catch (const Error &e) {
    g_create_err = e.msg;
    return nullptr;
  }
}

void akugpu_destroy(akugpu_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  session_destroy(ctx);       // a resident kernel would hold the device synchronisation until its idle timer
  cudaDeviceSynchronize();
  for (auto &p : ctx->timer.pending) { cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second); }
  for (auto e : ctx->timer.pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; i++) {
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
    if (ctx->ev_sc[i]) cudaEventDestroy(ctx->ev_sc[i]);
    if (ctx->ev_ln[i]) cudaEventDestroy(ctx->ev_ln[i]);
  }
  if (ctx->lna_stream) cudaStreamDestroy(ctx->lna_stream);
  if (ctx->stream_state.host) cudaFreeHost(ctx->stream_state.host);
  for (void *p : ctx->shared_peer) cudaIpcCloseMemHandle(p);
  for (void *p : ctx->shared_own) cudaFree(p);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  delete ctx;
}

const char *akugpu_last_error(akugpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int akugpu_set_stream(akugpu_ctx *ctx, void *cuda_stream)
{
  API_BEGIN
  if (cuda_stream) {
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
  } else if (!ctx->own_stream) {
    AKU_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  API_END
}

int akugpu_synchronize(akugpu_ctx *ctx)
{
  API_BEGIN
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  AKU_CUDA(cudaStreamSynchronize(ctx->copy_out));
  API_END
}

int64_t akugpu_launch_count(akugpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int akugpu_stage_times_reset(akugpu_ctx *ctx, int enable)
{
  API_BEGIN
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &p : ctx->timer.pending) { ctx->timer.pool.push_back(p.second.first); ctx->timer.pool.push_back(p.second.second); }
  ctx->timer.pending.clear();
  for (int i = 0; i < 3; i++) { ctx->timer.ms[i] = 0; ctx->timer.launches[i] = 0; }
  ctx->timer.enabled = enable != 0;
  API_END
}

int akugpu_stage_times(akugpu_ctx *ctx, double ms_out[3], int64_t launches_out[3])
{
  API_BEGIN
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &p : ctx->timer.pending) {
    float ms = 0;
    AKU_CUDA(cudaEventElapsedTime(&ms, p.second.first, p.second.second));
    ctx->timer.ms[p.first] += ms;
    ctx->timer.pool.push_back(p.second.first);
    ctx->timer.pool.push_back(p.second.second);
  }
  ctx->timer.pending.clear();
  for (int i = 0; i < 3; i++) {
    if (ms_out) ms_out[i] = ctx->timer.ms[i];
    if (launches_out) launches_out[i] = ctx->timer.launches[i];
  }
  API_END
}

// ---- front-end ---------------------------------------------------------------------
int akugpu_frontend_load_config_text(akugpu_ctx *ctx, const char *cfg_text)
{
  API_BEGIN
  if (!cfg_text) throw Error(AKUGPU_E_ARG, "cfg_text is NULL");
  frontend_parse(ctx, cfg_text);
  API_END
}

int akugpu_frontend_load_config(akugpu_ctx *ctx, const char *cfg_path)
{
  API_BEGIN
  if (!cfg_path) throw Error(AKUGPU_E_ARG, "cfg_path is NULL");
  std::ifstream in(cfg_path, std::ios::binary);
  if (!in) throw Error(AKUGPU_E_IO, std::string("could not open ") + cfg_path);
  std::ostringstream ss;
  ss << in.rdbuf();
  frontend_parse(ctx, ss.str());
  API_END
}

int akugpu_frontend_dim(akugpu_ctx *ctx) { return (ctx && ctx->fe.configured) ? ctx->fe.mods[ctx->fe.last].dim : AKUGPU_E_STATE; }
int akugpu_frontend_sample_rate(akugpu_ctx *ctx) { return (ctx && ctx->fe.configured) ? ctx->fe.mods[0].sample_rate : AKUGPU_E_STATE; }
float akugpu_frontend_frame_rate(akugpu_ctx *ctx) { return (ctx && ctx->fe.configured) ? ctx->fe.mods[0].frame_rate : -1.f; }
int akugpu_frontend_base_is_pre(akugpu_ctx *ctx) { return (ctx && ctx->fe.configured) ? (ctx->fe.mods[0].type == M_PRE ? 1 : 0) : AKUGPU_E_STATE; }
int akugpu_frontend_base_dim(akugpu_ctx *ctx) { return (ctx && ctx->fe.configured) ? ctx->fe.mods[0].dim : AKUGPU_E_STATE; }
int akugpu_frontend_pre_legacy(akugpu_ctx *ctx) { return (ctx && ctx->fe.configured) ? (ctx->fe.mods[0].type == M_PRE && ctx->fe.mods[0].legacy_file ? 1 : 0) : AKUGPU_E_STATE; }
int64_t akugpu_frontend_num_frames(akugpu_ctx *ctx, int64_t n_samples)
{
  if (!ctx || !ctx->fe.configured) return AKUGPU_E_STATE;
  return frontend_num_frames(ctx->fe, n_samples);
}

int akugpu_frontend_set_parameters(akugpu_ctx *ctx, const char *module_name, const char *text)
{
  API_BEGIN
  require_frontend(ctx);
  if (!module_name || !text) throw Error(AKUGPU_E_ARG, "module_name/text is NULL");
  frontend_set_parameters(ctx, module_name, text);
  API_END
}

static void frame_offsets_of(akugpu_ctx *ctx, const int64_t *utt_offsets, int n_utts, std::vector<int64_t> &uo,
                             std::vector<int64_t> &fo)
{
  if (n_utts < 0 || !utt_offsets) throw Error(AKUGPU_E_ARG, "utt_offsets is NULL or n_utts < 0");
  uo.assign(utt_offsets, utt_offsets + n_utts + 1);
  fo.assign(n_utts + 1, 0);
  for (int u = 0; u < n_utts; u++) {
    if (uo[u + 1] < uo[u]) throw Error(AKUGPU_E_ARG, "utt_offsets must be non-decreasing");
    int64_t n = frontend_num_frames(ctx->fe, uo[u + 1] - uo[u]);
    if (n <= 0) throw Error(AKUGPU_E_ARG, fmt("utterance %d: audio shorter than frame", u));   // aku/FeatureModules.cc:409
    fo[u + 1] = fo[u] + n;
  }
}

// Bytes per input unit (sample or stored row) of the configured base module; want_pre says which entry point was used.
static size_t base_unit_bytes(akugpu_ctx *ctx, bool want_pre)
{
  const bool pre = ctx->fe.mods[0].type == M_PRE;
  if (pre != want_pre)
    throw Error(AKUGPU_E_STATE, pre ? "the configured base module is `pre`: use akugpu_features_pre / akugpu_features_pre_range"
                                    : "the configured base module is `audiofile`: akugpu_features_pre needs a `pre` base module");
  return pre ? (size_t)ctx->fe.mods[0].dim * sizeof(float) : sizeof(int16_t);
}

static void features_impl(akugpu_ctx *ctx, const void *pcm, const int64_t *utt_offsets, int n_utts, void *out, int out_f64,
                          int64_t *frame_offsets, bool want_pre)
{
  require_frontend(ctx);
  const size_t unit = base_unit_bytes(ctx, want_pre);
  std::vector<int64_t> uo, fo;
  frame_offsets_of(ctx, utt_offsets, n_utts, uo, fo);
  if (frame_offsets) memcpy(frame_offsets, fo.data(), fo.size() * sizeof(int64_t));
  if (out && n_utts > 0) {
    if (!pcm) throw Error(AKUGPU_E_ARG, "pcm is NULL");
    const int dim = ctx->fe.mods[ctx->fe.last].dim;
    const size_t esz = out_f64 ? 8 : 4;
    const void *d_pcm = to_device(ctx, pcm, (size_t)uo[n_utts] * unit, ctx->d_pcm);
    const bool odev = is_device_ptr(out);
    void *d_out = out;
    if (!odev) { ctx->d_feats.reserve((size_t)fo[n_utts] * dim * esz); d_out = ctx->d_feats.p; }
    { StageScope sc(ctx, 0); frontend_run_batch(ctx, d_pcm, uo, fo, d_out, out_f64); }
    if (!odev) AKU_CUDA(cudaMemcpyAsync(out, d_out, (size_t)fo[n_utts] * dim * esz, cudaMemcpyDeviceToHost, ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  }
}

int akugpu_features(akugpu_ctx *ctx, const int16_t *pcm, const int64_t *utt_offsets, int n_utts, void *out, int out_f64,
                    int64_t *frame_offsets)
{
  API_BEGIN
  features_impl(ctx, pcm, utt_offsets, n_utts, out, out_f64, frame_offsets, false);
  API_END
}

int akugpu_features_pre(akugpu_ctx *ctx, const float *rows, const int64_t *row_offsets, int n_utts, void *out, int out_f64,
                        int64_t *frame_offsets)
{
  API_BEGIN
  features_impl(ctx, rows, row_offsets, n_utts, out, out_f64, frame_offsets, true);
  API_END
}

static void features_range_impl(akugpu_ctx *ctx, const void *pcm, int64_t n_samples, int start_frame, int end_frame,
                                const char *module_name, void *out, int out_f64, int *dim_out, bool want_pre)
{
  require_frontend(ctx);
  const size_t unit = base_unit_bytes(ctx, want_pre);
  int target = -1;
  if (module_name && module_name[0]) {
    for (size_t i = 0; i < ctx->fe.mods.size(); i++)
      if (ctx->fe.mods[i].name == module_name) target = (int)i;
    if (target < 0) throw Error(AKUGPU_E_ARG, std::string("unknown module requested: ") + module_name);
  }
  const int dim = ctx->fe.mods[target < 0 ? ctx->fe.last : target].dim;
  if (dim_out) *dim_out = dim;
  if (out && end_frame > start_frame) {
    if (!pcm) throw Error(AKUGPU_E_ARG, "pcm is NULL");
    const size_t esz = out_f64 ? 8 : 4;
    const int64_t n = end_frame - start_frame;
    const void *d_pcm = to_device(ctx, pcm, (size_t)n_samples * unit, ctx->d_pcm);
    const bool odev = is_device_ptr(out);
    void *d_out = out;
    if (!odev) { ctx->d_feats.reserve((size_t)n * dim * esz); d_out = ctx->d_feats.p; }
    { StageScope sc(ctx, 0); frontend_run_range(ctx, d_pcm, n_samples, start_frame, end_frame, target, d_out, out_f64); }
    if (!odev) AKU_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * dim * esz, cudaMemcpyDeviceToHost, ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  }
}

int akugpu_features_range(akugpu_ctx *ctx, const int16_t *pcm, int64_t n_samples, int start_frame, int end_frame,
                          const char *module_name, void *out, int out_f64, int *dim_out)
{
  API_BEGIN
  features_range_impl(ctx, pcm, n_samples, start_frame, end_frame, module_name, out, out_f64, dim_out, false);
  API_END
}

int akugpu_features_pre_range(akugpu_ctx *ctx, const float *rows, int64_t n_rows, int start_frame, int end_frame,
                              const char *module_name, void *out, int out_f64, int *dim_out)
{
  API_BEGIN
  features_range_impl(ctx, rows, n_rows, start_frame, end_frame, module_name, out, out_f64, dim_out, true);
  API_END
}

// ---- model -------------------------------------------------------------------------
int akugpu_model_read(akugpu_ctx *ctx, const char *base)
{
  API_BEGIN
  if (!base) throw Error(AKUGPU_E_ARG, "base is NULL");
  ctx->have_model = false;
  const std::string b(base);
  model_read_files(b + ".gk", b + ".mc", b + ".ph", ctx->hm);   // HmmSet::read_all, aku/HmmSet.cc:352-357
  model_pack(ctx);
  API_END
}

int akugpu_model_read_files(akugpu_ctx *ctx, const char *gk_path, const char *mc_path, const char *ph_path)
{
  API_BEGIN
  if (!gk_path || !mc_path || !ph_path) throw Error(AKUGPU_E_ARG, "a model path is NULL");
  ctx->have_model = false;
  model_read_files(gk_path, mc_path, ph_path, ctx->hm);
  model_pack(ctx);
  API_END
}

int akugpu_model_load_diag(akugpu_ctx *ctx, int n_states, int n_gauss, int dim, const int32_t *mix_offsets,
                           const int32_t *mix_gauss, const double *mix_weight, const double *means, const double *covs)
{
  API_BEGIN
  if (n_states < 0 || n_gauss < 0 || dim <= 0 || !mix_offsets || (n_gauss && (!means || !covs)))
    throw Error(AKUGPU_E_ARG, "bad model sizes / NULL arrays");
  ctx->have_model = false;
  HostModel &hm = ctx->hm;
  hm.S = n_states; hm.G = n_gauss; hm.D = dim;
  hm.mix_off.assign(mix_offsets, mix_offsets + n_states + 1);
  if (hm.mix_off[0] != 0) throw Error(AKUGPU_E_ARG, "mix_offsets[0] must be 0");
  const int K = hm.mix_off[n_states];
  for (int s = 0; s < n_states; s++)
    if (hm.mix_off[s + 1] < hm.mix_off[s]) throw Error(AKUGPU_E_ARG, "mix_offsets must be non-decreasing");
  if (K && (!mix_gauss || !mix_weight)) throw Error(AKUGPU_E_ARG, "mix_gauss/mix_weight is NULL");
  hm.mix_gauss.assign(mix_gauss, mix_gauss + K);
  hm.mix_w.assign(mix_weight, mix_weight + K);
  for (int k = 0; k < K; k++)
    if (hm.mix_gauss[k] < 0 || hm.mix_gauss[k] >= n_gauss) throw Error(AKUGPU_E_ARG, fmt("mix_gauss[%d] out of range", k));
  for (int s = 0; s < n_states; s++) {   // Mixture::normalize_weights, aku/Distributions.cc:2068-2075
    double sum = 0;
    for (int k = hm.mix_off[s]; k < hm.mix_off[s + 1]; k++) sum += hm.mix_w[k];
    for (int k = hm.mix_off[s]; k < hm.mix_off[s + 1]; k++) hm.mix_w[k] /= sum;
  }
  hm.mean.assign(means, means + (size_t)n_gauss * dim);
  hm.cov.assign(covs, covs + (size_t)n_gauss * dim);
  hm.full_index.clear(); hm.full_cov.clear(); hm.n_full = 0;
  hm.clear_clustering();
  hm.clear_cmllr();
  hm.ph_label.clear();
  hm.ph_states.clear();
  model_pack(ctx);
  API_END
}

int akugpu_model_load_full(akugpu_ctx *ctx, int n_states, int n_gauss, int dim, const int32_t *mix_offsets,
                           const int32_t *mix_gauss, const double *mix_weight, const double *means,
                           const double *full_covs)
{
  API_BEGIN
  if (n_states < 0 || n_gauss <= 0 || dim <= 0 || !mix_offsets || !means || !full_covs)
    throw Error(AKUGPU_E_ARG, "bad model sizes / NULL arrays");
  std::vector<double> zeros((size_t)n_gauss * dim, 1.0);
  int rc = akugpu_model_load_diag(ctx, n_states, n_gauss, dim, mix_offsets, mix_gauss, mix_weight, means, zeros.data());
  if (rc != 0) return rc;
  HostModel &hm = ctx->hm;
  hm.full_index.resize(n_gauss);
  for (int g = 0; g < n_gauss; g++) hm.full_index[g] = g;
  hm.full_cov.assign(full_covs, full_covs + (size_t)n_gauss * dim * dim);
  hm.n_full = n_gauss;
  ctx->have_model = false;
  model_pack(ctx);
  API_END
}

// ---- Gaussian clustering (phone_probs -C / --eval-minc / --eval-ming) ----------------
int akugpu_model_read_clustering(akugpu_ctx *ctx, const char *gcl_path)
{
  API_BEGIN
  require_model(ctx);
  if (!gcl_path) throw Error(AKUGPU_E_ARG, "gcl_path is NULL");
  model_read_clustering(gcl_path, ctx->hm);
  model_pack_clustering(ctx);
  API_END
}

int akugpu_model_set_clustering(akugpu_ctx *ctx, int n_clusters, const int32_t *gauss_index, const int32_t *cluster_index,
                                int64_t n_pairs)
{
  API_BEGIN
  require_model(ctx);
  if (n_pairs < 0 || (n_pairs && (!gauss_index || !cluster_index))) throw Error(AKUGPU_E_ARG, "bad n_pairs / NULL arrays");
  model_set_clustering(ctx->hm, n_clusters, gauss_index, cluster_index, n_pairs);
  model_pack_clustering(ctx);
  API_END
}

int akugpu_model_set_clustering_min_evals(akugpu_ctx *ctx, double min_clusters, double min_gaussians)
{
  API_BEGIN
  require_model(ctx);
  HostModel &hm = ctx->hm;                 // HmmSet::set_clustering_min_evals, aku/HmmSet.cc:1360-1366
  hm.eval_min_clusters = (int)(min_clusters * hm.n_clusters);
  hm.eval_min_gaussians = (int)(min_gaussians * hm.G);
  hm.use_clustering = true;
  API_END
}

int akugpu_model_use_clustering(akugpu_ctx *ctx, int on)
{
  API_BEGIN
  require_model(ctx);
  ctx->hm.use_clustering = on != 0;        // PDFPool::set_use_clustering, aku/Distributions.hh:238
  API_END
}

// ---- model-level CMLLR, global transform (phone_probs -S with a `model cmllr` entry, unitmode UNIT_NO) ----
int akugpu_model_set_cmllr(akugpu_ctx *ctx, const double *W)
{
  API_BEGIN
  require_model(ctx);
  HostModel &hm = ctx->hm;
  const int D = hm.D;
  if (!W && !hm.cmllr_on) return AKUGPU_OK;
  double factor = 1.0;
  if (W) {
    // AdaptedGaussian::compute_likelihood multiplies by AdaptedFeatureVector::determinant_A =
    // |LinearAlgebra::full_matrix_determinant(A)| (aku/ModelModules.hh:141,170).  That routine LU-factorises a copy but
    // multiplies the diagonal of A itself (aku/LinearAlgebra.cc:74-86), so the factor is |prod_i A(i,i)|, not |det A|;
    // the reference's numbers are the contract, hence the same factor here.
    for (int i = 0; i < D; i++) factor *= W[(size_t)i * (D + 1) + 1 + i];
    factor = fabs(factor);
    if (!(factor > 0) || std::isinf(factor)) throw Error(AKUGPU_E_ARG, "CMLLR: the diagonal of A must be finite and non-zero");
    for (size_t i = 0; i < (size_t)D * (D + 1); i++)
      if (!std::isfinite(W[i])) throw Error(AKUGPU_E_ARG, "CMLLR: W has a non-finite element");
  }
  if (hm.mix_w_base.empty()) hm.mix_w_base = hm.mix_w;
  ctx->have_model = false;
  hm.n_tr = 0; hm.g_tr.clear(); hm.tr_Ab.clear();        // a global transform replaces regression-class ones
  if (W) {
    hm.cmllr_W.assign(W, W + (size_t)D * (D + 1));
    std::vector<double> Ab((size_t)D * D + D);
    for (int i = 0; i < D; i++) {
      for (int j = 0; j < D; j++) Ab[(size_t)i * D + j] = W[(size_t)i * (D + 1) + 1 + j];
      Ab[(size_t)D * D + i] = W[(size_t)i * (D + 1)];
    }
    ctx->d_cmllr.reserve(Ab.size() * sizeof(double));
    AKU_CUDA(cudaMemcpyAsync(ctx->d_cmllr.p, Ab.data(), Ab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < hm.mix_w.size(); k++) hm.mix_w[k] = hm.mix_w_base[k] * factor;
    hm.cmllr_on = true;
  } else {
    hm.mix_w = hm.mix_w_base;
    hm.clear_cmllr();
  }
  model_pack(ctx);          // every scorer image carries the weights
  API_END
}

// ---- model-level CMLLR with regression classes (unitmode UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN) ----
static std::string center_phone(const std::string &label)      // Hmm::get_center_phone, aku/HmmSet.cc:22-40
{
  const size_t p1 = label.find_last_of('-'), p2 = label.find_first_of('+');
  std::string t;
  if (p1 != std::string::npos && p2 != std::string::npos) { if (p2 > p1 + 1) t = label.substr(p1 + 1, p2 - p1 - 1); }
  else if (p1 != std::string::npos) t = label.substr(p1 + 1);
  else if (p2 != std::string::npos) t = label.substr(0, p2);
  else t = label;
  if (t.empty()) throw Error(AKUGPU_E_MODEL, "Invalid phone label " + label);
  return t;
}

int akugpu_model_set_cmllr_units(akugpu_ctx *ctx, const char *unitmode, int n_transforms, const char *const *units, const double *W)
{
  API_BEGIN
  require_model(ctx);
  if (!unitmode) throw Error(AKUGPU_E_ARG, "unitmode is NULL");
  const std::string um(unitmode);
  if (um == "UNIT_NO") {
    if (n_transforms > 1) throw Error(AKUGPU_E_ARG, "ERROR: speaker can only contain one transform when UNIT_NO (global transform) is set");
    return akugpu_model_set_cmllr(ctx, n_transforms == 1 ? W : NULL);
  }
  if (um != "UNIT_PHONE" && um != "UNIT_MIX" && um != "UNIT_GAUSSIAN") throw Error(AKUGPU_E_ARG, "unknown unitmode " + um);
  if (n_transforms == 0) return akugpu_model_set_cmllr(ctx, NULL);
  if (n_transforms < 0 || !units || !W) throw Error(AKUGPU_E_ARG, "bad n_transforms / NULL arrays");
  HostModel &hm = ctx->hm;
  if (hm.n_full > 0) throw Error(AKUGPU_E_MODEL, "CMLLR regression classes are provided for diagonal pools only");
  if (clustering_on(ctx) || hm.n_clusters > 0)
    throw Error(AKUGPU_E_STATE, "CMLLR regression classes together with Gaussian clustering are not provided");
  const int D = hm.D, G = hm.G;
  if (um == "UNIT_PHONE" && hm.ph_label.empty())
    throw Error(AKUGPU_E_STATE, "UNIT_PHONE transforms need the phone table of a model read from files (akugpu_model_read)");
  // ConstrainedMllr::load_transform (aku/ModelModules.cc:172-236) visits the transforms in the order of their std::map
  // key -- the vector of unit strings -- and wraps every Gaussian of a transform anew: the last claimant wins
  std::vector<std::vector<std::string>> keys(n_transforms);
  for (int t = 0; t < n_transforms; t++) {
    if (!units[t]) throw Error(AKUGPU_E_ARG, "units[t] is NULL");
    std::istringstream is(units[t]);
    std::string u;
    while (is >> u) keys[t].push_back(u);
    if (keys[t].empty()) throw Error(AKUGPU_E_ARG, "a regression-class transform needs at least one unit");
    for (size_t i = 0; i < (size_t)D * (D + 1); i++)
      if (!std::isfinite(W[(size_t)t * D * (D + 1) + i])) throw Error(AKUGPU_E_ARG, "CMLLR: W has a non-finite element");
  }
  std::vector<int> order(n_transforms);
  for (int t = 0; t < n_transforms; t++) order[t] = t;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] < keys[b]; });
  for (int t = 1; t < n_transforms; t++)
    if (keys[order[t]] == keys[order[t - 1]]) throw Error(AKUGPU_E_ARG, "two transforms list the same units");   // one map entry in the reference
  std::vector<int32_t> g_tr(G, -1);
  std::vector<double> Ab((size_t)n_transforms * (D * D + D)), factor(n_transforms, 1.0);
  auto mixture_gaussians = [&](long m, std::vector<int> &out) {
    if (m < 0 || m >= hm.S) throw Error(AKUGPU_E_ARG, fmt("CMLLR: mixture %ld out of range", m));
    for (int k = hm.mix_off[m]; k < hm.mix_off[m + 1]; k++) out.push_back(hm.mix_gauss[k]);
  };
  for (int r = 0; r < n_transforms; r++) {
    const int t = order[r];
    const double *Wt = W + (size_t)t * D * (D + 1);
    double f = 1.0;
    for (int i = 0; i < D; i++) {
      for (int j = 0; j < D; j++) Ab[(size_t)r * (D * D + D) + (size_t)i * D + j] = Wt[(size_t)i * (D + 1) + 1 + j];
      Ab[(size_t)r * (D * D + D) + (size_t)D * D + i] = Wt[(size_t)i * (D + 1)];
      f *= Wt[(size_t)i * (D + 1) + 1 + i];          // the reference's "determinant": see akugpu_model_set_cmllr
    }
    factor[r] = fabs(f);
    if (!(factor[r] > 0) || std::isinf(factor[r])) throw Error(AKUGPU_E_ARG, "CMLLR: the diagonal of A must be finite and non-zero");
    std::vector<int> gs;
    if (um == "UNIT_PHONE") {          // RegClassTree::UnitPhoneme::get_gaussians, aku/RegClassTree.cc:302-322
      for (size_t h = 0; h < hm.ph_label.size(); h++) {
        if (std::find(keys[t].begin(), keys[t].end(), center_phone(hm.ph_label[h])) == keys[t].end()) continue;
        for (int32_t st : hm.ph_states[h]) mixture_gaussians(st, gs);
      }
    } else if (um == "UNIT_MIX") {     // UnitMixture::get_gaussians :368-385: units that are not numbers are skipped
      for (const std::string &u : keys[t]) {
        char *end = nullptr;
        const long m = strtol(u.c_str(), &end, 10);
        if (end == u.c_str() || *end) continue;
        mixture_gaussians(m, gs);
      }
    } else {                           // UnitGaussian::get_gaussians :444-454
      for (const std::string &u : keys[t]) {
        const long g = strtol(u.c_str(), nullptr, 10);
        if (g < 0 || g >= G) throw Error(AKUGPU_E_ARG, fmt("CMLLR: Gaussian %ld out of range", g));
        gs.push_back((int)g);
      }
    }
    for (int g : gs) g_tr[g] = r;
  }
  if (hm.mix_w_base.empty()) hm.mix_w_base = hm.mix_w;
  ctx->have_model = false;
  hm.cmllr_W.clear();
  hm.n_tr = n_transforms;
  hm.g_tr = g_tr;
  hm.tr_Ab = Ab;
  ctx->d_cmllr.reserve(Ab.size() * sizeof(double));
  AKU_CUDA(cudaMemcpyAsync(ctx->d_cmllr.p, Ab.data(), Ab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  for (size_t k = 0; k < hm.mix_w.size(); k++) {
    const int tr = g_tr[hm.mix_gauss[k]];
    hm.mix_w[k] = tr >= 0 ? hm.mix_w_base[k] * factor[tr] : hm.mix_w_base[k];
  }
  hm.cmllr_on = true;
  model_pack(ctx);
  API_END
}

int akugpu_model_num_states(akugpu_ctx *ctx) { return (ctx && ctx->have_model) ? ctx->hm.S : AKUGPU_E_STATE; }
int akugpu_model_dim(akugpu_ctx *ctx) { return (ctx && ctx->have_model) ? ctx->hm.D : AKUGPU_E_STATE; }
int akugpu_model_num_gaussians(akugpu_ctx *ctx) { return (ctx && ctx->have_model) ? ctx->hm.G : AKUGPU_E_STATE; }

// ---- scoring -----------------------------------------------------------------------
// Shared body of akugpu_gmm_score / akugpu_gmm_logprobs.  tiny > 0: decoder log-prob mode, out = float [F x S] of
// (float) log(max(likelihood, tiny)) whatever the precision.
static void gmm_score_impl(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, int precision, void *out, double tiny)
{
  require_model(ctx);
  if (n_frames < 0 || (n_frames && (!feats || !out))) throw Error(AKUGPU_E_ARG, "bad n_frames / NULL buffers");
  if (precision != AKUGPU_F32 && precision != AKUGPU_F64) throw Error(AKUGPU_E_ARG, "precision must be AKUGPU_F32 or AKUGPU_F64");
  const int S = ctx->hm.S, D = ctx->hm.D;
  if (n_frames == 0 || S == 0) return;
  const bool logmode = tiny > 0;
  // streaming regime (the decoder's per-frame feed): one launch, features in the parameter block, results and the
  // completion flag straight into mapped host memory.  false = a feature left the fp16 range: the general path below
  // redoes the call (and finds the same overflow itself)
  if (session_applicable(ctx, precision, n_frames, feats, out) &&
      session_score(ctx, feats, feats_f64, n_frames, (float *)out, logmode ? 1 : 0, logmode ? (float)log(tiny) : 0.f))
    return;
  session_quiesce(ctx);
  if (stream_applicable(ctx, precision, n_frames) &&
      stream_score(ctx, feats, feats_f64, n_frames, (float *)out, logmode ? 1 : 0, logmode ? (float)log(tiny) : 0.f))
    return;
  const void *d_feats = to_device(ctx, feats, (size_t)n_frames * D * (feats_f64 ? 8 : 4), ctx->d_feats);
  d_feats = adapt_feats(ctx, d_feats, feats_f64, n_frames);
  const size_t esz = (precision == AKUGPU_F64 && !logmode) ? 8 : 4;        // element size of the result
  const bool odev = is_device_ptr(out);
  uint8_t *d_out = (uint8_t *)out;
  if (!odev) { ctx->d_tmp.reserve((size_t)n_frames * S * esz); d_out = ctx->d_tmp.as<uint8_t>(); }
  ctx->tc16_suspended = false;
  bool redo = false;
  do {   // a second pass only when the fp16x2 scorer met a feature outside its range (see tc16_needs_redo)
  const int use_tc = tc_mode(ctx, precision);
  const bool dbl = precision == AKUGPU_F64 || (ctx->hm.n_full > 0 && !use_tc) || clustering_on(ctx) || ctx->hm.n_tr > 0;   // scored in double
  const bool want_f32 = esz == 4;
  const int64_t chunk = pick_chunk(ctx, n_frames, use_tc);
  ctx->d_sll.reserve((size_t)S * chunk * (dbl ? 8 : 4));
  if (dbl && want_f32) ctx->d_lna[0].reserve((size_t)S * chunk * 4);
  for (int64_t c0 = 0; c0 < n_frames; c0 += chunk) {
    const int64_t c1 = std::min(n_frames, c0 + chunk);
    StageScope sc(ctx, 1);
    if (dbl) {
      if (ctx->hm.n_full > 0) launch_gmm_full_f64(ctx, d_feats, feats_f64, c0, c1, ctx->d_sll.as<double>(), chunk);
      else launch_gmm_f64(ctx, d_feats, feats_f64, c0, c1, ctx->d_sll.as<double>(), chunk);
      if (want_f32) {
        launch_lin_to_log_f32(ctx, ctx->d_sll.as<double>(), (int64_t)S * chunk, ctx->d_lna[0].as<float>(), logmode ? tiny : 0.0);
        launch_transpose_f32(ctx, ctx->d_lna[0].as<float>(), chunk, S, c1 - c0, (float *)(d_out + (size_t)c0 * S * 4));
      } else {
        launch_transpose_f64(ctx, ctx->d_sll.as<double>(), chunk, S, c1 - c0, (double *)(d_out + (size_t)c0 * S * 8));
      }
    } else {
      if (use_tc == 2) launch_gmm_tc16(ctx, d_feats, feats_f64, c0, c1, ctx->d_sll.as<float>(), chunk, nullptr);
      else if (use_tc) launch_gmm_tc(ctx, d_feats, feats_f64, c0, c1, ctx->d_sll.as<float>(), chunk, nullptr);
      if (!use_tc || (ctx->ptc16.ready && ctx->ptc16.hybrid))     // whole model, or the states left out of the expanded form
        launch_gmm_f32(ctx, d_feats, feats_f64, c0, c1, ctx->d_sll.as<float>(), chunk);
      if (logmode) launch_floor_f32(ctx, ctx->d_sll.as<float>(), (int64_t)S * chunk, (float)log(tiny));
      launch_transpose_f32(ctx, ctx->d_sll.as<float>(), chunk, S, c1 - c0, (float *)(d_out + (size_t)c0 * S * 4));
    }
  }
  redo = !redo && tc16_needs_redo(ctx, use_tc);
  ctx->tc16_suspended = redo;
  } while (redo);
  if (!odev) AKU_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n_frames * S * esz, cudaMemcpyDeviceToHost, ctx->stream));
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
}

int akugpu_gmm_score(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, int precision, void *out)
{
  API_BEGIN_KEEP
  gmm_score_impl(ctx, feats, feats_f64, n_frames, precision, out, 0.0);
  API_END
}

int akugpu_gmm_logprobs(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, int precision, double tiny,
                        float *out)
{
  API_BEGIN_KEEP
  if (!(tiny > 0)) throw Error(AKUGPU_E_ARG, "tiny must be > 0 (decode-stream uses 1e-30, phone_probs 1e-50)");
  gmm_score_impl(ctx, feats, feats_f64, n_frames, precision, out, tiny);
  API_END
}

int akugpu_gmm_lna(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, int precision, int lnabytes,
                   int normalize, uint8_t *out)
{
  API_BEGIN
  require_model(ctx);
  if (n_frames < 0 || (n_frames && !feats)) throw Error(AKUGPU_E_ARG, "bad n_frames / NULL feats");
  const void *d_feats = to_device(ctx, feats, (size_t)n_frames * ctx->hm.D * (feats_f64 ? 8 : 4), ctx->d_feats);
  d_feats = adapt_feats(ctx, d_feats, feats_f64, n_frames);
  score_to_lna(ctx, d_feats, feats_f64, n_frames, precision, lnabytes, normalize, out, nullptr);
  API_END
}

int akugpu_phone_probs_ex(akugpu_ctx *ctx, const int16_t *pcm, const int64_t *utt_offsets, int n_utts, int precision,
                          int lnabytes, int normalize, uint8_t *out, int64_t *frame_offsets, uint64_t *checksum_out,
                          uint64_t *utt_checksums)
{
  API_BEGIN
  require_frontend(ctx);
  require_model(ctx);
  base_unit_bytes(ctx, false);       // PCM in: needs an audiofile base module
  const int dim = ctx->fe.mods[ctx->fe.last].dim;
  if (dim != ctx->hm.D)              // aku/phone_probs.cc:119-124
    throw Error(AKUGPU_E_STATE, fmt("Gaussian dimension is %d but feature dimension is %d.", ctx->hm.D, dim));
  std::vector<int64_t> uo, fo;
  frame_offsets_of(ctx, utt_offsets, n_utts, uo, fo);
  if (frame_offsets) memcpy(frame_offsets, fo.data(), fo.size() * sizeof(int64_t));
  if (n_utts == 0) { if (checksum_out) *checksum_out = 0; return AKUGPU_OK; }
  if (!pcm) throw Error(AKUGPU_E_ARG, "pcm is NULL");
  const int feats_f64 = precision == AKUGPU_F64;
  const int16_t *d_pcm = (const int16_t *)to_device(ctx, pcm, (size_t)uo[n_utts] * 2, ctx->d_pcm);
  ctx->d_feats.reserve((size_t)fo[n_utts] * dim * (feats_f64 ? 8 : 4));
  { StageScope sc(ctx, 0); frontend_run_batch(ctx, d_pcm, uo, fo, ctx->d_feats.p, feats_f64); }
  const void *d_feats = adapt_feats(ctx, ctx->d_feats.p, feats_f64, fo[n_utts]);
  score_to_lna(ctx, d_feats, feats_f64, fo[n_utts], precision, lnabytes, normalize, out, checksum_out, &fo, utt_checksums);
  API_END
}

int akugpu_phone_probs(akugpu_ctx *ctx, const int16_t *pcm, const int64_t *utt_offsets, int n_utts, int precision,
                       int lnabytes, int normalize, uint8_t *out, int64_t *frame_offsets, uint64_t *checksum_out)
{
  return akugpu_phone_probs_ex(ctx, pcm, utt_offsets, n_utts, precision, lnabytes, normalize, out, frame_offsets, checksum_out, NULL);
}

int akugpu_lna_header(int n_states, int lnabytes, uint8_t out5[5])
{
  if (!out5) return AKUGPU_E_ARG;
  out5[0] = (n_states >> 24) & 255; out5[1] = (n_states >> 16) & 255; out5[2] = (n_states >> 8) & 255; out5[3] = n_states & 255;
  out5[4] = (uint8_t)lnabytes;
  return AKUGPU_OK;
}

int akugpu_set_chunk_frames(akugpu_ctx *ctx, int64_t frames)
{
  API_BEGIN
  if (frames != 0 && frames < 128) throw Error(AKUGPU_E_ARG, "chunk must be 0 (auto) or >= 128 frames");
  ctx->chunk_frames = frames;
  API_END
}

int akugpu_set_scorer_variant(akugpu_ctx *ctx, int variant)
{
  API_BEGIN
  if (variant < 0 || variant > 4) throw Error(AKUGPU_E_ARG, "variant must be 0..4");
  ctx->scorer_variant = variant;
  if (ctx->have_model) model_pack(ctx);
  API_END
}

double akugpu_model_expanded_form_q(akugpu_ctx *ctx)
{
  if (!ctx || !ctx->have_model) return -1.0;
  return std::max(ctx->ptc16.q_max, ctx->ptc.q_max);
}

int akugpu_scorer_in_use(akugpu_ctx *ctx)
{
  if (!ctx || !ctx->have_model) return AKUGPU_E_STATE;
  if (ctx->hm.use_clustering && ctx->hm.n_clusters > 0) return 0;
  if (ctx->ptc16.ready) return ctx->ptc16.hybrid ? 5 : (ctx->ptc16.stream ? 4 : 3);
  if (ctx->ptc.ready) return 2;
  return ctx->hm.n_full > 0 ? 0 : 1;
}

int akugpu_set_streaming(akugpu_ctx *ctx, int enable)
{
  API_BEGIN
  ctx->streaming_enabled = enable != 0;
  API_END
}

int akugpu_stream_open(akugpu_ctx *ctx, double idle_ms)
{
  API_BEGIN
  require_model(ctx);
  if (!(idle_ms > 0)) idle_ms = 100.0;
  StreamState &st = ctx->stream_state;
  st.session_idle_ms = idle_ms;
  st.session_want = true;
  if (!session_applicable(ctx, AKUGPU_F32, 1, nullptr, nullptr)) {
    st.session_want = false;
    throw Error(AKUGPU_E_STATE, "the resident scorer serves diagonal models of the fp16x2 tensor-core scorer (no clustering, no model-level CMLLR, no hybrid split)");
  }
  try {
    session_launch(ctx);
  } catch (...) {
    st.session_want = false;
    throw;
  }
  API_END
}

int akugpu_stream_logprobs(akugpu_ctx *ctx, const float *feats, int n_frames, double tiny, const float **rows)
{
  API_BEGIN_KEEP
  require_model(ctx);
  if (!feats || !rows || n_frames < 1 || n_frames > STREAM_MAX_FRAMES) throw Error(AKUGPU_E_ARG, "akugpu_stream_logprobs: 1..16 frames of host float features, rows != NULL");
  if (!ctx->stream_state.session_want) throw Error(AKUGPU_E_STATE, "akugpu_stream_logprobs: no session (akugpu_stream_open)");
  const bool logmode = tiny > 0;
  const float floor_at = logmode ? (float)log(tiny) : 0.f;
  if (stream_applicable(ctx, AKUGPU_F32, n_frames) && session_score(ctx, feats, 0, n_frames, nullptr, logmode ? 1 : 0, floor_at)) {
    *rows = session_rows(ctx);
  } else {
    // more frames than the model allows per message, or a feature outside the fp16 range: the general path, into a
    // buffer of the context
    ctx->stream_fallback.resize((size_t)n_frames * ctx->hm.S);
    gmm_score_impl(ctx, feats, 0, n_frames, AKUGPU_F32, ctx->stream_fallback.data(), logmode ? tiny : 0.0);
    *rows = ctx->stream_fallback.data();
  }
  API_END
}

int akugpu_stream_latency(akugpu_ctx *ctx, const float *feats, int n_frames, double tiny, int n_calls, double out_us[4])
{
  API_BEGIN_KEEP
  require_model(ctx);
  if (!feats || !out_us || n_frames < 1 || n_calls < 1) throw Error(AKUGPU_E_ARG, "akugpu_stream_latency: bad arguments");
  std::vector<double> us((size_t)n_calls);
  std::vector<float> scratch((size_t)n_frames * ctx->hm.S);
  const bool logmode = tiny > 0;
  const float floor_at = logmode ? (float)log(tiny) : 0.f;
  for (int i = 0; i < n_calls; i++) {
    const auto t0 = std::chrono::steady_clock::now();
    // what akugpu_stream_logprobs does inside an open session; one launch per call (akugpu_gmm_logprobs) otherwise
    if (!(ctx->stream_state.session_want && n_frames <= STREAM_MAX_FRAMES && stream_applicable(ctx, AKUGPU_F32, n_frames) &&
          session_score(ctx, feats, 0, n_frames, nullptr, logmode ? 1 : 0, floor_at)))
      gmm_score_impl(ctx, feats, 0, n_frames, AKUGPU_F32, scratch.data(), logmode ? tiny : 0.0);
    us[(size_t)i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  }
  double sum = 0;
  for (double v : us) sum += v;
  std::sort(us.begin(), us.end());
  out_us[0] = sum / n_calls;
  out_us[1] = us[(size_t)(n_calls / 2)];
  out_us[2] = us[(size_t)std::min<int64_t>(n_calls - 1, (int64_t)(0.99 * n_calls))];
  out_us[3] = us.back();
  API_END
}

int akugpu_stream_close(akugpu_ctx *ctx)
{
  API_BEGIN          /* ends the kernel */
  ctx->stream_state.session_want = false;
  session_release_device(ctx);
  API_END
}

int akugpu_stream_stats(akugpu_ctx *ctx, int64_t out[8])
{
  if (!ctx || !out) return AKUGPU_E_ARG;
  const StreamState &st = ctx->stream_state;
  out[0] = st.session_want ? 1 : 0;
  out[1] = st.session_live ? 1 : 0;
  out[2] = st.session_launches;
  out[3] = st.session_calls;
  const volatile unsigned int *f = reinterpret_cast<const volatile unsigned int *>(st.host);
  out[4] = f ? f[4] : 0;
  out[5] = f ? f[5] : 0;
  out[6] = f ? f[6] : 0;
  out[7] = f ? f[7] : 0;
  return AKUGPU_OK;
}

int akugpu_stream_probe(akugpu_ctx *ctx, double out[8])
{
  API_BEGIN
  require_model(ctx);
  if (!out) throw Error(AKUGPU_E_ARG, "out is NULL");
  stream_probe(ctx, out);
  API_END
}

int akugpu_pipe_rates(akugpu_ctx *ctx, double out[8])
{
  API_BEGIN
  if (!out) throw Error(AKUGPU_E_ARG, "out is NULL");
  pipe_rates(ctx, out);
  API_END
}

}  // extern "C"
