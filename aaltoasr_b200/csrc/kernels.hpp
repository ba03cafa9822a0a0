// kernels.hpp -- launchers shared between the translation units of libakugpu.so.
#pragma once
#include "ctx.hpp"
#include <algorithm>

namespace akugpu {

// gmm_kernels.cu
size_t gmm_f32_smem_bytes(const PackedF32 &p);
int64_t gmm_wave_frames(akugpu_ctx *ctx);
int gmm_frame_tile();
void launch_gmm_f32(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, float *sll,
                    int64_t ldF);
void launch_gmm_f64(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, double *lin,
                    int64_t ldF);
void launch_transpose_f32(akugpu_ctx *ctx, const float *in, int64_t ldF, int S, int64_t F, float *out);
void launch_transpose_f64(akugpu_ctx *ctx, const double *in, int64_t ldF, int S, int64_t F, double *out);
void pipe_rates(akugpu_ctx *ctx, double out[8]);

// gmm_full.cu
void launch_gmm_full_f64(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, double *lin,
                         int64_t ldF);
void launch_lin_to_log_f32(akugpu_ctx *ctx, const double *lin, int64_t n, float *out, double tiny = 0.0);
void launch_floor_f32(akugpu_ctx *ctx, float *x, int64_t n, float floor_at);
// out[f] = A in[f] + b for F feature rows (float or double, accumulated in double); Ab = A [D x D] row-major, then b [D]
void launch_affine_rows(akugpu_ctx *ctx, const void *in, int is_f64, int64_t F, int D, const double *Ab, void *out);

// gmm_tc.cu (tensor-core scorer, experimental)
void model_pack_tc(akugpu_ctx *ctx);
int64_t gmm_tc_wave_frames(akugpu_ctx *ctx);
// Returns true when the launch also produced the per-frame normaliser norm[frame] = {max, log1p(sum of the others)}
// (only when one CTA sweeps every component tile of its frames).
bool launch_gmm_tc(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, float *sll, int64_t ldF,
                   float2 *norm);
// gmm_tc16.cu (fp16x2 tensor-core scorer for diagonal pools; A' built in the kernel and resident in shared memory)
bool tc16_supported(const HostModel &hm);
void model_pack_tc16(akugpu_ctx *ctx);
int64_t gmm_tc16_wave_frames(akugpu_ctx *ctx);
bool launch_gmm_tc16(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, float *sll, int64_t ldF,
                     float2 *norm);
// true (and clears the flag) when a launch since the last call met a feature outside the fp16 range of the scaled terms
bool gmm_tc16_overflowed(akugpu_ctx *ctx);
bool host_cholesky(const std::vector<double> &A, int n, std::vector<double> &Lw);
void host_lu_inverse(const std::vector<double> &M, int n, std::vector<double> &inv);

// gmm_stream.cu (streaming-regime scorer: <= STREAM_MAX_FRAMES frames against the whole model in one launch)
bool stream_applicable(akugpu_ctx *ctx, int precision, int64_t n_frames);
bool stream_score(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, float *out, int use_floor, float floor_at);
void stream_probe(akugpu_ctx *ctx, double out[8]);
// gmm_resident.cu (resident scorer: a kernel that stays on the device between calls, parameter image in shared memory)
bool session_applicable(akugpu_ctx *ctx, int precision, int64_t n_frames, const void *feats, const void *out);
bool session_score(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, float *out, int use_floor, float floor_at);
void session_launch(akugpu_ctx *ctx);
void session_quiesce(akugpu_ctx *ctx);
void session_destroy(akugpu_ctx *ctx);
void session_release_device(akugpu_ctx *ctx);   // akugpu_stream_close: another context may open a session on the device
const float *session_rows(akugpu_ctx *ctx);      // the pinned result rows of the last session call

// lna_kernels.cu
// norm_scratch: where a normaliser pass of its own writes (default: ctx->d_norm)
void launch_lna_f32(akugpu_ctx *ctx, const float *sll, int64_t ldF, int S, int64_t nf, int lnabytes, int normalize,
                    const float2 *norm, uint8_t *out, float2 *norm_scratch = nullptr);
void launch_lna_f64(akugpu_ctx *ctx, const double *lin, int64_t ldF, int S, int64_t nf, int lnabytes, int normalize,
                    uint8_t *out);
void launch_checksum(akugpu_ctx *ctx, const uint8_t *buf, int64_t nbytes, unsigned long long *acc);

// multigpu.cu: per-utterance checksum sink
void checksum_begin(akugpu_ctx *ctx, const int64_t *frame_offsets, int n_utts, int64_t rec_bytes);
void checksum_update(akugpu_ctx *ctx, const uint8_t *records, int64_t first_frame, int64_t n_frames);
void checksum_end(akugpu_ctx *ctx, uint64_t *out);

// model.cu
void model_read_clustering(const std::string &gcl_path, HostModel &hm);
void model_set_clustering(HostModel &hm, int n_clusters, const int32_t *gauss_index, const int32_t *cluster_index, int64_t n_pairs);
void model_pack_clustering(akugpu_ctx *ctx);
void model_read_files(const std::string &gk_path, const std::string &mc_path, const std::string &ph_path, HostModel &hm);
void model_pack(akugpu_ctx *ctx);

// frontend_config.cc / frontend_kernels.cu
void frontend_parse(akugpu_ctx *ctx, const std::string &text);
void frontend_set_parameters(akugpu_ctx *ctx, const std::string &module, const std::string &text);
int64_t frontend_num_frames(const Frontend &fe, int64_t n_samples);
// Computes module `target` (or the last module when target<0) for ONE utterance whose PCM is on the
// device, for frames [start,end) in reference frame numbering; writes [end-start][dim] float or double.
// d_pcm: int16 samples (audiofile base module) or float32 rows [n x dim] (`pre` base module: n_samples / utt_off count rows)
void frontend_run_range(akugpu_ctx *ctx, const void *d_pcm, int64_t n_samples, int start, int end, int target,
                        void *d_out, int out_f64);
// Batch: all utterances, frames 0..n_u-1 each, into d_out [sum n_u][dim].
void frontend_run_batch(akugpu_ctx *ctx, const void *d_pcm, const std::vector<int64_t> &utt_off,
                        const std::vector<int64_t> &frame_off, void *d_out, int out_f64);

// Opt-in to `bytes` of dynamic shared memory for `kernel` on the context's device.  The attribute is per device and the
// cache is shared by every context of the process, so it is keyed by (device, kernel) and guarded by a mutex
// (one context per host thread, several threads / devices per process).  (api.cu)
void ensure_dynamic_smem(akugpu_ctx *ctx, const void *kernel, size_t bytes);

// stage timing (api.cu)
struct StageScope {
  akugpu_ctx *ctx;
  int stage;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st = nullptr;
  int64_t l0;
  StageScope(akugpu_ctx *c, int s);
  ~StageScope();
};

}  // namespace akugpu
