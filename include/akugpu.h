/* akugpu.h -- C ABI of the B200-native acoustic front-end for AaltoASR.
 *
 * One shared library (aaltoasr_b200/libakugpu.so), plain pointers and sizes, no
 * C++/torch types, no exceptions across the boundary.  Each entry point names
 * the reference interface it replaces (paths relative to the AaltoASR tree).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (AKUGPU_E_*); the message
 *     is available from akugpu_last_error(ctx)   [reference: `throw std::string`,
 *     e.g. aku/FeatureModules.cc:334, aku/Distributions.cc:2872].
 *   - data pointers may be HOST or DEVICE memory; the library detects which with
 *     cudaPointerGetAttributes.  Host buffers are staged through pinned memory and
 *     overlapped with compute; device buffers are used in place.
 *   - one context per host thread / GPU; a context is not thread-safe (the
 *     reference classes are not re-entrant either: aku/HmmSet.hh:550-553).
 *   - frames are rows; features are row-major [frames x dim].
 *   - there is NO CPU fallback: without a CUDA device akugpu_create() fails.
 */
#ifndef AKUGPU_H
#define AKUGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct akugpu_ctx akugpu_ctx;

enum {
  AKUGPU_OK = 0,
  AKUGPU_E_CUDA = -1,      /* CUDA runtime error                                  */
  AKUGPU_E_ARG = -2,       /* bad argument / inconsistent sizes                   */
  AKUGPU_E_CONFIG = -3,    /* feature configuration not supported / malformed     */
  AKUGPU_E_MODEL = -4,     /* model files malformed / unsupported PDF type        */
  AKUGPU_E_STATE = -5,     /* call order (no model / no front-end configured)     */
  AKUGPU_E_IO = -6         /* file could not be read / written                    */
};

/* Arithmetic of the Gaussian scorer + LNA epilogue.
 * F32: throughput mode (float log-probs within 1e-4 relative of the reference).
 * F64: parity mode, follows aku/Distributions.cc:1041-1062,2079-2086 and
 *      aku/phone_probs.cc:225-259 operation by operation in double. */
enum { AKUGPU_F32 = 0, AKUGPU_F64 = 1 };

/* ---- context ----------------------------------------------------------------- */
akugpu_ctx *akugpu_create(int device);             /* NULL if no usable CUDA device */
void        akugpu_destroy(akugpu_ctx *ctx);
const char *akugpu_last_error(akugpu_ctx *ctx);    /* ctx may be NULL: last create() error */
/* Use an existing CUDA stream (cudaStream_t) for all work; NULL = context's own. */
int         akugpu_set_stream(akugpu_ctx *ctx, void *cuda_stream);
int         akugpu_synchronize(akugpu_ctx *ctx);
/* Number of kernels this library has launched on ctx since creation. */
int64_t     akugpu_launch_count(akugpu_ctx *ctx);
/* Device time (ms, CUDA events on the launching stream) spent in the named stage
 * since the last akugpu_stage_times_reset(): 0 front-end, 1 GMM scorer, 2 LNA epilogue. */
int         akugpu_stage_times(akugpu_ctx *ctx, double ms_out[3], int64_t launches_out[3]);
int         akugpu_stage_times_reset(akugpu_ctx *ctx, int enable);

/* ---- feature front-end ---------------------------------------------------------
 * Replaces FeatureGenerator::load_configuration (aku/FeatureGenerator.cc:97-219)
 * and the module classes of aku/FeatureModules.cc.  The text is the reference's
 * own `module { name .. type .. sources .. }` format.  Supported module types:
 * audiofile, pre, fft, vtln (incl. all-pass), mel, power, mel_power, dct, delta, merge, concat, normalization,
 * lin_transform, mean_subtractor, sr_norm, quanteq -- every module type of aku/FeatureModules.cc. */
int   akugpu_frontend_load_config(akugpu_ctx *ctx, const char *cfg_path);
int   akugpu_frontend_load_config_text(akugpu_ctx *ctx, const char *cfg_text);
int   akugpu_frontend_dim(akugpu_ctx *ctx);            /* FeatureGenerator::dim()         */
int   akugpu_frontend_sample_rate(akugpu_ctx *ctx);    /* FeatureGenerator::sample_rate() */
float akugpu_frontend_frame_rate(akugpu_ctx *ctx);     /* FeatureGenerator::frame_rate()  */
int   akugpu_frontend_base_is_pre(akugpu_ctx *ctx);    /* 1: the base module is `pre` (stored features), 0: audiofile */
/* Output dimension of the base module: window_width for audiofile, the configured `dim` for pre -- the value
 * PreModule::set_file compares with the file header ("The file has invalid dimension", aku/FeatureModules.cc:622-626). */
int   akugpu_frontend_base_dim(akugpu_ctx *ctx);
/* 1: the `pre` base module was configured with `legacy_file 1`: the stored-feature file starts with a ONE-byte
 * dimension instead of an int32 (aku/FeatureModules.cc:608-615).  Reading the header is the host reader's business. */
int   akugpu_frontend_pre_legacy(akugpu_ctx *ctx);
/* Number of frames the reference generates before eof() for an audio file of
 * n_samples samples (aku/FeatureModules.cc:371-424: frame f is valid iff
 * (int)(f*window_advance) + window_width + 1 <= n_samples). */
int64_t akugpu_frontend_num_frames(akugpu_ctx *ctx, int64_t n_samples);
/* FeatureModule::set_parameters for a named module (aku/FeatureModule.hh:107):
 * `text` holds `key value...` lines as in a config block (lin_transform: "matrix ...", "bias ..."; normalization:
 * "mean ...", "scale ..."; vtln: "warp_factor w" or "slapt_coef ..."; sr_norm: "speech_rate r"; quanteq: "alpha ...",
 * "gamma ...", "quant_max ..."). */
int   akugpu_frontend_set_parameters(akugpu_ctx *ctx, const char *module_name, const char *text);

/* Batch feature computation: replaces the per-frame FeatureGenerator::generate(f)
 * loop (aku/FeatureGenerator.cc:267-273) for whole utterances.
 *   pcm            int16 mono samples of all utterances, concatenated
 *   utt_offsets    [n_utts+1] sample offsets into pcm (host memory)
 *   out            [total_frames x dim] features, float (out_f64=0) or double
 *   frame_offsets  [n_utts+1] receives the frame offset of each utterance (host)
 * Frames of utterance u are 0 .. num_frames(len_u)-1, borders handled as the
 * reference does (first/last window replicated, aku/FeatureModules.cc:381-422).
 * Pass out=NULL to only fill frame_offsets. */
int akugpu_features(akugpu_ctx *ctx, const int16_t *pcm, const int64_t *utt_offsets, int n_utts,
                    void *out, int out_f64, int64_t *frame_offsets);
/* Same pipeline for an explicit frame range [start,end) of ONE utterance, frames
 * outside the file behaving as in the reference (feacat --start-frame/--end-frame,
 * aku/feacat.cc:96-110).  module_name != NULL returns that module's output
 * instead of the last module's (FeatureGenerator::module(name)->at(f)). */
int akugpu_features_range(akugpu_ctx *ctx, const int16_t *pcm, int64_t n_samples,
                          int start_frame, int end_frame, const char *module_name,
                          void *out, int out_f64, int *dim_out);

/* The same two calls for a configuration whose base module is `pre` (PreModule, aku/FeatureModules.cc:603-755: stored
 * float32 feature rows instead of audio, the format feacat -H --raw-output writes: int32 dim, then rows; the header
 * is the caller's business).  rows = all utterances' rows back to back, [n x dim] float32 with dim = the `pre`
 * module's configured dimension; row_offsets / n_rows count rows.  Frames before / after the stored rows repeat the
 * first / last one (:713-728).  With a `pre` base the int16 entry points (and akugpu_phone_probs) return AKUGPU_E_STATE,
 * and vice versa. */
int akugpu_features_pre(akugpu_ctx *ctx, const float *rows, const int64_t *row_offsets, int n_utts,
                        void *out, int out_f64, int64_t *frame_offsets);
int akugpu_features_pre_range(akugpu_ctx *ctx, const float *rows, int64_t n_rows, int start_frame, int end_frame,
                              const char *module_name, void *out, int out_f64, int *dim_out);

/* ---- acoustic model --------------------------------------------------------------
 * Replaces HmmSet::read_all / read_mc / read_ph / read_gk (aku/HmmSet.cc:157-357,
 * aku/Distributions.cc:2812-2910) and the parameter side of DiagonalGaussian
 * (aku/Distributions.cc:1132-1150,1274-1288) and Mixture (:2419-2434,2068-2075). */
int akugpu_model_read(akugpu_ctx *ctx, const char *base);   /* base.mc, base.ph, base.gk */
/* The same with the three files named separately (phone_probs -g/-m/-p, aku/phone_probs.cc:99-105). */
int akugpu_model_read_files(akugpu_ctx *ctx, const char *gk_path, const char *mc_path, const char *ph_path);
/* Direct load.  State s owns components mix_offsets[s] .. mix_offsets[s+1]-1;
 * component k refers to Gaussian mix_gauss[k] with weight mix_weight[k]
 * (weights are re-normalised per state like Mixture::normalize_weights()).
 * means/covs are [G x D]; cov <= 0 disables that dimension exactly as the
 * reference does. */
int akugpu_model_load_diag(akugpu_ctx *ctx, int n_states, int n_gauss, int dim,
                           const int32_t *mix_offsets, const int32_t *mix_gauss,
                           const double *mix_weight, const double *means, const double *covs);
/* All-full-covariance pool (FullCovarianceGaussian, aku/Distributions.cc:1467-1488,1560-1586):
 * full_covs is [G x D x D] row-major.  Pools mixing `diag` and `full` lines come through
 * akugpu_model_read.  Precision F64 scores them in double (the reference's exponential form,
 * aku/Distributions.cc:1437-1446); F32 uses the fp16 hi/lo-split tensor-core kernel on the same expanded form
 * (gmm_tc16_kernel<0>) when the pool is all-full-covariance and well conditioned, the double path otherwise. */
int akugpu_model_load_full(akugpu_ctx *ctx, int n_states, int n_gauss, int dim,
                           const int32_t *mix_offsets, const int32_t *mix_gauss,
                           const double *mix_weight, const double *means, const double *full_covs);
/* Gaussian clustering approximation (phone_probs -C file.gcl --eval-minc x --eval-ming y, aku/phone_probs.cc:73-75,112-117):
 * per frame the cluster centres are scored, the Gaussians of the best clusters are evaluated exactly until both minima are
 * reached, the others take their centre's likelihood (PDFPool::precompute_likelihoods, aku/Distributions.cc:2685-2722).
 *   akugpu_model_read_clustering           HmmSet::read_clustering (aku/HmmSet.cc:1354, aku/Distributions.cc:3115-3170):
 *                                          `n_clusters` then `gauss_index cluster_index` pairs; centres by moment matching
 *   akugpu_model_set_clustering            the same from memory (pairs exactly as they would be read)
 *   akugpu_model_set_clustering_min_evals  HmmSet::set_clustering_min_evals (aku/HmmSet.cc:1360): ratios of clusters /
 *                                          Gaussians; switches the approximation on
 *   akugpu_model_use_clustering            PDFPool::set_use_clustering
 * While it is on, scoring runs in the double path whatever precision is requested (diagonal pools only); loading a
 * model clears it. */
int akugpu_model_read_clustering(akugpu_ctx *ctx, const char *gcl_path);
int akugpu_model_set_clustering(akugpu_ctx *ctx, int n_clusters, const int32_t *gauss_index,
                                const int32_t *cluster_index, int64_t n_pairs);
int akugpu_model_set_clustering_min_evals(akugpu_ctx *ctx, double min_clusters, double min_gaussians);
int akugpu_model_use_clustering(akugpu_ctx *ctx, int on);
/* Model-level constrained MLLR with ONE global transform: the `model cmllr` entry of a speaker file with
 * `unitmode UNIT_NO` (ConstrainedMllr::set_parameters / load_transform, aku/ModelModules.cc:62-95,172-236;
 * phone_probs -S, aku/SpeakerConfig.cc:236-285).  W = [dim x (dim+1)] doubles, row-major, exactly the `w1` value:
 * column 0 is the bias b, columns 1..dim the matrix A.  Every Gaussian (and cluster centre) is then evaluated at
 * A f + b and its likelihood multiplied by the reference's factor |prod_i A(i,i)| (AdaptedGaussian, aku/ModelModules.hh:
 * 161-171; its "determinant" is the product of A's own diagonal, aku/LinearAlgebra.cc:74-86 -- reproduced as it is).
 * W = NULL removes the transform (ModelTransformer::reset_transforms); loading a model clears it.  Regression-class
 * transforms (unitmode UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN): akugpu_model_set_cmllr_units. */
int akugpu_model_set_cmllr(akugpu_ctx *ctx, const double *W);
/* The same module with regression classes: `unitmode UNIT_PHONE | UNIT_MIX | UNIT_GAUSSIAN` and entries `w<i> <units...>
 * <D x (D+1) numbers>` (ConstrainedMllr::set_parameters / load_transform, aku/ModelModules.cc:62-95,172-236).  units[t] =
 * the unit strings of transform t separated by blanks: centre-phone labels (Hmm::get_center_phone, aku/HmmSet.cc:22-40;
 * needs a model read from files), mixture indices, or Gaussian indices (RegClassTree::Unit*::get_gaussians,
 * aku/RegClassTree.cc:302-322,368-385,444-454); W = [n_transforms x D x (D+1)].  A Gaussian claimed by several
 * transforms takes the last one in the order of the reference's std::map (unit lists compared lexicographically); a
 * Gaussian no transform claims is left alone.  unitmode "UNIT_NO" = akugpu_model_set_cmllr; n_transforms = 0 removes the
 * transforms.  While regression-class transforms are loaded every request is served by the double path (each class has its
 * own feature row; diagonal pools, no Gaussian clustering). */
int akugpu_model_set_cmllr_units(akugpu_ctx *ctx, const char *unitmode, int n_transforms, const char *const *units, const double *W);
int akugpu_model_num_states(akugpu_ctx *ctx);   /* HmmSet::num_states()  */
int akugpu_model_dim(akugpu_ctx *ctx);          /* HmmSet::dim()         */
int akugpu_model_num_gaussians(akugpu_ctx *ctx);

/* ---- scoring -----------------------------------------------------------------------
 * Replaces, for F frames at once, HmmSet::precompute_likelihoods +
 * HmmSet::state_likelihood (aku/HmmSet.cc:485-501, aku/HmmSet.hh:309):
 *   precision F32: out = float  [F x S] natural-log state likelihoods
 *   precision F64: out = double [F x S] linear state likelihoods floored at 1e-50
 *                  (exactly what state_likelihood() returns). */
int akugpu_gmm_score(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames,
                     int precision, void *out);

/* The in-process decoder feed (decoder/decode-stream.cc:59-67,191-207 -> Toolbox::set_one_frame -> OneFrameAcoustics::set,
 * decoder/src/OneFrameAcoustics.cc:23-30): un-normalised log-probabilities
 *   out[f][s] = (float) safe_log(state_likelihood(s)),   safe_log(x) = log(max(x, tiny))
 * for F frames at once (F = 1 serves a per-frame loop: ~40 us per call).  decode-stream uses tiny = 1e-30; the
 * likelihood itself is floored at 1e-50 by HmmSet as always.  out is float [F x S] for either precision. */
int akugpu_gmm_logprobs(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, int precision, double tiny,
                        float *out);

/* Scores + the normalise/quantise loop of aku/phone_probs.cc:225-262.
 *   lnabytes 2: big-endian uint16 codes; 4: IEEE float32 little-endian
 *   normalize 0 == phone_probs --no-normalization
 *   out       [n_frames x S x lnabytes] bytes, no header; a DEVICE buffer must be 4-byte aligned (AKUGPU_E_ARG
 *             otherwise), host buffers may have any alignment.  */
int akugpu_gmm_lna(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames,
                   int precision, int lnabytes, int normalize, uint8_t *out);

/* Whole path for a batch of utterances: PCM -> features -> scores -> LNA records
 * (aku/phone_probs.cc:145-267 minus file I/O).  out receives the records of all
 * utterances back to back (no per-file header); frame_offsets as in akugpu_features.
 * out may be NULL (discard; for kernel-only timing) and checksum_out non-NULL to
 * receive a 64-bit sum of all output bytes computed on the device. */
int akugpu_phone_probs(akugpu_ctx *ctx, const int16_t *pcm, const int64_t *utt_offsets, int n_utts,
                       int precision, int lnabytes, int normalize,
                       uint8_t *out, int64_t *frame_offsets, uint64_t *checksum_out);

/* The same call that also returns one order-sensitive 64-bit checksum per utterance (utt_checksums [n_utts], host; may be
 * NULL), computed on the device from the records as they are written (see akugpu_checksum_begin for the definition).
 * An utterance's checksum does not depend on which other utterances shared the call or on which GPU ran it: the
 * utterance-sharded run over N GPUs (the reference's `-B N -I i` split, aku/phone_probs.cc:78-79,
 * 135-139) is checked against the single-GPU run with it. */
int akugpu_phone_probs_ex(akugpu_ctx *ctx, const int16_t *pcm, const int64_t *utt_offsets, int n_utts,
                          int precision, int lnabytes, int normalize,
                          uint8_t *out, int64_t *frame_offsets, uint64_t *checksum_out, uint64_t *utt_checksums);

/* ---- utterance-sharded runs over several GPUs (one process per GPU) ----------------------------------------------
 * No counterpart in the reference (its batch processes share nothing and write their own files); these are the
 * device-side pieces of "a final gather of LNA buffers" (BASELINE.json north_star).
 *
 * Checksum sink: per-utterance checksums of LNA records that are resident on this device -- produced here or received
 * from another rank.  With w_j = little-endian 32-bit word j of a frame's record (zero padded to whole words):
 *     row(f) = sum_j (2j+1) w_j          utt(u) = sum_i (2i+1) row(frame_offsets[u] + i)          (mod 2^64)
 *   begin   frame_offsets [n_utts+1] (host) numbers the frames of the stream the records belong to; rec_bytes = S * lnabytes
 *   update  records (DEVICE) = frames [first_frame, first_frame + n_frames) of that stream, in any order of calls,
 *           asynchronous on the context's stream
 *   end     waits and returns utt_checksums [n_utts] (host) */
int akugpu_checksum_begin(akugpu_ctx *ctx, const int64_t *frame_offsets, int n_utts, int64_t rec_bytes);
int akugpu_checksum_update(akugpu_ctx *ctx, const uint8_t *records, int64_t first_frame, int64_t n_frames);
int akugpu_checksum_end(akugpu_ctx *ctx, uint64_t *utt_checksums);
/* Device buffers that another process of the same node can map (CUDA IPC): the writer rank allocates its rotating
 * receive buffer with akugpu_shared_alloc and hands the 64-byte handle to the other ranks (any transport: the bench uses
 * a torch.distributed broadcast); they map it with akugpu_shared_open and pass the mapped address (+ their slot offset)
 * as the DEVICE `out` of akugpu_phone_probs / akugpu_gmm_lna: the LNA kernel then stores its records straight into
 * the writer's memory over NVLink -- epilogue and gather in one kernel, no send buffer.  akugpu_shared_release frees /
 * unmaps (also done by akugpu_destroy). */
int akugpu_shared_alloc(akugpu_ctx *ctx, size_t bytes, void **dev_ptr, unsigned char handle[64]);
int akugpu_shared_open(akugpu_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
int akugpu_shared_release(akugpu_ctx *ctx, void *dev_ptr);
/* cudaMemcpyAsync(dst, src, bytes) between device buffers (local or mapped peer memory) on the context's stream: the
 * copy-engine variant of the gather (records into a local slot, then over NVLink while the next sub-batch is scored). */
int akugpu_copy_async(akugpu_ctx *ctx, void *dst, const void *src, size_t bytes);

/* Writes the 5-byte LNA header (aku/phone_probs.cc:213-214): big-endian uint32
 * num_states, then lnabytes. */
int akugpu_lna_header(int n_states, int lnabytes, uint8_t out5[5]);

/* ---- tuning / introspection ---------------------------------------------------- */
/* Frames scored per chunk of the pipelined batch path.  0 (default) = one full wave of the
 * scorer (two waves of SM count x 128 frames = 37888 on B200). */
int akugpu_set_chunk_frames(akugpu_ctx *ctx, int64_t frames);
/* Kernel variant of the throughput-mode (F32) scorer:
 *   0 = default: tensor-core scorers (tcgen05 + TMEM + TMA, expanded form): the fp16 hi/lo-split kernel for diagonal
 *       pools with <= 64 components per state and dim <= 63, the bf16x3-split kernel for full-covariance pools (and as
 *       the fallback when a feature leaves the fp16 range); otherwise the FP32-pipe kernel (diagonal) or the double
 *       path (mixed pools);
 *       The expanded form is used only where it is well conditioned (1/2 sum_d (mu_d - c_d)^2 / var_d <= 200 around the
 *       minimax centre c: predicted error <= 8e-5 on a log-likelihood): states with a sharper component are scored by
 *       the direct-form FP32-pipe kernel in the same pass (diagonal pools; all states when they are half the model or
 *       more), ill-conditioned full-covariance pools by the double path;
 *   1 = FP32-pipe kernel with plain FFMA, 2 = FP32-pipe kernel with packed FFMA2,
 *   3 = the tensor-core scorers as in 0 but without the conditioning check,
 *   4 = force the bf16x3-split tensor-core kernel. */
int akugpu_set_scorer_variant(akugpu_ctx *ctx, int variant);
/* Conditioning of the expanded (GEMM) form for the loaded model: max over Gaussians of 1/2 sum_d (mu_d - c_d)^2 / var_d
 * around the centre the library chose (0 when no tensor-core image was considered; -1 without a model). */
double akugpu_model_expanded_form_q(akugpu_ctx *ctx);
/* Which kernel serves throughput-mode (F32) requests for the loaded model: 0 = the double path (mixed / ill-conditioned
 * full-covariance pools, Gaussian clustering), 1 = FP32-pipe direct form, 2 = bf16x3 tensor-core, 3 = fp16x2 tensor-core
 * with resident A', 4 = fp16x2 tensor-core streaming A', 5 = fp16x2 tensor-core for the well-conditioned states + FP32-pipe
 * kernel for the others. */
int akugpu_scorer_in_use(akugpu_ctx *ctx);
/* Streaming-regime scorer.  Calls of akugpu_gmm_score (precision F32) / akugpu_gmm_logprobs with at most 16 frames (8 for models of more
 * than 800 component tiles) --
 * the decoder's per-frame loop, decoder/decode-stream.cc:191-207 -- are served by ONE launch that spreads the component
 * tiles of the whole model over all SMs (gmm_stream_kernel); host features travel in the kernel's parameter block,
 * results and the completion flag are written straight into pinned, mapped host memory.  Diagonal pools served by the
 * fp16x2 tensor-core image only (akugpu_scorer_in_use() == 3, no model-level CMLLR, no Gaussian clustering); everything
 * else, and calls whose features leave the fp16 range, take the general path.  akugpu_set_streaming(ctx, 0) switches the
 * fast path off (tests compare the two). */
int akugpu_set_streaming(akugpu_ctx *ctx, int enable);
/* What the streaming regime is measured against, for the loaded model: out[0] = bytes of the parameter image swept per
 * call; out[1] / out[2] = seconds per gmm_stream_kernel launch (one frame, CUDA events around the kernel) with the image
 * L2-resident / after an L2 flush (a 512 MB fill before every launch); out[3] / out[4] = seconds of a plain uint4 read
 * sweep of the same image by all SMs, L2-resident / after a flush: the L2 and HBM read rooflines of this very buffer;
 * out[5] = SM clock (MHz) right after the isolated launches (an idle GPU clocks down between them); out[6] = seconds per
 * launch in a train of 200 back-to-back launches (GPU busy, image L2-resident); out[7] = SM clock (MHz) after the train. */
int akugpu_stream_probe(akugpu_ctx *ctx, double out[8]);
/* Resident scorer for a stream decoder's session (decoder/decode-stream.cc:150-207: one Toolbox + one acoustic model
 * for the life of the stream, one scoring call per frame).  akugpu_stream_open starts ONE kernel that stays on the
 * device: each of its CTAs keeps its share of the parameter image in shared memory (the config-2 model, 5000 x 16 =
 * 30.7 MB, fits the 148 SMs whole; of a larger model the tiles that do not fit are streamed from L2 per call).  While
 * it runs, calls of akugpu_gmm_score (F32) / akugpu_gmm_logprobs with host buffers and at most 16 frames are MESSAGES to
 * that kernel through pinned, mapped memory -- no launch and no CUDA call on the path of a frame -- and return the
 * bits the launch-per-call scorer returns.  The kernel ends when it is told to (akugpu_stream_close, and ANY other
 * entry point of this context sends that message first: nothing of the library runs beside it) or by itself after
 * idle_ms without a call (<= 0: 100 ms); the next eligible call starts it again.  Other CUDA work of the process on
 * the same device waits for that moment too: open a session on a GPU the decoder owns (one session per device and process:
 * a second context's akugpu_stream_open fails with AKUGPU_E_STATE until the first closes its session).  Models the streaming scorer
 * does not serve (akugpu_set_streaming) are refused with AKUGPU_E_STATE.
 * akugpu_stream_stats: out[0] = session requested, out[1] = kernel believed to be running, out[2] = kernel launches,
 * out[3] = calls served by it; out[4] / out[5] = the last call as the last CTA to finish saw it, in nanoseconds of the
 * device's timer: command seen -> A' built, command seen -> every CTA done (host-side latency minus out[5] is PCIe and
 * polling); out[6] / out[7] = command seen -> that CTA's results stored / fenced. */
int akugpu_stream_open(akugpu_ctx *ctx, double idle_ms);
/* The per-frame call of an open session without a copy and without pointer queries: feats = n_frames (1..16) rows of
 * HOST float features, *rows = [n_frames][n_states] floats of (float) log(max(likelihood, tiny)) (tiny <= 0: the plain
 * log-likelihoods) in pinned memory of the context, valid until its next call -- what decoder/decode-stream.cc:191-207
 * computes into its vector before Toolbox::set_one_frame.  AKUGPU_E_STATE without akugpu_stream_open. */
int akugpu_stream_logprobs(akugpu_ctx *ctx, const float *feats, int n_frames, double tiny, const float **rows);
int akugpu_stream_close(akugpu_ctx *ctx);
/* Measurement aid: n_calls per-frame-loop calls on the same n_frames rows of host float features, each timed on the host
 * (steady clock) inside the library, i.e. without the caller's own call overhead: out_us = {mean, median, 99th
 * percentile, maximum} microseconds per call.  Inside an open session a call is what akugpu_stream_logprobs does;
 * otherwise what akugpu_gmm_logprobs (tiny > 0) / akugpu_gmm_score does with host buffers. */
int akugpu_stream_latency(akugpu_ctx *ctx, const float *feats, int n_frames, double tiny, int n_calls, double out_us[4]);
int akugpu_stream_stats(akugpu_ctx *ctx, int64_t out[8]);

/* Micro-benchmarks of the issue pipes the scorer depends on (lane-ops per second):
 * out[0] FFMA, out[1] FFMA2 (counted as 2 lane-ops), out[2] DFMA, out[3] MUFU.EX2 with constant
 * operands; out[4] FFMA and out[5] FFMA2 with three distinct register operands; out[6] / out[7]
 * the scorer's own register tile (FFMA 8x8 / FFMA2 8x4) without memory traffic. */
int akugpu_pipe_rates(akugpu_ctx *ctx, double out[8]);

#ifdef __cplusplus
}
#endif
#endif /* AKUGPU_H */
