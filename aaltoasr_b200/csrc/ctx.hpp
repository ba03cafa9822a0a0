// ctx.hpp -- context object behind the C ABI (include/akugpu.h).
// Owns every device allocation; nothing allocated here crosses the boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include <map>
#include <chrono>
#include <memory>
#include "../../include/akugpu.h"

namespace akugpu {

struct Error {
  int code;
  std::string msg;
  Error(int c, const std::string &m) : code(c), msg(m) {}
};

inline std::string fmt(const char *f, ...) {
  char buf[1024];
  va_list ap; va_start(ap, f); vsnprintf(buf, sizeof buf, f, ap); va_end(ap);
  return std::string(buf);
}

#define AKU_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      throw ::akugpu::Error(AKUGPU_E_CUDA, ::akugpu::fmt("%s:%d: %s: %s", __FILE__, __LINE__, #call, \
                                                       cudaGetErrorString(e_)));       \
  } while (0)

// Growable device buffer (never shrinks).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    // a buffer that has to grow again gets 1/8 of headroom: batches of slightly different sizes (utterances of
    // jittered length) must not pay a cudaFree + cudaMalloc -- hundreds of ms for GB-sized buffers -- every few calls
    const size_t want = p ? bytes + bytes / 8 : bytes;
    if (p) AKU_CUDA(cudaFree(p));
    p = nullptr; cap = 0;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; AKU_CUDA(cudaMalloc(&p, bytes)); cap = bytes; return; }
    cap = want;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T *as() const { return (T *)p; }
  ~DevBuf() { release(); }
};
struct PinnedBuf {
  void *p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) AKU_CUDA(cudaFreeHost(p));
    p = nullptr; cap = 0;
    AKU_CUDA(cudaMallocHost(&p, bytes));
    cap = bytes;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T *as() const { return (T *)p; }
  ~PinnedBuf() { release(); }
};

inline bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------------------------
// Acoustic model, host view (doubles, as read) and packed device images.
struct HostModel {
  int S = 0, G = 0, D = 0;
  std::vector<int32_t> mix_off;     // [S+1]
  std::vector<int32_t> mix_gauss;   // [K]
  std::vector<double> mix_w;        // [K] normalised
  std::vector<double> mean, cov;    // [G*D]  (cov: diagonal; unused rows for full Gaussians)
  // full-covariance Gaussians (aku/Distributions.cc:1467-1488): full_index[g] = row in full_cov, -1 = diagonal
  std::vector<int32_t> full_index;  // [G] (empty = all diagonal)
  std::vector<double> full_cov;     // [n_full * D * D] row-major, as read
  int n_full = 0;
  // Gaussian clustering (PDFPool::read_clustering aku/Distributions.cc:3115-3170, precompute_likelihoods :2685-2722)
  int n_clusters = 0;
  std::vector<std::vector<int32_t>> cluster_gauss;   // members as listed (a Gaussian may be listed twice, see model.cu)
  std::vector<int32_t> gauss_cluster;                // [G] cluster of each Gaussian, -1 = not listed
  std::vector<double> c_mean, c_cov;                 // [C * D] moment-matched diagonal centres (Gaussian::merge :854-897)
  bool use_clustering = false;
  int eval_min_clusters = 1, eval_min_gaussians = 1; // PDFPool defaults (:2572-2574)
  void clear_clustering() {
    n_clusters = 0; cluster_gauss.clear(); gauss_cluster.clear(); c_mean.clear(); c_cov.clear();
    use_clustering = false; eval_min_clusters = eval_min_gaussians = 1;
  }
  // Global model-level constrained MLLR (ConstrainedMllr with unitmode UNIT_NO, aku/ModelModules.cc:172-236): every
  // Gaussian is evaluated at A f + b and its likelihood multiplied by the module's "determinant" (see api.cu).  The
  // factor is folded into the mixture weights (mix_w = mix_w_base * factor) so that every scorer image carries it.
  bool cmllr_on = false;
  std::vector<double> cmllr_W;      // [D x (D+1)] row-major as given: column 0 = b, columns 1..D = A
  std::vector<double> mix_w_base;   // the normalised weights while cmllr_on
  // Regression-class transforms (unitmode UNIT_PHONE / UNIT_MIX / UNIT_GAUSSIAN, aku/ModelModules.cc:62-95,172-236):
  // Gaussian g with g_tr[g] = t >= 0 is evaluated at A_t f + b_t and multiplied by transform t's factor.  cmllr_on is set
  // too; the throughput scorers do not serve such a model (one feature row per class), the double path does.
  int n_tr = 0;
  std::vector<int32_t> g_tr;        // [G]
  std::vector<double> tr_Ab;        // [n_tr][D*D + D]: A row-major, then b
  void clear_cmllr() { cmllr_on = false; cmllr_W.clear(); mix_w_base.clear(); n_tr = 0; g_tr.clear(); tr_Ab.clear(); }
  // phones of the .ph file (label, states) -- what UNIT_PHONE transforms are resolved with; empty for direct loads
  std::vector<std::string> ph_label;
  std::vector<std::vector<int32_t>> ph_states;
};

// fp32 scorer image: tiles of 8 slots x 16 components, see gmm_kernels.cu.
struct PackedF32 {
  bool packed_ffma2 = true;  // FFMA2 (default) or plain FFMA instructions
  int DP = 0;                // dim pairs
  int n_tiles = 0;
  size_t tile_floats = 0;
  DevBuf params;             // n_tiles * tile_floats floats
  DevBuf center;             // float [2*DP] feature centre
  DevBuf center64;           // double [2*DP]
  std::vector<char> clean;   // [n_tiles] 1 = every queue starts a new state at this tile
  std::map<int, std::pair<int, std::shared_ptr<DevBuf>>> ranges;   // ysplit -> (count, device int[count+1])
};
// fp64 scorer image: plain arrays in the reference's own order.
struct PackedF64 {
  DevBuf mean, prec;     // double [G*D]
  DevBuf cst;            // double [G]   m_constant
  DevBuf mix_off, mix_gauss;  // int32
  DevBuf mix_w;          // double [K]
  // full-covariance part (exponential form, aku/Distributions.cc:1437-1446,1530-1547)
  int n_full = 0, L = 0;       // L = D(D+3)/2
  DevBuf theta;                // double [L][n_full]   exponential parameters, one column per full Gaussian
  DevBuf full_norm, full_cst;  // double [n_full]      m_exponential_normalizer, m_constant
  DevBuf full_gauss;           // int32  [n_full]      pool index of each full Gaussian
  DevBuf diag_gauss;           // int32  [G - n_full]  pool indices of the diagonal Gaussians
  // Gaussian clustering: centres as a pool of C one-component "states", cluster of each Gaussian, member counts
  DevBuf c_mean, c_prec, c_cst, c_mix_off, c_mix_gauss, c_mix_w, g2c, c_size;
  DevBuf g_tr;                 // int32 [G] transform of each Gaussian (regression-class CMLLR), -1 = none
};

// tensor-core scorer image (gmm_tc.cu): bf16x3-split expanded parameters, slot-ordered rows
struct PackedTC {
  bool ready = false, full = false;
  double q_max = 0;            // conditioning of the expanded form for this model (gmm_tc.cu: tc_expanded_params)
  int L = 0, Lm = 0, Kp = 0, n_tiles = 0;
  DevBuf B, bias, meta, center;
  std::vector<char> clean;
  std::map<int, std::pair<int, std::shared_ptr<DevBuf>>> ranges;
};

// fp16x2 tensor-core scorer image (gmm_tc16.cu): diagonal pools, hi/lo-split expanded parameters with per-term
// power-of-two scaling, the component constant folded into two extra K terms
struct PackedTC16 {
  bool ready = false, full = false, stream = false;   // stream: A' too wide for shared memory (full covariance), streamed like B'
  int D = 0, L = 0, NCH = 0, KB = 0, n_tiles = 0;   // L = terms incl. the 2 constant ones, NCH = K16 chunks per half, KB = 64-wide k-blocks
  int half = 0, Kp = 0;                              // columns of Bh (= offset of Bl) and of a whole row
  double q_max = 0;                                  // conditioning of the expanded form for this model
  bool hybrid = false;                               // states in bad_state are left to the FP32-pipe kernel (packed into p32)
  std::vector<char> bad_state;                       // [S] 1 = a component of the state is too ill-conditioned for the expanded form
  int n_bad = 0;
  DevBuf B, meta, center, escale, flag;
  std::vector<char> clean;
  std::map<int, std::pair<int, std::shared_ptr<DevBuf>>> ranges;
  std::vector<double> h_center;                      // host copy of the feature centre (the streaming scorer centres on the host)
  alignas(64) unsigned char map_b[128];              // CUtensorMap of B', encoded once per model (streaming scorer)
  alignas(64) unsigned char map_b64[128];            // the same rows as 32-column boxes with SWIZZLE_64B (resident scorer)
  bool map_ready = false;
};

// streaming-regime scorer (gmm_stream.cu): pinned, mapped host block {flags | centred features | results}, CTA counter
constexpr int STREAM_MAX_FRAMES = 16;     // measured (tests/test_gpu_stream.py): beyond 16 frames (8 at 1250 tiles) the general path is as fast
struct StreamState {
  void *host = nullptr, *dev_view = nullptr;
  size_t bytes = 0;
  DevBuf cnt, relay;          // relay: {command word (16 B) | pad | features of the call} the polling CTA hands to the others
  unsigned int seq = 0;
  // resident scorer (gmm_resident.cu): a kernel that stays on the device between calls, parameter image in shared memory
  bool session_want = false, session_live = false;
  double session_idle_ms = 100.0;
  cudaStream_t session_stream = nullptr;
  void *pkt_host = nullptr, *pkt_dev = nullptr;      // one 512-byte command packet per CTA, pinned + mapped
  int session_grid = 0;
  const void *known_host[2] = {nullptr, nullptr};    // buffers of recent session calls already classified as host memory
  std::vector<int> session_order;                    // CTA indices, most component tiles first
  int64_t session_launches = 0, session_calls = 0;
  std::chrono::steady_clock::time_point session_last;
};

// ---------------------------------------------------------------------------------
// Front-end module graph (parsed from the reference's feature configuration).
enum ModType { M_AUDIOFILE, M_FFT, M_MEL, M_POWER, M_MEL_POWER, M_DCT, M_DELTA, M_MERGE, M_CONCAT,
               M_NORMALIZATION, M_LIN_TRANSFORM, M_MEAN_SUBTRACTOR, M_PRE, M_VTLN, M_SR_NORM, M_QUANTEQ };

struct Module {
  std::string name;
  ModType type;
  std::vector<int> src;     // indices of source modules
  int dim = 0;
  int left = 0, right = 0;  // own context (frames)
  // audiofile
  int sample_rate = 0, window_width = 0, copy_borders = 1;
  int legacy_file = 0;      // pre: 1-byte dimension header (PreModule::set_file, aku/FeatureModules.cc:608-615)
  float frame_rate = 125.f, window_advance = 0.f, emph = 0.97f;
  // fft
  int magnitude = 1, log = 0;
  // mel
  int root = 0;
  std::vector<float> bin_edges;
  // dct
  int zeroth = 0;
  // delta
  int width = 2;
  float norm = 10.f;
  // normalization / lin_transform
  std::vector<float> v_mean, v_scale, matrix, bias;
  bool matrix_defined = false, bias_defined = false;
  // mean_subtractor: left/right hold the +1-extended offsets, width = left+right-1
  int ms_width = 0;
  // vtln (VtlnModule, aku/FeatureModules.cc:1505-1934): warp of the spectrum bins, sinc / linear interpolation
  int use_pwlin = 0, use_slapt = 0, sinc_rad = 8, lanczos = 1, all_pass = 0;
  float pwlin_turn_point = 0.8f, warp_factor = 1.0f;
  std::vector<float> slapt_params, vtln_bins;
  // sr_norm (SRNormModule :1936-2069): Lanczos resampling across the stacked frames of a concat module
  int in_frames = 0, out_frames = 0, frame_dim = 0, lanczos_order = 4;
  float speech_rate = 1.0f;
  // quanteq (QuantEqModule :2072-2148): per-channel power-law equalisation, parameters set at run time only
  std::vector<float> q_alpha, q_gamma, q_max;
  // device copies of parameters
  std::shared_ptr<DevBuf> d_a, d_b, d_c;
  std::shared_ptr<DevBuf> d_d;   // double copy of a float table (mel triangle weights, dct basis): the warp-per-frame front-end kernel
                                 // reads the widened values instead of converting them term by term
  // extended frame range this module must be evaluated on for an utterance:
  // [-ext_left, n_frames-1+ext_right]
  int ext_left = 0, ext_right = 0;
};

struct Frontend {
  bool configured = false;
  std::vector<Module> mods;
  int last = -1;
  // fused static path: audiofile -> fft -> {mel->dct, power, mel_power...}
  std::shared_ptr<DevBuf> d_window;   // float [W] hamming
  std::shared_ptr<DevBuf> d_twiddle;  // float2 [W/2] or DFT table
};

struct StageTimer {
  bool enabled = false;
  double ms[3] = {0, 0, 0};
  int64_t launches[3] = {0, 0, 0};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
  std::vector<cudaEvent_t> pool;
};

}  // namespace akugpu

struct akugpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;       // compute stream
  bool own_stream = true;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  // second compute stream: the LNA epilogue of chunk k runs here next to the scorer of chunk k+1 (api.cu, score_to_lna_impl)
  cudaStream_t lna_stream = nullptr;
  cudaEvent_t ev_sc[2] = {nullptr, nullptr}, ev_ln[2] = {nullptr, nullptr};
  bool overlap_lna = false;      // measured (profiles/r02_overlap_experiment.txt): no gain, the step is power-capped; AKUGPU_OVERLAP=1 switches it on
  std::string err;
  int64_t launches = 0;
  int sm_count = 148;
  int64_t chunk_frames = 0;   // 0 = auto: one full wave of the scorer per chunk
  int scorer_variant = 0;
  bool tc16_suspended = false;   // set while a call is redone with the bf16x3 kernel after an fp16 range overflow
  bool streaming_enabled = true; // small calls (<= STREAM_MAX_FRAMES frames) of the decoder feed take gmm_stream_kernel
  akugpu::StreamState stream_state;
  std::vector<float> stream_fallback;   // akugpu_stream_logprobs: rows of a call the resident scorer could not serve

  akugpu::HostModel hm;
  bool have_model = false;
  akugpu::PackedF32 p32;
  akugpu::PackedF64 p64;
  akugpu::PackedTC ptc;
  akugpu::PackedTC16 ptc16;
  bool have_p64 = false;

  akugpu::Frontend fe;
  akugpu::StageTimer timer;

  // scratch
  akugpu::DevBuf d_feats, d_sll, d_lna[2], d_pcm, d_tmp, d_chk, d_norm, d_clik, d_csel;
  int64_t adapt_stride = 0;            // elements between the adapted feature arrays of two transforms (F * D of the current call)
  akugpu::DevBuf d_sll2, d_norm2;      // second score / normaliser buffers of the overlapped pipeline
  akugpu::DevBuf d_cmllr, d_adapt;     // double [n][D*D + D] (A row-major, then b) per transform; adapted features of the current call
                                       // (global transform: [F][D]; regression classes: [n_tr][F][D])
  akugpu::DevBuf d_fe[8];
  std::vector<std::shared_ptr<akugpu::DevBuf>> fe_bufs;   // per-module output matrices (grow-only)
  akugpu::PinnedBuf h_in[2], h_out[2];
  // per-utterance checksum sink (multigpu.cu): frame-offset table and 64-bit accumulators on the device
  akugpu::DevBuf d_chk_fo, d_chk_acc;
  bool chk_open = false;
  int chk_n_utts = 0;
  int64_t chk_rec = 0, chk_frames = 0, chk_first = 0;
  std::vector<void *> shared_own, shared_peer;   // akugpu_shared_alloc / akugpu_shared_open buffers still held
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr};
};
