// lna_kernels.cu -- normalise + quantise epilogue (K6): state likelihoods -> LNA records.
//
// Replaces the per-frame loop of aku/phone_probs.cc:225-262 (twin:
// aku/PhoneProbsToolbox.cc:80-131).  The reference's arithmetic is a float/double
// hybrid that has to be reproduced, not improved (hard part 1):
//   obs[s]  = (float) max(lik_s, 1e-50)          float cast BEFORE the sum: values
//                                                below 2^-150 become 0, values in the
//                                                fp32 denormal range are rounded to a
//                                                multiple of 2^-149
//   Z       = sum_s (double) obs[s] ;  Z == 0 or --no-normalization -> Z = 1
//   lp[s]   = (float) safe_log(obs[s] / Z)       safe_log floors at log(1e-50)
//   4 bytes : raw IEEE float, little-endian
//   2 bytes : lp < -36.008 -> 0xFFFF, else big-endian (int)(-1820.0*lp + .5)
//
// lna_f32: input = natural-log state likelihoods (fp32, state-major [S][ldF]) from
//          gmm_diag_f32.  Everything above is done in the log domain; the denormal
//          rounding is emulated as rint(exp(ll + 149 ln2)); the S-way normalisation
//          excludes the maximum term (lp_max = -log1p(sum of the others)) so the
//          dominant state does not lose its value to cancellation.
// lna_f64: input = linear double likelihoods from gmm_diag_f64; literally the
//          reference's sequence of operations, states summed in index order.
#include "ctx.hpp"
#include "kernels.hpp"
#include "lna_common.cuh"
#include <cuda.h>
#include "tc_common.cuh"

namespace akugpu {

template <int B>
__device__ __forceinline__ void lna_store(uint8_t *dst, float lp);
template <>
__device__ __forceinline__ void lna_store<4>(uint8_t *dst, float lp)
{
  uint32_t u = __float_as_uint(lp);
  dst[0] = u & 255; dst[1] = (u >> 8) & 255; dst[2] = (u >> 16) & 255; dst[3] = (u >> 24) & 255;
}
__device__ __forceinline__ uint32_t lna_code16(float lp)
{
  if ((double)lp < -36.008) return 0xFFFFu;
  int temp = (int)(-1820.0 * (double)lp + .5);
  return (uint32_t)(((temp >> 8) & 255) << 8 | (temp & 255));
}
template <>
__device__ __forceinline__ void lna_store<2>(uint8_t *dst, float lp)
{
  uint32_t c = lna_code16(lp);
  dst[0] = (c >> 8) & 255; dst[1] = c & 255;
}

// CTA = 32 frames (lanes) x 8 warps; warp w owns states r*64 + w*8 .. +7 of every round r.
template <int B>
__global__ void __launch_bounds__(256)
lna_f32(const float *__restrict__ sll, int64_t ldF, int S, int64_t nf, int normalize, const float2 *__restrict__ norm,
        uint8_t *__restrict__ out)
{
  typedef typename std::conditional<B == 2, uint16_t, uint32_t>::type elem_t;
  __shared__ float sh_M[8][32];
  __shared__ double sh_R[8][32];
  __shared__ elem_t tile[32][64 + (B == 2 ? 2 : 1)];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const int64_t f = f0 + lane;
  const bool fvalid = f < nf;
  const float *col = sll + (fvalid ? f : 0);

  float Mx = -INFINITY;
  double lognorm = 0.0;
  if (normalize && norm) {            // normaliser already computed by the scorer's epilogue (gmm_tc.cu)
    const float2 nm = norm[fvalid ? f : 0];
    Mx = nm.x;
    lognorm = (double)nm.y;
  } else if (normalize) {
    double R = 0.0;
    for (int sb = w * 8; sb < S; sb += 64) {
      // 8 independent loads in flight, then one rescale per batch
      float L[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) L[j] = (sb + j < S) ? __ldcs(col + (int64_t)(sb + j) * ldF) : -INFINITY;
      float bm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) { L[j] = log_of_float_cast(L[j]); bm = fmaxf(bm, L[j]); }
      if (bm == -INFINITY) continue;
      // the running maximum is excluded from R: R = sum over all other terms of exp(L - Mx)
      bool excl = false;
      if (bm > Mx) {
        R = (Mx == -INFINITY) ? 0.0 : (R + 1.0) * (double)__expf(Mx - bm);
        Mx = bm;
        excl = true;          // the first element equal to the new maximum is the excluded one
      }
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (excl && L[j] == Mx) { excl = false; continue; }
        if (L[j] != -INFINITY) acc += __expf(L[j] - Mx);
      }
      R += (double)acc;
    }
    sh_M[w][lane] = Mx;
    sh_R[w][lane] = R;
    __syncthreads();
    float gM = sh_M[0][lane];
    int wstar = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k)
      if (sh_M[k][lane] > gM) { gM = sh_M[k][lane]; wstar = k; }
    double Rt = 0.0;
    if (gM != -INFINITY) {
      Rt = sh_R[wstar][lane];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k != wstar && sh_M[k][lane] != -INFINITY) Rt += (1.0 + sh_R[k][lane]) * exp((double)(sh_M[k][lane] - gM));
    }
    Mx = gM;
    lognorm = (double)(float)log1p(Rt);     // the normaliser is a float everywhere (scorer epilogue, lna_f32_norm)
  }

  for (int s_round = 0; s_round < S; s_round += 64) {
    float Lv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int s = s_round + w * 8 + j;
      Lv[j] = (s < S) ? __ldcs(col + (int64_t)s * ldF) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int s = s_round + w * 8 + j;
      float lp = LP_FLOOR;
      if (s < S) {
        float L = log_of_float_cast(Lv[j]);
        if (normalize) {
          if (Mx != -INFINITY && L != -INFINITY) lp = (float)((double)(L - Mx) - lognorm);
        } else {
          lp = L;
        }
        if (!(lp >= LP_FLOOR)) lp = LP_FLOOR;
      }
      if (B == 2) {
        uint32_t c = lna_code16(lp);
        tile[lane][w * 8 + j] = (elem_t)(((c >> 8) & 255) | ((c & 255) << 8));   // big-endian in memory
      } else {
        tile[lane][w * 8 + j] = (elem_t)__float_as_uint(lp);
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int fr = w * 4 + r;
      int64_t gf = f0 + fr;
      if (gf < nf) {
        elem_t *dst = reinterpret_cast<elem_t *>(out) + gf * S + s_round;
        if (s_round + lane < S) dst[lane] = tile[fr][lane];
        if (s_round + 32 + lane < S) dst[32 + lane] = tile[fr][32 + lane];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// lna_f32_norm: the first pass of lna_f32 on its own -- per frame the maximum of the float-cast state likelihoods and
// log1p of the sum of all the others relative to it -- for scorers that do not produce it in their epilogue (FP32-pipe
// kernel, launches whose component tiles are split over several CTAs).  Same layout of work as lna_f32; the result
// feeds lna_f32_rows, which then reads the scores once more.
__global__ void __launch_bounds__(256)
lna_f32_norm(const float *__restrict__ sll, int64_t ldF, int S, int64_t nf, float2 *__restrict__ norm)
{
  __shared__ float sh_M[8][32];
  __shared__ double sh_R[8][32];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t f = (int64_t)blockIdx.x * 32 + lane;
  const bool fvalid = f < nf;
  const float *col = sll + (fvalid ? f : 0);
  float Mx = -INFINITY;
  double R = 0.0;
  for (int sb = w * 8; sb < S; sb += 64) {
    float L[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) L[j] = (sb + j < S) ? __ldg(col + (int64_t)(sb + j) * ldF) : -INFINITY;
    float bm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) { L[j] = log_of_float_cast(L[j]); bm = fmaxf(bm, L[j]); }
    if (bm == -INFINITY) continue;
    bool excl = false;
    if (bm > Mx) {
      R = (Mx == -INFINITY) ? 0.0 : (R + 1.0) * (double)__expf(Mx - bm);
      Mx = bm;
      excl = true;          // the first element equal to the new maximum is the excluded one
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (excl && L[j] == Mx) { excl = false; continue; }
      if (L[j] != -INFINITY) acc += __expf(L[j] - Mx);
    }
    R += (double)acc;
  }
  sh_M[w][lane] = Mx;
  sh_R[w][lane] = R;
  __syncthreads();
  if (w != 0 || !fvalid) return;
  float gM = sh_M[0][lane];
  int wstar = 0;
#pragma unroll
  for (int k = 1; k < 8; ++k)
    if (sh_M[k][lane] > gM) { gM = sh_M[k][lane]; wstar = k; }
  double Rt = 0.0;
  if (gM != -INFINITY) {
    Rt = sh_R[wstar][lane];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k != wstar && sh_M[k][lane] != -INFINITY) Rt += (1.0 + sh_R[k][lane]) * exp((double)(sh_M[k][lane] - gM));
  }
  norm[f] = make_float2(gM, (float)log1p(Rt));
}

// ------------------------------------------------------------------------------------------------
// lna_f32_rows: same results as lna_f32, laid out for HBM bandwidth.  CTA = 32 frames (lanes) x 8 warps; a round
// covers 256 states (2-byte codes) or 128 states (floats): every lane reads its frame's column (a warp reads one
// 128 B line per state, 16 loads in flight per thread), converts in fp32 only, and packs the records into a
// double-buffered shared tile [frame][128 words]; after ONE barrier per round each warp streams four frame rows out
// as 128 B line stores.  Used when the rows are 4-byte aligned (S * B % 4 == 0).
//
// All arithmetic is fp32 and bit-identical to the double expressions of lna_f32:
//   * lp = (float)((double)(L - Mx) - (double)lognorm_f) is the correctly rounded difference of two floats, i.e. the
//     fp32 subtraction (the scorer's normaliser is a float);
//   * (double)lp < -36.008 holds exactly when lp < 0xc2100831 (the smallest float above -36.008);
//   * (int)(-1820.0 * (double)lp + .5): the double expression is exact (24 x 11 bits), so it is the truncation
//     towards zero of the exact value -- for lp <= 0 one round-down fma (lna_code16_f32_neg); for any sign the
//     nearest integer of fmaf(-1820, lp, .5), moved one step towards zero when the sign of fmaf(-1820, lp, .5 - k)
//     (an fma's sign is the exact sign) says k overshot.
__device__ __noinline__ uint32_t lna_code16_f32_wide(float lp)     // absurd magnitudes / NaN: the double expression itself
{
  const int temp = (int)(-1820.0 * (double)lp + .5);
  return (uint32_t)temp & 0xFFFFu;
}
// any sign (un-normalised log-likelihoods can be positive: negative codes wrap like the reference's two's complement)
__device__ __forceinline__ uint32_t lna_code16_f32(float lp)
{
  if (lp < __uint_as_float(0xc2100831u)) return 0xFFFFu;
  const float t = fmaf(-1820.f, lp, 0.5f);
  if (!(fabsf(t) < 4194304.f)) return lna_code16_f32_wide(lp);
  float kf = (t + 12582912.f) - 12582912.f;                   // nearest integer (ties to even), |kf| < 2^22
  const float d = fmaf(-1820.f, lp, 0.5f - kf);               // sign of (exact value - kf); 0.5 - kf is exact
  if (kf > 0.f && d < 0.f) kf -= 1.f;                         // (int) truncates towards zero
  else if (kf < 0.f && d > 0.f) kf += 1.f;
  return (__float_as_uint(kf + 12582912.f) - 0x4B400000u) & 0xFFFFu;
}
// lp <= 0 (normalised records): floor(-1820 lp + .5) in ONE round-down fma: the exact value + 2^22 lands in
// [2^22, 2^23) where the fp32 grid is 0.5, rounding down to that grid and dropping the half bit is the floor.
__device__ __forceinline__ uint32_t lna_code16_f32_neg(float lp)
{
  const uint32_t m = __float_as_uint(__fmaf_rd(-1820.f, lp, 4194304.5f));
  return (lp < __uint_as_float(0xc2100831u)) ? 0xFFFFu : ((m >> 1) & 0xFFFFu);
}
__device__ __noinline__ float log_of_float_cast_rare(float v) { return log_of_float_cast(v); }

template <int B, bool NORM>
__global__ void __launch_bounds__(256)
lna_f32_rows(const float *__restrict__ sll, int64_t ldF, int S, int64_t nf, const float2 *__restrict__ norm, uint8_t *__restrict__ out)
{
  constexpr int SPR = (B == 2) ? 256 : 128;      // states per round
  constexpr int SPW = SPR / 8;                   // states per warp and round
  constexpr int NB = 16;                         // loads in flight per thread
  constexpr int TW = 128;                        // 32-bit words per frame row of the tile
  __shared__ uint32_t tile[2][32][TW + 1];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const int64_t f = f0 + lane;
  const bool fvalid = f < nf;
  float Mx = 0.f, lognorm = 0.f;
  if (NORM) {
    const float2 nm = norm[fvalid ? f : 0];
    // every state flushed to zero (Mx = -inf): Z = 1 and every record is the floor; +inf makes L - Mx = -inf
    Mx = (nm.x == -INFINITY) ? INFINITY : nm.x;
    lognorm = nm.y;
  }
  const int rounds = (S + SPR - 1) / SPR;
  const float *src = sll + (fvalid ? f : 0) + (int64_t)(w * SPW) * ldF;
  for (int r = 0; r < rounds; r++, src += (int64_t)SPR * ldF) {
    uint32_t(*tl)[TW + 1] = tile[r & 1];
    const int sw = r * SPR + w * SPW;
#pragma unroll
    for (int b = 0; b < SPW; b += NB) {
      float lp[NB];
      const float *p = src + (int64_t)b * ldF;
#pragma unroll
      for (int j = 0; j < NB; ++j, p += ldF) lp[j] = (sw + b + j < S) ? __ldcs(p) : 0.f;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float L = lp[j];
        if (L < LN_2M126) L = log_of_float_cast_rare(L);       // fp32-denormal / zero range of the float cast: rare
        const float v = NORM ? (L - Mx) - lognorm : L;
        lp[j] = fmaxf(v, LP_FLOOR);                            // also catches -inf and NaN
      }
      if (B == 2) {
#pragma unroll
        for (int j = 0; j < NB; j += 2) {
          const uint32_t c0 = NORM ? lna_code16_f32_neg(lp[j]) : lna_code16_f32(lp[j]);
          const uint32_t c1 = NORM ? lna_code16_f32_neg(lp[j + 1]) : lna_code16_f32(lp[j + 1]);
          tl[lane][(w * SPW + b + j) >> 1] = __byte_perm(c0, c1, 0x4501);   // big-endian codes, two per word
        }
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) tl[lane][w * SPW + b + j] = __float_as_uint(lp[j]);
      }
    }
    __syncthreads();
    const int nwords = min(SPR, S - r * SPR) * B / 4;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int fr = w * 4 + rr;
      const int64_t gf = f0 + fr;
      if (gf < nf) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(out + ((size_t)gf * S + (size_t)r * SPR) * B);
#pragma unroll
        for (int k = 0; k < TW / 32; ++k)
          if (lane + 32 * k < nwords) __stcs(dst + lane + 32 * k, tl[fr][lane + 32 * k]);
      }
    }
    // no second barrier: the next round fills the other buffer, and the one after that is behind the next barrier
  }
}

// Parity mode: one thread per frame, states in index order, doubles throughout.
// lna_f32_rows with the scores brought in by TMA: one elected thread fetches a round's [states][32 frames] box into a
// 3-deep shared-memory ring (48 KB in flight per CTA, no registers tied up by loads in flight: the register-staged kernel
// above keeps 16 per thread and sits at 47 % warps active, stalled on them); the warps convert from shared memory -- lane =
// frame, rows of 32 floats: conflict free -- with the arithmetic of lna_f32_rows, bit for bit, and store rows the same way.
// Boxes that reach past the last state / frame are zero filled by the copy, like the guarded loads above.
template <int B, bool NORM, int NST>
__global__ void __launch_bounds__(256)
lna_f32_rows_tma(const __grid_constant__ CUtensorMap mapS, int S, int64_t nf, const float2 *__restrict__ norm, uint8_t *__restrict__ out)
{
  constexpr int SPR = 128;                       // states per round
  constexpr int SPW = SPR / 8;                   // states per warp and round (= one batch of 16)
  constexpr int TW = SPR * B / 4;                // 32-bit words per frame row of the tile
  constexpr uint32_t STAGE_BYTES = SPR * 32 * 4;
  extern __shared__ __align__(128) unsigned char lna_sm[];
  float (*in)[SPR][32] = reinterpret_cast<float (*)[SPR][32]>(lna_sm);
  uint32_t (*tile)[32][TW + 1] = reinterpret_cast<uint32_t (*)[32][TW + 1]>(lna_sm + NST * STAGE_BYTES);
  __shared__ uint64_t full[NST];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const int64_t f = f0 + lane;
  const bool fvalid = f < nf;
  const int rounds = (S + SPR - 1) / SPR;
  if (tid == 0) {
    for (int i = 0; i < NST; i++) tc::mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int r = 0; r < NST && r < rounds; r++) {
      tc::mbar_expect_tx(&full[r], STAGE_BYTES);
      tc::tma_load_2d(in[r], &mapS, (int)f0, r * SPR, &full[r]);
    }
  }
  float Mx = 0.f, lognorm = 0.f;
  if (NORM) {
    const float2 nm = norm[fvalid ? f : 0];
    // every state flushed to zero (Mx = -inf): Z = 1 and every record is the floor; +inf makes L - Mx = -inf
    Mx = (nm.x == -INFINITY) ? INFINITY : nm.x;
    lognorm = nm.y;
  }
  __syncthreads();                               // the barriers exist
  for (int r = 0; r < rounds; r++) {
    const int st = r % NST;
    tc::mbar_wait(&full[st], (uint32_t)(r / NST) & 1u);
    uint32_t(*tl)[TW + 1] = tile[r & 1];
    float lp[SPW];
#pragma unroll
    for (int j = 0; j < SPW; ++j) lp[j] = in[st][w * SPW + j][lane];
#pragma unroll
    for (int j = 0; j < SPW; ++j) {
      float L = lp[j];
      if (L < LN_2M126) L = log_of_float_cast_rare(L);       // fp32-denormal / zero range of the float cast: rare
      const float v = NORM ? (L - Mx) - lognorm : L;
      lp[j] = fmaxf(v, LP_FLOOR);                            // also catches -inf and NaN
    }
    if (B == 2) {
#pragma unroll
      for (int j = 0; j < SPW; j += 2) {
        const uint32_t c0 = NORM ? lna_code16_f32_neg(lp[j]) : lna_code16_f32(lp[j]);
        const uint32_t c1 = NORM ? lna_code16_f32_neg(lp[j + 1]) : lna_code16_f32(lp[j + 1]);
        tl[lane][(w * SPW + j) >> 1] = __byte_perm(c0, c1, 0x4501);   // big-endian codes, two per word
      }
    } else {
#pragma unroll
      for (int j = 0; j < SPW; ++j) tl[lane][w * SPW + j] = __float_as_uint(lp[j]);
    }
    __syncthreads();                             // the stage has been read by every warp, the tile is complete
    if (tid == 0 && r + NST < rounds) {
      tc::mbar_expect_tx(&full[st], STAGE_BYTES);
      tc::tma_load_2d(in[st], &mapS, (int)f0, (r + NST) * SPR, &full[st]);
    }
    const int nwords = min(SPR, S - r * SPR) * B / 4;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int fr = w * 4 + rr;
      const int64_t gf = f0 + fr;
      if (gf < nf) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(out + ((size_t)gf * S + (size_t)r * SPR) * B);
#pragma unroll
        for (int k = 0; k < (TW + 31) / 32; ++k)
          if (lane + 32 * k < nwords) __stcs(dst + lane + 32 * k, tl[fr][lane + 32 * k]);
      }
    }
    // no second barrier: the next round fills the other tile, and the one after that is behind the next barrier
  }
}

template <int B>
__global__ void __launch_bounds__(128)
lna_f64(const double *__restrict__ lin, int64_t ldF, int S, int64_t nf, int normalize, uint8_t *__restrict__ out)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const double *col = lin + f;
  double Z = 0.0;
  for (int s = 0; s < S; ++s) {
    double l = col[(int64_t)s * ldF];
    if (l < 1e-50) l = 1e-50;
    float o = (float)l;
    Z = __dadd_rn(Z, (double)o);
  }
  if (!normalize || Z == 0.0) Z = 1.0;
  const double log_tiny = log(1e-50);
  for (int s = 0; s < S; ++s) {
    double l = col[(int64_t)s * ldF];
    if (l < 1e-50) l = 1e-50;
    float o = (float)l;
    double x = __ddiv_rn((double)o, Z);
    float lp = (float)(x < 1e-50 ? log_tiny : log(x));
    lna_store<B>(out + (f * S + s) * B, lp);
  }
}

__global__ void checksum_kernel(const uint8_t *__restrict__ buf, int64_t nbytes, unsigned long long *acc)
{
  unsigned long long local = 0;
  const int64_t nw = nbytes / 4;
  const uint32_t *wbuf = reinterpret_cast<const uint32_t *>(buf);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = wbuf[i];
    local += (v & 255) + ((v >> 8) & 255) + ((v >> 16) & 255) + (v >> 24);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = nw * 4; i < nbytes; ++i) local += buf[i];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(acc, local);
}

constexpr int LNA_TMA_DEFAULT = 1;      // lna_f32_rows_tma (8.6 -> 7.2 ms per config-2 step); AKUGPU_LNA_TMA=0: the register-staged kernel
typedef CUresult (*LnaEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void lna_make_map(CUtensorMap *map, const float *sll, int S, int64_t nf, int64_t ldF)
{
  static LnaEncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void *p = nullptr;
    AKU_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw Error(AKUGPU_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    fn = (LnaEncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {(cuuint64_t)nf, (cuuint64_t)S};            // innermost first: frames, then states
  cuuint64_t strides[1] = {(cuuint64_t)ldF * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(sll), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(AKUGPU_E_CUDA, fmt("cuTensorMapEncodeTiled (LNA scores) failed (%d)", (int)r));
}

void launch_lna_f32(akugpu_ctx *ctx, const float *sll, int64_t ldF, int S, int64_t nf, int lnabytes, int normalize,
                    const float2 *norm, uint8_t *out, float2 *norm_scratch)
{
  if (nf <= 0) return;
  unsigned grid = (unsigned)((nf + 31) / 32);
  // row-streaming kernel: needs 4-byte aligned rows; the per-frame normaliser comes from the scorer's epilogue or from
  // a pass of its own
  if (((size_t)S * lnabytes) % 4 == 0 && ((uintptr_t)out & 3) == 0 && !getenv("AKUGPU_LNA_OLD")) {
    if (normalize && !norm) {
      if (!norm_scratch) { ctx->d_norm.reserve((size_t)(nf + 31) / 32 * 32 * sizeof(float2)); norm_scratch = ctx->d_norm.as<float2>(); }
      lna_f32_norm<<<grid, 256, 0, ctx->stream>>>(sll, ldF, S, nf, norm_scratch);
      AKU_CUDA(cudaGetLastError());
      ctx->launches++;
      norm = norm_scratch;
    }
    static const int use_tma = getenv("AKUGPU_LNA_TMA") ? atoi(getenv("AKUGPU_LNA_TMA")) : LNA_TMA_DEFAULT;
    if (use_tma && (ldF * 4) % 16 == 0 && ((uintptr_t)sll & 15) == 0 && nf < (int64_t)1 << 31) {
      // scores as a 2-D tensor [S][nf] of floats (row pitch ldF), boxes of 128 states x 32 frames
      CUtensorMap mapS;
      lna_make_map(&mapS, sll, S, nf, ldF);
      const int nst = (use_tma >= 2 && use_tma <= 4) ? use_tma : 3;       // ring depth (AKUGPU_LNA_TMA=2|3|4; 1 = default 3)
      auto go = [&](auto kernel, int B_) {
        const size_t smem = (size_t)nst * 128 * 32 * 4 + 2 * (size_t)32 * (128 * B_ / 4 + 1) * 4;
        ensure_dynamic_smem(ctx, (const void *)kernel, smem);
        kernel<<<grid, 256, smem, ctx->stream>>>(mapS, S, nf, norm, out);
      };
#define LNA_TMA_GO(NST_)                                                         \
      if (lnabytes == 2 && normalize) go(lna_f32_rows_tma<2, true, NST_>, 2);    \
      else if (lnabytes == 2) go(lna_f32_rows_tma<2, false, NST_>, 2);           \
      else if (normalize) go(lna_f32_rows_tma<4, true, NST_>, 4);                \
      else go(lna_f32_rows_tma<4, false, NST_>, 4);
      if (nst == 2) { LNA_TMA_GO(2) } else if (nst == 4) { LNA_TMA_GO(4) } else { LNA_TMA_GO(3) }
#undef LNA_TMA_GO
      AKU_CUDA(cudaGetLastError());
      ctx->launches++;
      return;
    }
    if (lnabytes == 2 && normalize) lna_f32_rows<2, true><<<grid, 256, 0, ctx->stream>>>(sll, ldF, S, nf, norm, out);
    else if (lnabytes == 2) lna_f32_rows<2, false><<<grid, 256, 0, ctx->stream>>>(sll, ldF, S, nf, norm, out);
    else if (normalize) lna_f32_rows<4, true><<<grid, 256, 0, ctx->stream>>>(sll, ldF, S, nf, norm, out);
    else lna_f32_rows<4, false><<<grid, 256, 0, ctx->stream>>>(sll, ldF, S, nf, norm, out);
    AKU_CUDA(cudaGetLastError());
    ctx->launches++;
    return;
  }
  if (lnabytes == 2) lna_f32<2><<<grid, 256, 0, ctx->stream>>>(sll, ldF, S, nf, normalize, norm, out);
  else lna_f32<4><<<grid, 256, 0, ctx->stream>>>(sll, ldF, S, nf, normalize, norm, out);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_lna_f64(akugpu_ctx *ctx, const double *lin, int64_t ldF, int S, int64_t nf, int lnabytes, int normalize,
                    uint8_t *out)
{
  if (nf <= 0) return;
  unsigned grid = (unsigned)((nf + 127) / 128);
  if (lnabytes == 2) lna_f64<2><<<grid, 128, 0, ctx->stream>>>(lin, ldF, S, nf, normalize, out);
  else lna_f64<4><<<grid, 128, 0, ctx->stream>>>(lin, ldF, S, nf, normalize, out);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_checksum(akugpu_ctx *ctx, const uint8_t *buf, int64_t nbytes, unsigned long long *acc)
{
  if (nbytes <= 0) return;
  int grid = ctx->sm_count * 8;
  checksum_kernel<<<grid, 256, 0, ctx->stream>>>(buf, nbytes, acc);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace akugpu
