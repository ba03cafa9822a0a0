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
    lognorm = log1p(Rt);
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

// Parity mode: one thread per frame, states in index order, doubles throughout.
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

void launch_lna_f32(akugpu_ctx *ctx, const float *sll, int64_t ldF, int S, int64_t nf, int lnabytes, int normalize,
                    const float2 *norm, uint8_t *out)
{
  if (nf <= 0) return;
  unsigned grid = (unsigned)((nf + 31) / 32);
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
