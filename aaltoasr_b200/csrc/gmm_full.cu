// gmm_full.cu -- full-covariance (and mixed diag/full pool) scoring in double: parity path.
//
// Follows the reference's exponential-family evaluation
//   PDFPool::precompute_likelihoods            aku/Distributions.cc:2663-2682
//     exponential feature  phi = [f ; map_m2v(f f^T)]   (lower triangle, off-diagonals x sqrt 2)
//   FullCovarianceGaussian::compute_log_likelihood_exponential   aku/Distributions.cc:1437-1446
//     ll = <phi, theta> + normalizer + constant ,  lik = exp(ll)
//   diagonal members of a "variable" pool keep DiagonalGaussian::compute_likelihood (:1034-1062)
//   Mixture::compute_likelihood (:2079-2086) and the 1e-50 floor (aku/HmmSet.cc:497-498)
// i.e. a dense [frames x L] . [L x G_full] contraction with L = D(D+3)/2 (819 at D = 39), done here as a
// register-tiled double GEMM on the FP64 pipe with every dot product accumulated in index order
// (separate multiply and add, like the x86 build).  A tensor-core (bf16x3 / tf32x3 split) throughput
// mode for this contraction is the next step for BASELINE config 5; this kernel is its parity anchor.
#include "ctx.hpp"
#include "kernels.hpp"

namespace akugpu {

__global__ void expfeat_f64(const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t nf, int D, int L,
                            double *__restrict__ phi)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nf * L) return;
  const int64_t fr = i / L;
  const int l = (int)(i - fr * L);
  auto x = [&](int d) -> double {
    return feats_f64 ? reinterpret_cast<const double *>(feats)[(f_begin + fr) * D + d]
                     : (double)reinterpret_cast<const float *>(feats)[(f_begin + fr) * D + d];
  };
  double v;
  if (l < D) {
    v = x(l);
  } else {
    int pos = l - D, r = 0;
    while ((r + 1) * (r + 2) / 2 <= pos) r++;      // row r of the lower triangle holds r+1 entries
    const int c = pos - r * (r + 1) / 2;
    const double m = __dmul_rn(x(r), x(c));
    v = (r == c) ? m : __dmul_rn(sqrt(2.0), m);
  }
  phi[fr * L + l] = v;
}

// lik_g[gauss][frame] = exp( <phi[frame], theta[:, j]> + norm[j] + cst[j] ) for the full Gaussians.
// CTA 16x16 threads, 64 frames x 64 Gaussians, thread tile 4x4, k-chunks of 16 through shared memory.
__global__ void __launch_bounds__(256)
fullcov_gemm_f64(const double *__restrict__ phi, int64_t nf, int L, const double *__restrict__ theta, int n_full,
                 const double *__restrict__ fnorm, const double *__restrict__ fcst, const int *__restrict__ full_gauss,
                 double *__restrict__ lik_g, int64_t ldF)
{
  __shared__ double sp[16][64 + 1];   // [k][frame]
  __shared__ double st[16][64 + 1];   // [k][gauss]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // tx: frames, ty: gaussians
  const int64_t fb = (int64_t)blockIdx.x * 64;
  const int gb = blockIdx.y * 64;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < L; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int fr = i >> 4, k = i & 15;           // phi is [frame][L]: k fastest
      const int64_t f = fb + fr;
      sp[k][fr] = (f < nf && k0 + k < L) ? phi[f * L + k0 + k] : 0.0;
      const int kk = i >> 6, g = i & 63;           // theta is [L][n_full]: gauss fastest
      st[kk][g] = (gb + g < n_full && k0 + kk < L) ? theta[(size_t)(k0 + kk) * n_full + gb + g] : 0.0;
    }
    __syncthreads();
    const int kmax = min(16, L - k0);
    for (int k = 0; k < kmax; k++) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { a[q] = sp[k][tx + 16 * q]; b[q] = st[k][ty + 16 * q]; }
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int r = 0; r < 4; r++) acc[q][r] = __dadd_rn(acc[q][r], __dmul_rn(a[q], b[r]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int j = gb + ty + 16 * r;
    if (j >= n_full) continue;
    const double nrm = fnorm[j], c = fcst[j];
    const int g = full_gauss[j];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int64_t f = fb + tx + 16 * q;
      if (f < nf) lik_g[(int64_t)g * ldF + f] = exp(__dadd_rn(__dadd_rn(acc[q][r], nrm), c));
    }
  }
}

// Diagonal members of a mixed pool: DiagonalGaussian::compute_likelihood, one thread per (gaussian, frame).
__global__ void gauss_diag_f64(const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t nf, int D,
                               const double *__restrict__ mean, const double *__restrict__ prec,
                               const double *__restrict__ cst, const int *__restrict__ diag_gauss, int n_diag,
                               double *__restrict__ lik_g, int64_t ldF)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (f >= nf || j >= n_diag) return;
  const int g = diag_gauss[j];
  double ll = 0.0;
  for (int d = 0; d < D; d++) {
    const double x = feats_f64 ? reinterpret_cast<const double *>(feats)[(f_begin + f) * D + d]
                               : (double)reinterpret_cast<const float *>(feats)[(f_begin + f) * D + d];
    const double dd = __dsub_rn(x, mean[(size_t)g * D + d]);
    ll = __dadd_rn(ll, __dmul_rn(__dmul_rn(dd, dd), prec[(size_t)g * D + d]));
  }
  lik_g[(int64_t)g * ldF + f] = exp(__dadd_rn(__dmul_rn(ll, -0.5), cst[g]));
}

// Mixture::compute_likelihood in component order + the 1e-50 floor.
__global__ void mixture_f64(const double *__restrict__ lik_g, int64_t ldF, int64_t nf, const int *__restrict__ mix_off,
                            const int *__restrict__ mix_gauss, const double *__restrict__ mix_w, int S,
                            double *__restrict__ lin, int64_t ld_out)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (f >= nf || s >= S) return;
  double l = 0.0;
  for (int k = mix_off[s]; k < mix_off[s + 1]; k++)
    l = __dadd_rn(l, __dmul_rn(mix_w[k], lik_g[(int64_t)mix_gauss[k] * ldF + f]));
  if (l < 1e-50) l = 1e-50;
  lin[(int64_t)s * ld_out + f] = l;
}

// (float) safe_log(x): log(tiny) below tiny (tiny = 0: plain log)
__global__ void lin_to_log_f32(const double *__restrict__ lin, int64_t n, double tiny, float *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double x = lin[i]; out[i] = (float)log(x < tiny ? tiny : x); }
}
__global__ void floor_f32(float *__restrict__ x, int64_t n, float floor_at)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], floor_at);          // NaN -> floor
}

// ConstrainedMllr::AdaptedFeatureVector::calculate_new_ada_vector (aku/ModelModules.hh:208-212): o = b + A f, in double.
template <class T>
__global__ void affine_rows(const T *__restrict__ in, int64_t F, int D, const double *__restrict__ Ab, T *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * D) return;
  const int64_t r = i / D;
  const int c = (int)(i - r * D);
  const T *x = in + r * D;
  const double *a = Ab + (size_t)c * D;
  double acc = 0;                      // row sum in index order, then + b: separate multiplies and adds like the CPU code
  for (int j = 0; j < D; j++) acc = __dadd_rn(acc, __dmul_rn(a[j], (double)x[j]));
  out[i] = (T)__dadd_rn(acc, Ab[(size_t)D * D + c]);
}

// Frames [f_begin, f_end) -> lin[state][ldF] (linear double likelihoods, floored) for pools with full Gaussians.
void launch_gmm_full_f64(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, double *lin,
                         int64_t ldF)
{
  const HostModel &hm = ctx->hm;
  const PackedF64 &p = ctx->p64;
  const int64_t nf = f_end - f_begin;
  if (nf <= 0) return;
  cudaStream_t st = ctx->stream;
  const int64_t ldG = (nf + 127) / 128 * 128;
  ctx->d_fe[2].reserve((size_t)nf * p.L * sizeof(double));          // phi
  ctx->d_fe[3].reserve((size_t)hm.G * ldG * sizeof(double));        // per-Gaussian likelihoods
  double *phi = ctx->d_fe[2].as<double>(), *lik_g = ctx->d_fe[3].as<double>();
  const int64_t ne = nf * p.L;
  expfeat_f64<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(feats, feats_f64, f_begin, nf, hm.D, p.L, phi);
  dim3 grid((unsigned)((nf + 63) / 64), (p.n_full + 63) / 64);
  fullcov_gemm_f64<<<grid, 256, 0, st>>>(phi, nf, p.L, p.theta.as<double>(), p.n_full, p.full_norm.as<double>(),
                                         p.full_cst.as<double>(), p.full_gauss.as<int>(), lik_g, ldG);
  ctx->launches += 2;
  const int n_diag = hm.G - p.n_full;
  if (n_diag > 0) {
    dim3 g2((unsigned)((nf + 127) / 128), n_diag);
    gauss_diag_f64<<<g2, 128, 0, st>>>(feats, feats_f64, f_begin, nf, hm.D, p.mean.as<double>(), p.prec.as<double>(),
                                       p.cst.as<double>(), p.diag_gauss.as<int>(), n_diag, lik_g, ldG);
    ctx->launches++;
  }
  dim3 g3((unsigned)((nf + 127) / 128), hm.S);
  mixture_f64<<<g3, 128, 0, st>>>(lik_g, ldG, nf, p.mix_off.as<int>(), p.mix_gauss.as<int>(), p.mix_w.as<double>(), hm.S,
                                  lin, ldF);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_lin_to_log_f32(akugpu_ctx *ctx, const double *lin, int64_t n, float *out, double tiny)
{
  lin_to_log_f32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(lin, n, tiny, out);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}
void launch_affine_rows(akugpu_ctx *ctx, const void *in, int is_f64, int64_t F, int D, const double *Ab, void *out)
{
  if (F <= 0) return;
  const unsigned blocks = (unsigned)((F * D + 255) / 256);
  if (is_f64) affine_rows<double><<<blocks, 256, 0, ctx->stream>>>((const double *)in, F, D, Ab, (double *)out);
  else affine_rows<float><<<blocks, 256, 0, ctx->stream>>>((const float *)in, F, D, Ab, (float *)out);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}
void launch_floor_f32(akugpu_ctx *ctx, float *x, int64_t n, float floor_at)
{
  floor_f32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(x, n, floor_at);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace akugpu
