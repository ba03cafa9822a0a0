// gmm_kernels.cu -- diagonal-Gaussian mixture scorer (K5) for sm_100a.
//
// Replaces, for a whole tile of frames at once, the per-frame CPU loops
//   DiagonalGaussian::compute_log_likelihood   aku/Distributions.cc:1041-1062
//   DiagonalGaussian::compute_likelihood       aku/Distributions.cc:1034-1037
//   Mixture::compute_likelihood                aku/Distributions.cc:2079-2086
//   PDFPool::precompute_likelihoods            aku/Distributions.cc:2648-2682
//   HmmSet::precompute_likelihoods             aku/HmmSet.cc:485-501
//
// Two kernels:
//   gmm_diag_f32<GR,F2>  throughput mode.  Direct form on the FP32 pipe:
//        t = x'*s + m ; acc += t*t        (2 FMA per frame x component x dim)
//     with s = sqrt(prec/2), m = -(mu-centre)*s prepared in double on the host, and
//     x' = x - centre.  Tensor cores are not used: the expanded (GEMM) form loses the
//     result to cancellation at fp32 and is off by orders of magnitude at tf32/bf16
//     (experiment 1b).  Mixture log-sum-exp is fused.
//   gmm_diag_f64         parity mode.  Same operation order as the reference in
//     double (no FMA contraction), linear-domain mixture sum, 1e-50 floor.
//
// Data layout (fp32 image, built by pack_model_f32 in model.cu)
//   component slots are grouped into tiles of TC = 16*GR slots; a state's slots are
//   contiguous, padded to a multiple of GR, and never straddle a tile.  One tile is
//   one contiguous "stage image" in HBM:
//        float4 P[DP][TC]   {s(2dp), s(2dp+1), m(2dp), m(2dp+1)}  slot-permuted
//        float  C[TC]       -(log w + log sqrt(prod prec))   (i.e. -c ; +1e30 for pads)
//   fetched with ONE cp.async.bulk (TMA 1-D, UBLKCP) per stage into a 2-deep ring,
//   completion on an mbarrier.  Slot permutation: the j-th slot of thread-group cg
//   sits at position j*16+cg so that a warp's LDS.128 are conflict free.
//
// CTA = 256 threads = 16 frame-groups x 16 component-groups; thread tile = 8 frames
// x GR components; CTA tile = 128 frames x TC components per stage.  Output is
// state-major  sll[state][ldF]  (fp32 natural-log likelihood), coalesced over frames,
// which is also the order the LNA epilogue (lna_kernels.cu) sweeps it in.
#include "ctx.hpp"
#include "kernels.hpp"

namespace akugpu {

constexpr int TF = 64;           // frames per CTA tile
constexpr int FR = 8;            // frames per thread
constexpr int NTH = 128;
constexpr int NFG = TF / FR;     // 8 frame groups
constexpr int NCG = NTH / NFG;   // 16 component groups
#ifndef GMM_DP_UNROLL
#define GMM_DP_UNROLL 2
#endif
constexpr int TAB_INTS = 32;     // per-tile state table appended to the stage image
constexpr int kDpUnroll = GMM_DP_UNROLL;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int GR, bool F2>
struct AccT;
template <int GR>
struct AccT<GR, false> { float v[FR][GR]; };
template <int GR>
struct AccT<GR, true> { float2 v[FR][GR]; };

// One (frame, component, dim-pair) update: t = x*s + m ; acc += t*t.
template <bool F2, bool FIRST, class A>
__device__ __forceinline__ void gmm_step(A &acc, float x0, float x1, const float4 &p, float nc)
{
  if constexpr (F2) {
    float2 tt = __ffma2_rn(make_float2(x0, x1), make_float2(p.x, p.y), make_float2(p.z, p.w));
    acc = __ffma2_rn(tt, tt, FIRST ? make_float2(nc, 0.f) : acc);
  } else {
    float t0 = fmaf(x0, p.x, p.z);
    float t1 = fmaf(x1, p.y, p.w);
    acc = fmaf(t0, t0, FIRST ? nc : acc);
    acc = fmaf(t1, t1, acc);
  }
}

// ------------------------------------------------------------------------------------
// grid.x = frame tiles (64 frames), grid.y = split of the component tiles (only used when there
// are too few frame tiles to fill the chip).  4 CTAs of 128 threads are resident per SM so that
// one CTA's log-sum-exp epilogue hides under the other CTAs' FMA loops.
template <int GR, bool F2>
__global__ void __launch_bounds__(NTH, F2 ? 4 : 2)
gmm_diag_f32(const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t f_end, int D, int DP,
             const float *__restrict__ params, size_t tile_floats, int n_tiles, int tiles_per_cta,
             const float *__restrict__ center, const double *__restrict__ center64,
             float *__restrict__ sll, int64_t ldF)
{
  constexpr int TC = NCG * GR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t stage_bytes = (uint32_t)(tile_floats * sizeof(float));
  float2 *xs = reinterpret_cast<float2 *>(smem_raw);                         // [DP][TF]
  unsigned char *stage0 = smem_raw + (size_t)DP * TF * sizeof(float2);
  __shared__ uint64_t full_bar[2];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int fg = lane & 7;                 // frame group: frames k*16 + fg*2 + {0,1}, k = 0..3
  const int cg = warp * 4 + (lane >> 3);   // component group

  const int64_t f0 = f_begin + (int64_t)blockIdx.x * TF;
  const int t_begin = blockIdx.y * tiles_per_cta;
  const int t_end = min(n_tiles, t_begin + tiles_per_cta);
  if (t_begin >= t_end) return;

  if (tid == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&full_bar[0], stage_bytes);
    tma_load_1d(stage0, params + (size_t)t_begin * tile_floats, stage_bytes, &full_bar[0]);
    if (t_begin + 1 < t_end) {
      mbar_expect_tx(&full_bar[1], stage_bytes);
      tma_load_1d(stage0 + stage_bytes, params + (size_t)(t_begin + 1) * tile_floats, stage_bytes, &full_bar[1]);
    }
  }

  // Frame tile -> shared, transposed to [dim pair][frame] and centred.
  {
    const int D2 = 2 * DP;
    float *xsf = reinterpret_cast<float *>(xs);
    for (int idx = tid; idx < TF * D2; idx += NTH) {
      int fr = idx / D2, d = idx - fr * D2;
      int64_t gf = f0 + fr;
      float v = 0.f;
      if (d < D && gf < f_end) {
        if (feats_f64)
          v = (float)(reinterpret_cast<const double *>(feats)[gf * D + d] - center64[d]);
        else
          v = reinterpret_cast<const float *>(feats)[gf * D + d] - center[d];
      }
      xsf[((d >> 1) * TF + fr) * 2 + (d & 1)] = v;
    }
  }
  __syncthreads();

  uint32_t phase[2] = {0, 0};
  for (int t = t_begin; t < t_end; ++t) {
    const int b = (t - t_begin) & 1;
    unsigned char *stage = stage0 + (size_t)b * stage_bytes;
    const float4 *ps = reinterpret_cast<const float4 *>(stage);
    const float *cs = reinterpret_cast<const float *>(stage + (size_t)DP * TC * sizeof(float4));
    const int *tab = reinterpret_cast<const int *>(cs + TC);
    mbar_wait(&full_bar[b], phase[b]);
    phase[b] ^= 1;

    AccT<GR, F2> acc;
#pragma unroll
    for (int j = 0; j < GR; ++j) {
      const float nc = cs[j * NCG + cg];     // accumulators start at -c
#pragma unroll
      for (int i = 0; i < FR; ++i) {
        if constexpr (F2) acc.v[i][j] = make_float2(nc, 0.f);
        else acc.v[i][j] = nc;
      }
    }
#pragma unroll (kDpUnroll)
    for (int dp = 0; dp < DP; ++dp) {
      float4 xv[FR / 2], pv[GR];
#pragma unroll
      for (int k = 0; k < FR / 2; ++k) xv[k] = *reinterpret_cast<const float4 *>(&xs[dp * TF + k * 16 + fg * 2]);
#pragma unroll
      for (int j = 0; j < GR; ++j) pv[j] = ps[dp * TC + j * NCG + cg];
#pragma unroll
      for (int i = 0; i < FR; ++i) {
        const float x0 = (i & 1) ? xv[i >> 1].z : xv[i >> 1].x;
        const float x1 = (i & 1) ? xv[i >> 1].w : xv[i >> 1].y;
#pragma unroll
        for (int j = 0; j < GR; ++j) gmm_step<F2, false>(acc.v[i][j], x0, x1, pv[j], 0.f);
      }
    }
    __syncthreads();   // everyone is done reading this stage's parameters: reuse its head for the partials
                       // (the state table sits at the tail of the image and stays intact)

    // Thread-level log-sum-exp over its GR components: a = min(-ll), sum = sum exp(ll + a).
    float2 *part = reinterpret_cast<float2 *>(stage);   // [NCG][TF]
#pragma unroll
    for (int i = 0; i < FR; ++i) {
      float v[GR];
#pragma unroll
      for (int j = 0; j < GR; ++j) {
        if constexpr (F2) v[j] = acc.v[i][j].x + acc.v[i][j].y;
        else v[j] = acc.v[i][j];
      }
      float a = v[0];
#pragma unroll
      for (int j = 1; j < GR; ++j) a = fminf(a, v[j]);
      const float al = a * LOG2E;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < GR; ++j) sum += ex2f(fmaf(v[j], -LOG2E, al));
      const int fr = (i >> 1) * 16 + fg * 2 + (i & 1);
      part[cg * TF + fr] = make_float2(a, sum);
    }
    __syncthreads();

    // Combine the thread-groups of each state of this tile; write state log-likelihoods.
    {
      const int nst = tab[0], s0 = tab[1];
      const int fr = tid & (TF - 1);
      // thread handles frame (tid & 63) of local states (tid >> 6), +2, +4, ..
      for (int ls = tid >> 6; ls < nst; ls += NTH / TF) {
        {
          const int e = tab[2 + ls];
          const int g0 = e & 255, ng = e >> 8;
          float A = part[g0 * TF + fr].x;
          for (int g = 1; g < ng; ++g) A = fminf(A, part[(g0 + g) * TF + fr].x);
          float tot = 0.f;
          for (int g = 0; g < ng; ++g) {
            float2 p = part[(g0 + g) * TF + fr];
            tot = fmaf(p.y, ex2f((A - p.x) * LOG2E), tot);
          }
          sll[(int64_t)(s0 + ls) * ldF + (f0 - f_begin) + fr] = fmaf(lg2f(tot), LN2, -A);
        }
      }
    }
    __syncthreads();
    if (tid == 0 && t + 2 < t_end) {
      fence_proxy_async();
      mbar_expect_tx(&full_bar[b], stage_bytes);
      tma_load_1d(stage, params + (size_t)(t + 2) * tile_floats, stage_bytes, &full_bar[b]);
    }
  }
}

// ------------------------------------------------------------------------------------
// Parity mode: the reference's arithmetic, operation by operation, in double.
//   ll = sum_i (f_i-mu_i)*(f_i-mu_i)*prec_i  (left to right, separate mul/mul/add)
//   ll = ll*(-0.5) + const ; lik = exp(ll)                aku/Distributions.cc:1052-1059,1036
//   l  = sum_k w_k*lik_k (in component order)             aku/Distributions.cc:2082-2084
//   l  = max(l, 1e-50)                                    aku/HmmSet.cc:497-498
// CTA: 128 frames; 8 warps take states round-robin; lane handles frames lane+32*i.
__global__ void __launch_bounds__(256)
gmm_diag_f64(const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t f_end, int D,
             const double *__restrict__ mean, const double *__restrict__ prec, const double *__restrict__ cst,
             const int *__restrict__ mix_off, const int *__restrict__ mix_gauss, const double *__restrict__ mix_w,
             int S, double *__restrict__ lin, int64_t ldF)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *xs = reinterpret_cast<double *>(smem_raw);   // [D][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t f0 = f_begin + (int64_t)blockIdx.x * 128;
  for (int idx = tid; idx < 128 * D; idx += 256) {
    int fr = idx / D, d = idx - fr * D;
    int64_t gf = f0 + fr;
    double v = 0.0;
    if (gf < f_end)
      v = feats_f64 ? reinterpret_cast<const double *>(feats)[gf * D + d]
                    : (double)reinterpret_cast<const float *>(feats)[gf * D + d];
    xs[d * 128 + fr] = v;
  }
  __syncthreads();
  for (int s = blockIdx.y * 8 + warp; s < S; s += gridDim.y * 8) {
    double l[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = mix_off[s]; k < mix_off[s + 1]; ++k) {
      const int g = mix_gauss[k];
      const double w = mix_w[k];
      const double *mu = mean + (size_t)g * D;
      const double *pr = prec + (size_t)g * D;
      double ll[4] = {0.0, 0.0, 0.0, 0.0};
      for (int d = 0; d < D; ++d) {
        const double m = mu[d], p = pr[d];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double dd = __dsub_rn(xs[d * 128 + lane + 32 * i], m);
          ll[i] = __dadd_rn(ll[i], __dmul_rn(__dmul_rn(dd, dd), p));
        }
      }
      const double c = cst[g];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double v = __dadd_rn(__dmul_rn(ll[i], -0.5), c);
        l[i] = __dadd_rn(l[i], __dmul_rn(w, exp(v)));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double v = l[i];
      if (v < 1e-50) v = 1e-50;
      lin[(int64_t)s * ldF + (f0 - f_begin) + lane + 32 * i] = v;
    }
  }
}

// [S][ldF] state-major  ->  [F][S] frame-major (the layout akugpu_gmm_score returns).
template <class T>
__global__ void transpose_sf(const T *__restrict__ in, int64_t ldF, int S, int64_t F, T *__restrict__ out)
{
  __shared__ T tile[32][33];
  int64_t fb = (int64_t)blockIdx.x * 32;
  int sb = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int s = sb + r;
    int64_t f = fb + threadIdx.x;
    if (s < S && f < F) tile[r][threadIdx.x] = in[(int64_t)s * ldF + f];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    int64_t f = fb + r;
    int s = sb + threadIdx.x;
    if (s < S && f < F) out[f * S + s] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------------------
// launchers
size_t gmm_f32_smem_bytes(const PackedF32 &p) {
  return (size_t)p.DP * TF * sizeof(float2) + 2 * p.tile_floats * sizeof(float);
}

template <int GR, bool F2>
static void launch_f32_t(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end,
                         float *sll, int64_t ldF)
{
  const PackedF32 &p = ctx->p32;
  const HostModel &hm = ctx->hm;
  size_t smem = gmm_f32_smem_bytes(p);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AKU_CUDA(cudaFuncSetAttribute(gmm_diag_f32<GR, F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int64_t nf = f_end - f_begin;
  int ftiles = (int)((nf + TF - 1) / TF);
  // One wave = sm_count * resident CTAs.  Full chunks are sized to whole waves by the caller
  // (gmm_wave_frames); a short chunk splits the component tiles over grid.y to fill the chip.
  const int wave = ctx->sm_count * (F2 ? 4 : 2);
  int ysplit = 1;
  if (ftiles < wave) ysplit = std::min(p.n_tiles, std::max(1, wave / ftiles));
  int tiles_per_cta = (p.n_tiles + ysplit - 1) / ysplit;
  ysplit = (p.n_tiles + tiles_per_cta - 1) / tiles_per_cta;
  dim3 grid(ftiles, ysplit);
  gmm_diag_f32<GR, F2><<<grid, NTH, smem, ctx->stream>>>(
      feats, feats_f64, f_begin, f_end, hm.D, p.DP, p.params.as<float>(), p.tile_floats, p.n_tiles, tiles_per_cta,
      p.center.as<float>(), p.center64.as<double>(), sll, ldF);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

// Frames per full wave of the fp32 scorer (chunks should be a multiple of this).
int64_t gmm_wave_frames(akugpu_ctx *ctx) { return (int64_t)ctx->sm_count * (ctx->p32.GR == 4 ? 4 : 2) * TF; }
int gmm_frame_tile() { return TF; }

void launch_gmm_f32(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, float *sll,
                    int64_t ldF)
{
  if (ctx->p32.GR == 4) launch_f32_t<4, true>(ctx, feats, feats_f64, f_begin, f_end, sll, ldF);
  else launch_f32_t<8, false>(ctx, feats, feats_f64, f_begin, f_end, sll, ldF);
}

void launch_gmm_f64(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, double *lin,
                    int64_t ldF)
{
  const HostModel &hm = ctx->hm;
  const PackedF64 &p = ctx->p64;
  size_t smem = (size_t)hm.D * 128 * sizeof(double);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AKU_CUDA(cudaFuncSetAttribute(gmm_diag_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int64_t nf = f_end - f_begin;
  int ftiles = (int)((nf + 127) / 128);
  int ysplit = std::max(1, std::min((hm.S + 7) / 8, (4 * ctx->sm_count + ftiles - 1) / ftiles));
  dim3 grid(ftiles, ysplit);
  gmm_diag_f64<<<grid, 256, smem, ctx->stream>>>(feats, feats_f64, f_begin, f_end, hm.D, p.mean.as<double>(),
                                                 p.prec.as<double>(), p.cst.as<double>(), p.mix_off.as<int>(),
                                                 p.mix_gauss.as<int>(), p.mix_w.as<double>(), hm.S, lin, ldF);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_transpose_f32(akugpu_ctx *ctx, const float *in, int64_t ldF, int S, int64_t F, float *out)
{
  dim3 grid((unsigned)((F + 31) / 32), (S + 31) / 32), block(32, 8);
  transpose_sf<float><<<grid, block, 0, ctx->stream>>>(in, ldF, S, F, out);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}
void launch_transpose_f64(akugpu_ctx *ctx, const double *in, int64_t ldF, int S, int64_t F, double *out)
{
  dim3 grid((unsigned)((F + 31) / 32), (S + 31) / 32), block(32, 8);
  transpose_sf<double><<<grid, block, 0, ctx->stream>>>(in, ldF, S, F, out);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

// ------------------------------------------------------------------------------------
// Issue-pipe micro-benchmarks (akugpu_pipe_rates): 8 independent chains per thread, enough warps to
// saturate each pipe.  Modes 0-3 are the classic constant-operand chains; 4-7 use the scorer's own
// operand pattern (all-register, no memory) to expose register-bandwidth limits:
//   4: a = fma(b, c, a) with three distinct varying registers
//   5: same with FFMA2
//   6: the scorer's inner tile, FFMA:  t = fma(x_i, s_j, m_j); acc_ij = fma(t, t, acc_ij)   (8x8)
//   7: the scorer's inner tile, FFMA2 (8x4 packed)
template <int MODE>
__global__ void pipe_rate_kernel(float *out, int iters)
{
  float a[8], b[8], c[8];
  double da[8];
  float2 a2[8], b2[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-3f + i;
    b[i] = 0.999f + (threadIdx.x + i) * 1e-9f;
    c[i] = 1e-3f * (i + 1);
    da[i] = a[i];
    a2[i] = make_float2(a[i], a[i] + 0.5f);
    b2[i] = make_float2(b[i], b[i]);
    c2[i] = make_float2(c[i], c[i]);
  }
  const float m = 0.999f + threadIdx.x * 1e-9f, cc = 1e-3f;
  float acc[8][8];
  float2 acc2[8][4];
  if (MODE == 6) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = i + j;
  }
  if (MODE == 7) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(i, j);
  }
  for (int it = 0; it < iters; ++it) {
    if (MODE <= 5) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (MODE == 0) a[i] = fmaf(a[i], m, cc);
          if (MODE == 1) a2[i] = __ffma2_rn(a2[i], make_float2(m, m), make_float2(cc, cc));
          if (MODE == 2) da[i] = fma(da[i], (double)m, (double)cc);
          if (MODE == 3) a[i] = ex2f(a[i]);
          if (MODE == 4) a[i] = fmaf(b[i], c[i], a[i]);
          if (MODE == 5) a2[i] = __ffma2_rn(b2[i], c2[i], a2[i]);
        }
      }
    } else if (MODE == 6) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = fmaf(a[i], b[j], c[j]);
          acc[i][j] = fmaf(t, t, acc[i][j]);
        }
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += 1e-7f;   // keep the loop from being hoisted
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t = __ffma2_rn(a2[i], b2[j], c2[j]);
          acc2[i][j] = __ffma2_rn(t, t, acc2[i][j]);
        }
#pragma unroll
      for (int i = 0; i < 8; ++i) a2[i].x += 1e-7f;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + (float)da[i] + a2[i].x + a2[i].y;
  if (MODE == 6) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += acc[i][j];
  }
  if (MODE == 7) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s += acc2[i][j].x + acc2[i][j].y;
  }
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
static float time_pipe(akugpu_ctx *ctx, float *d, int iters, int blocks, int threads, cudaEvent_t e0, cudaEvent_t e1)
{
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    AKU_CUDA(cudaEventRecord(e0, ctx->stream));
    pipe_rate_kernel<MODE><<<blocks, threads, 0, ctx->stream>>>(d, iters);
    AKU_CUDA(cudaEventRecord(e1, ctx->stream));
    AKU_CUDA(cudaEventSynchronize(e1));
    float ms;
    AKU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
    ctx->launches++;
  }
  return best;
}

void pipe_rates(akugpu_ctx *ctx, double out[8])
{
  DevBuf d; d.reserve(16);
  cudaEvent_t e0, e1;
  AKU_CUDA(cudaEventCreate(&e0));
  AKU_CUDA(cudaEventCreate(&e1));
  const int iters = 2048, blocks = ctx->sm_count * 4, threads = 256;
  float ms[8];
  ms[0] = time_pipe<0>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  ms[1] = time_pipe<1>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  ms[2] = time_pipe<2>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  ms[3] = time_pipe<3>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  ms[4] = time_pipe<4>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  ms[5] = time_pipe<5>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  ms[6] = time_pipe<6>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  ms[7] = time_pipe<7>(ctx, d.as<float>(), iters, blocks, threads, e0, e1);
  const double per_iter[8] = {32, 64, 32, 32, 32, 64, 128, 128};   // lane-ops per thread per iteration
  for (int mode = 0; mode < 8; ++mode) out[mode] = (double)blocks * threads * iters * per_iter[mode] / (ms[mode] * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

}  // namespace akugpu
