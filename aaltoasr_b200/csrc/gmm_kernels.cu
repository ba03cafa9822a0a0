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
// Data layout (fp32 image, built by model_pack in model.cu)
//   A "slot" is 16 component positions of ONE state (padded with weight-0 components); a state
//   with K components owns ceil(K/16) slots.  Slots are dealt into 8 queues, one per warp of a
//   CTA, a state's slots staying consecutive in one queue.  Tile t = the t-th slot of each queue
//   = 128 components, stored as one contiguous "stage image" in HBM:
//        float4 P[DP][128]  {s(2dp), s(2dp+1), m(2dp), m(2dp+1)}     component = warp*16 + j
//        float  C[128]      -(log w + log sqrt(prod prec))           (+1e30 for padding)
//        int    META[8]     per warp: state<<2 | first<<1 | last     (-1: padding slot)
//   fetched with ONE cp.async.bulk (TMA 1-D, UBLKCP) per stage into a 2-deep ring, completion on
//   an mbarrier; the last warp to finish a stage re-arms its buffer (no CTA-wide barrier).
//
// CTA = 256 threads = 8 independent warps over the same 128 frames; warp tile = 128 frames x 16
// components, thread tile = 4 frames x 16 components.  The FMAs are packed FFMA2 over FRAME
// pairs: (x[f0],x[f1]) * s + m with s and m as scalar-broadcast operands (SASS `R.F32`), which
// measured 94-96% of the FP32 pipe in isolation against 74% for three 64-bit operands
// (scripts/micro_occ2.cu).  Because a thread owns every component of its slot, the mixture
// log-sum-exp is entirely in registers, carried across the slots of a big state; there is no
// shared-memory exchange and no __syncthreads in the loop.  Output is state-major
// sll[state][ldF] (fp32 natural-log likelihood), coalesced 256-byte row segments per warp,
// which is also the order the LNA epilogue (lna_kernels.cu) sweeps it in.
#include "ctx.hpp"
#include "kernels.hpp"

namespace akugpu {

constexpr int TF = 128;          // frames per CTA (and per warp) tile
constexpr int NTH = 256;
constexpr int NW = NTH / 32;     // warps = slots per tile
constexpr int GR = 16;           // components per slot
constexpr int TC = NW * GR;      // components per tile
constexpr int NPAIR = TF / 2;    // frame pairs per tile
constexpr int CTAS_PER_SM = 2;
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

// grid.x = frame tiles (128 frames); grid.y = ranges of component tiles (more than one only when
// there are too few frame tiles to fill the chip; range boundaries never cut a state's slots).
// 2 CTAs x 8 warps are resident per SM; every warp runs its own tile loop.
// DPC > 0: number of dim pairs known at compile time (20 for 39/40-dim features); 0: runtime.
template <bool F2, int DPC>
__global__ void __launch_bounds__(NTH, CTAS_PER_SM)
gmm_diag_f32(const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t f_end, int D, int DP_rt,
             const float *__restrict__ params, size_t tile_floats, const int *__restrict__ range_begin,
             const float *__restrict__ center, const double *__restrict__ center64,
             float *__restrict__ sll, int64_t ldF, int dbg)
{
  const int DP = DPC > 0 ? DPC : DP_rt;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t stage_bytes = (uint32_t)(tile_floats * sizeof(float));
  // xs[dp][pair] = {x[f0][2dp], x[f1][2dp], x[f0][2dp+1], x[f1][2dp+1]},  f0 = 2*pair, f1 = 2*pair+1
  float4 *xs = reinterpret_cast<float4 *>(smem_raw);
  unsigned char *stage0 = smem_raw + (size_t)DP * NPAIR * sizeof(float4);
  __shared__ uint64_t full_bar[2];
  __shared__ int done_cnt[2];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t f0 = f_begin + (int64_t)blockIdx.x * TF;
  const int t_begin = range_begin[blockIdx.y], t_end = range_begin[blockIdx.y + 1];
  if (t_begin >= t_end) return;

  if (tid == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    done_cnt[0] = done_cnt[1] = 0;
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
  // Frame tile -> shared, centred, regrouped into frame pairs.
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
      xsf[((d >> 1) * NPAIR + (fr >> 1)) * 4 + (d & 1) * 2 + (fr & 1)] = v;
    }
  }
  __syncthreads();

  // log-sum-exp state carried across the slots of a multi-slot state lives in shared memory
  // (only touched by states with more than 16 components): run[q][tid] = {a.x, a.y, s.x, s.y}
  float4 *run = reinterpret_cast<float4 *>(stage0 + 2 * (size_t)stage_bytes);
  uint32_t phase = 0;                                   // bit b = parity to wait for on buffer b
  float *const out_col = sll + (f0 - f_begin) + 2 * lane;
  for (int t = t_begin; t < t_end; ++t) {
    const int b = (t - t_begin) & 1;
    unsigned char *stage = stage0 + (size_t)b * stage_bytes;
    const float4 *ps = reinterpret_cast<const float4 *>(stage) + warp * GR;
    const float *cs = reinterpret_cast<const float *>(stage + (size_t)DP * TC * sizeof(float4));
    if (!(dbg & 2) || t < t_begin + 2) mbar_wait(&full_bar[b], (phase >> b) & 1);
    phase ^= 1u << b;

    float2 acc[2][GR];                                // [pair][component] = -(log-likelihood) of 2 frames
    // One dim pair of the slot for this thread's 4 frames.  FIRST: the accumulators start at -c, which
    // enters the first FFMA2 as a scalar-broadcast addend (no accumulator initialisation moves).
    auto dim_pair = [&](int dp, auto first_tag) {
      constexpr bool FIRST = decltype(first_tag)::value;
      const float4 xa = xs[dp * NPAIR + lane];        // frames 2*lane, 2*lane+1
      const float4 xb = xs[dp * NPAIR + 32 + lane];   // frames 64+2*lane, 65+2*lane
#pragma unroll
      for (int c = 0; c < GR; ++c) {
        const float4 p = ps[dp * TC + c];             // warp-uniform address: broadcast
        float nc = 0.f;
        if constexpr (FIRST) nc = cs[warp * GR + c];
        if constexpr (F2) {
          // s and m enter as scalar-broadcast operands
          float2 t0 = __ffma2_rn(make_float2(xa.x, xa.y), make_float2(p.x, p.x), make_float2(p.z, p.z));
          float2 u0 = __ffma2_rn(make_float2(xb.x, xb.y), make_float2(p.x, p.x), make_float2(p.z, p.z));
          float2 t1 = __ffma2_rn(make_float2(xa.z, xa.w), make_float2(p.y, p.y), make_float2(p.w, p.w));
          float2 u1 = __ffma2_rn(make_float2(xb.z, xb.w), make_float2(p.y, p.y), make_float2(p.w, p.w));
          acc[0][c] = __ffma2_rn(t0, t0, FIRST ? make_float2(nc, nc) : acc[0][c]);
          acc[1][c] = __ffma2_rn(u0, u0, FIRST ? make_float2(nc, nc) : acc[1][c]);
          acc[0][c] = __ffma2_rn(t1, t1, acc[0][c]);
          acc[1][c] = __ffma2_rn(u1, u1, acc[1][c]);
        } else {
          float a0 = fmaf(xa.x, p.x, p.z), a1 = fmaf(xa.y, p.x, p.z), a2 = fmaf(xa.z, p.y, p.w), a3 = fmaf(xa.w, p.y, p.w);
          float b0 = fmaf(xb.x, p.x, p.z), b1 = fmaf(xb.y, p.x, p.z), b2 = fmaf(xb.z, p.y, p.w), b3 = fmaf(xb.w, p.y, p.w);
          acc[0][c].x = fmaf(a0, a0, FIRST ? nc : acc[0][c].x); acc[0][c].y = fmaf(a1, a1, FIRST ? nc : acc[0][c].y);
          acc[1][c].x = fmaf(b0, b0, FIRST ? nc : acc[1][c].x); acc[1][c].y = fmaf(b1, b1, FIRST ? nc : acc[1][c].y);
          acc[0][c].x = fmaf(a2, a2, acc[0][c].x); acc[0][c].y = fmaf(a3, a3, acc[0][c].y);
          acc[1][c].x = fmaf(b2, b2, acc[1][c].x); acc[1][c].y = fmaf(b3, b3, acc[1][c].y);
        }
      }
    };
    dim_pair(0, std::true_type());
#pragma unroll 2
    for (int dp = 1; dp < DP; ++dp) dim_pair(dp, std::false_type());
    const int meta = reinterpret_cast<const int *>(cs + TC)[warp];

    // This warp is done with the stage buffer; the last of the 8 warps re-arms it.
    // Hand-off of the buffer to the async proxy (WAR): every warp's reads of the stage are ordered before its count
    // (syncwarp + release fence + atomic), the last warp's count is followed by an acquire fence and the proxy fence,
    // then the bulk copy.  compute-sanitizer racecheck does not model this counter and reports the copy against the
    // reads (profiles/r02_sanitizer.txt); the first fill / refill -> read direction is the mbarrier's complete_tx.
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const int old = atomicAdd(&done_cnt[b], 1);
      if (old == NW - 1) {
        done_cnt[b] = 0;
        __threadfence_block();
        if (t + 2 < t_end && !(dbg & 2)) {
          fence_proxy_async();
          mbar_expect_tx(&full_bar[b], stage_bytes);
          tma_load_1d(stage, params + (size_t)(t + 2) * tile_floats, stage_bytes, &full_bar[b]);
        }
      }
    }
    if (dbg & 1) {   // experiment: no log-sum-exp epilogue
      float r = 0.f;
#pragma unroll
      for (int c = 0; c < GR; ++c) r += acc[0][c].x + acc[0][c].y + acc[1][c].x + acc[1][c].y;
      if (r == 1.2345f) out_col[0] = r;
      continue;
    }

    // Mixture log-sum-exp over the slot, in registers: a = min(-ll), s = sum exp(ll + a).
    const bool first = (meta & 2) != 0, last = (meta & 1) != 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float2 a = acc[q][0];
#pragma unroll
      for (int c = 1; c < GR; ++c) { a.x = fminf(a.x, acc[q][c].x); a.y = fminf(a.y, acc[q][c].y); }
      float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!first) {
        prev = run[q * NTH + tid];
        a.x = fminf(a.x, prev.x); a.y = fminf(a.y, prev.y);
      }
      const float2 al = make_float2(a.x * LOG2E, a.y * LOG2E);
      const float2 nl = make_float2(-LOG2E, -LOG2E);
      float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < GR; c += 2) {
        float2 e0 = __ffma2_rn(acc[q][c], nl, al), e1 = __ffma2_rn(acc[q][c + 1], nl, al);
        s0 = __fadd2_rn(s0, make_float2(ex2f(e0.x), ex2f(e0.y)));
        s1 = __fadd2_rn(s1, make_float2(ex2f(e1.x), ex2f(e1.y)));
      }
      float2 sum = __fadd2_rn(s0, s1);
      if (!first) {
        float2 e = __ffma2_rn(make_float2(prev.x, prev.y), nl, al);
        sum = __ffma2_rn(make_float2(prev.z, prev.w), make_float2(ex2f(e.x), ex2f(e.y)), sum);
      }
      if (!last) run[q * NTH + tid] = make_float4(a.x, a.y, sum.x, sum.y);
      if (last && meta >= 0) {
        const float2 res = make_float2(fmaf(lg2f(sum.x), LN2, -a.x), fmaf(lg2f(sum.y), LN2, -a.y));
        *reinterpret_cast<float2 *>(out_col + (int64_t)(meta >> 2) * ldF + q * 64) = res;
      }
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
             int S, double *__restrict__ lin, int64_t ldF, double floor_at, const int *__restrict__ g2c,
             const unsigned char *__restrict__ csel, const double *__restrict__ clik,
             const int *__restrict__ g_tr = nullptr, const void *__restrict__ tfeats = nullptr, int64_t t_stride = 0)
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
      // regression-class CMLLR (AdaptedGaussian, aku/ModelModules.hh:161-171): this Gaussian sees its class's A f + b
      const int tr = g_tr ? g_tr[g] : -1;
      if (tr < 0) {
        for (int d = 0; d < D; ++d) {
          const double m = mu[d], p = pr[d];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double dd = __dsub_rn(xs[d * 128 + lane + 32 * i], m);
            ll[i] = __dadd_rn(ll[i], __dmul_rn(__dmul_rn(dd, dd), p));
          }
        }
      } else {
        for (int d = 0; d < D; ++d) {
          const double m = mu[d], p = pr[d];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int64_t gf = f0 + lane + 32 * i;
            double xv = 0.0;
            if (gf < f_end)
              xv = feats_f64 ? reinterpret_cast<const double *>(tfeats)[(int64_t)tr * t_stride + gf * D + d]
                             : (double)reinterpret_cast<const float *>(tfeats)[(int64_t)tr * t_stride + gf * D + d];
            double dd = __dsub_rn(xv, m);
            ll[i] = __dadd_rn(ll[i], __dmul_rn(__dmul_rn(dd, dd), p));
          }
        }
      }
      const double c = cst[g];
      // Gaussian clustering (PDFPool::precompute_likelihoods, aku/Distributions.cc:2685-2722): a Gaussian of a cluster
      // that was not among the best ones takes its cluster centre's likelihood -- unless that is not > 0, in which
      // case PDFPool::compute_likelihood (:2637-2644) evaluates the Gaussian after all
      const int cl = g2c ? g2c[g] : -1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double v = __dadd_rn(__dmul_rn(ll[i], -0.5), c);
        double e = exp(v);
        if (cl >= 0) {
          const int64_t at = (int64_t)cl * ldF + (f0 - f_begin) + lane + 32 * i;
          if (!csel[at]) { const double ce = clik[at]; if (ce > 0) e = ce; }
        }
        l[i] = __dadd_rn(l[i], __dmul_rn(w, e));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double v = l[i];
      if (v < floor_at) v = floor_at;
      lin[(int64_t)s * ldF + (f0 - f_begin) + lane + 32 * i] = v;
    }
  }
}

// Which clusters are evaluated exactly for a frame: clusters in descending order of their centre's likelihood, taken
// while (clusters so far < min_clusters) or (Gaussians so far < min_gaussians) -- the reference pops a
// std::priority_queue (aku/Distributions.cc:2686-2708).  One CTA per frame: bitonic sort of (likelihood, index) in
// shared memory (ties go to the lower index; among equal POSITIVE likelihoods the reference's order is that of
// libstdc++'s heap, among zero likelihoods the choice does not change any result because such clusters' members are
// evaluated exactly either way), then a prefix sum of the member counts in sorted order.
__global__ void __launch_bounds__(256)
cluster_select(const double *__restrict__ clik, int64_t ldF, int C, int Cp, int64_t nf, const int *__restrict__ csize,
               int min_clusters, int min_gaussians, unsigned char *__restrict__ csel)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *key = reinterpret_cast<double *>(smem_raw);                 // [Cp]
  int *idx = reinterpret_cast<int *>(key + Cp);                       // [Cp]
  int *cum = idx + Cp;                                                // [Cp] inclusive prefix of member counts
  __shared__ int part[256];
  const int64_t f = blockIdx.x;
  if (f >= nf) return;
  for (int i = threadIdx.x; i < Cp; i += 256) {
    key[i] = i < C ? clik[(int64_t)i * ldF + f] : -1.0;               // padding sorts last (likelihoods are >= 0)
    idx[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= Cp; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < Cp; i += 256) {
        const int p = i ^ j;
        if (p > i) {
          const bool desc = (i & k) == 0;                               // final order: descending
          const double a = key[i], b = key[p];
          const int ia = idx[i], ib = idx[p];
          const bool a_first = (a > b) || (a == b && ia < ib);          // a belongs before b
          if (desc ? !a_first : a_first) { key[i] = b; key[p] = a; idx[i] = ib; idx[p] = ia; }
        }
      }
      __syncthreads();
    }
  // inclusive prefix sum of the member counts in sorted order (each thread a contiguous run)
  const int per = (Cp + 255) / 256;
  const int b0 = threadIdx.x * per, b1 = min(Cp, b0 + per);
  int run = 0;
  for (int i = b0; i < b1; i++) { run += (idx[i] < C) ? csize[idx[i]] : 0; cum[i] = run; }
  part[threadIdx.x] = run;
  __syncthreads();
  int base = 0;
  for (int t = 0; t < threadIdx.x; t++) base += part[t];
  for (int i = b0; i < b1; i++) {
    const int c = idx[i];
    if (c >= C) continue;
    const int before = base + cum[i] - csize[c];                       // Gaussians evaluated before this cluster
    csel[(int64_t)c * ldF + f] = (i < min_clusters || before < min_gaussians) ? 1 : 0;
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
  return (size_t)p.DP * NPAIR * sizeof(float4) + 2 * p.tile_floats * sizeof(float) + 2 * NTH * sizeof(float4);
}

// Frames per full wave of the fp32 scorer (chunks should be a multiple of this).
int64_t gmm_wave_frames(akugpu_ctx *ctx) { return (int64_t)ctx->sm_count * CTAS_PER_SM * TF; }
int gmm_frame_tile() { return TF; }

// Splits the component tiles into `want` ranges whose boundaries are "clean" tiles (every queue
// starts a new state there).  Cached per ysplit in the packed model.
static const int *tile_ranges(akugpu_ctx *ctx, int want, int &got)
{
  PackedF32 &p = ctx->p32;
  auto it = p.ranges.find(want);
  if (it == p.ranges.end()) {
    std::vector<int> r(1, 0);
    for (int k = 1; k < want; ++k) {
      int target = (int)((int64_t)p.n_tiles * k / want);
      while (target < p.n_tiles && !p.clean[target]) target++;
      if (target > r.back() && target < p.n_tiles) r.push_back(target);
    }
    r.push_back(p.n_tiles);
    auto buf = std::make_shared<DevBuf>();
    buf->reserve(r.size() * sizeof(int));
    AKU_CUDA(cudaMemcpyAsync(buf->p, r.data(), r.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
    it = p.ranges.emplace(want, std::make_pair((int)r.size() - 1, buf)).first;
  }
  got = it->second.first;
  return it->second.second->as<int>();
}

template <bool F2, int DPC>
static void launch_f32_t(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end,
                         float *sll, int64_t ldF)
{
  const PackedF32 &p = ctx->p32;
  const HostModel &hm = ctx->hm;
  size_t smem = gmm_f32_smem_bytes(p);
  ensure_dynamic_smem(ctx, (const void *)gmm_diag_f32<F2, DPC>, smem);
  int64_t nf = f_end - f_begin;
  int ftiles = (int)((nf + TF - 1) / TF);
  // One wave = sm_count * 2 resident CTAs.  Full chunks are sized to whole waves by the caller
  // (gmm_wave_frames); a short chunk splits the component tiles over grid.y to fill the chip.
  const int wave = ctx->sm_count * CTAS_PER_SM;
  int want = 1;
  if (ftiles < wave) want = std::min(p.n_tiles, std::max(1, wave / ftiles));
  int ysplit = 1;
  const int *ranges = tile_ranges(ctx, want, ysplit);
  dim3 grid(ftiles, ysplit);
  gmm_diag_f32<F2, DPC><<<grid, NTH, smem, ctx->stream>>>(feats, feats_f64, f_begin, f_end, hm.D, p.DP, p.params.as<float>(),
                                                     p.tile_floats, ranges, p.center.as<float>(),
                                                     p.center64.as<double>(), sll, ldF,
                                                     getenv("AKUGPU_DBG") ? atoi(getenv("AKUGPU_DBG")) : 0);
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_gmm_f32(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, float *sll,
                    int64_t ldF)
{
  const bool dp20 = ctx->p32.DP == 20;
  if (ctx->p32.packed_ffma2) {
    if (dp20) launch_f32_t<true, 20>(ctx, feats, feats_f64, f_begin, f_end, sll, ldF);
    else launch_f32_t<true, 0>(ctx, feats, feats_f64, f_begin, f_end, sll, ldF);
  } else {
    if (dp20) launch_f32_t<false, 20>(ctx, feats, feats_f64, f_begin, f_end, sll, ldF);
    else launch_f32_t<false, 0>(ctx, feats, feats_f64, f_begin, f_end, sll, ldF);
  }
}

void launch_gmm_f64(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, double *lin,
                    int64_t ldF)
{
  const HostModel &hm = ctx->hm;
  const PackedF64 &p = ctx->p64;
  size_t smem = (size_t)hm.D * 128 * sizeof(double);
  ensure_dynamic_smem(ctx, (const void *)gmm_diag_f64, smem);
  int64_t nf = f_end - f_begin;
  int ftiles = (int)((nf + 127) / 128);
  const int *g2c = nullptr;
  const unsigned char *csel = nullptr;
  const double *clik = nullptr;
  if (hm.use_clustering && hm.n_clusters > 0) {
    // centre likelihoods (unfloored), then the per-frame choice of the clusters to evaluate exactly
    const int C = hm.n_clusters;
    int Cp = 2;
    while (Cp < C) Cp <<= 1;
    const size_t sel_smem = (size_t)Cp * (sizeof(double) + 2 * sizeof(int));
    if (sel_smem > 200 * 1024) throw Error(AKUGPU_E_MODEL, fmt("too many Gaussian clusters (%d) for the selection kernel", C));
    ctx->d_clik.reserve((size_t)C * ldF * sizeof(double));
    ctx->d_csel.reserve((size_t)C * ldF);
    int ys = std::max(1, std::min((C + 7) / 8, (4 * ctx->sm_count + ftiles - 1) / ftiles));
    gmm_diag_f64<<<dim3(ftiles, ys), 256, smem, ctx->stream>>>(feats, feats_f64, f_begin, f_end, hm.D, p.c_mean.as<double>(),
                                                              p.c_prec.as<double>(), p.c_cst.as<double>(), p.c_mix_off.as<int>(),
                                                              p.c_mix_gauss.as<int>(), p.c_mix_w.as<double>(), C,
                                                              ctx->d_clik.as<double>(), ldF, -1.0, nullptr, nullptr, nullptr);
    AKU_CUDA(cudaGetLastError());
    ensure_dynamic_smem(ctx, (const void *)cluster_select, sel_smem);
    cluster_select<<<(unsigned)nf, 256, sel_smem, ctx->stream>>>(ctx->d_clik.as<double>(), ldF, C, Cp, nf, p.c_size.as<int>(),
                                                                 hm.eval_min_clusters, hm.eval_min_gaussians,
                                                                 ctx->d_csel.as<unsigned char>());
    AKU_CUDA(cudaGetLastError());
    ctx->launches += 2;
    g2c = p.g2c.as<int>();
    csel = ctx->d_csel.as<unsigned char>();
    clik = ctx->d_clik.as<double>();
  }
  int ysplit = std::max(1, std::min((hm.S + 7) / 8, (4 * ctx->sm_count + ftiles - 1) / ftiles));
  dim3 grid(ftiles, ysplit);
  gmm_diag_f64<<<grid, 256, smem, ctx->stream>>>(feats, feats_f64, f_begin, f_end, hm.D, p.mean.as<double>(),
                                                 p.prec.as<double>(), p.cst.as<double>(), p.mix_off.as<int>(),
                                                 p.mix_gauss.as<int>(), p.mix_w.as<double>(), hm.S, lin, ldF, 1e-50, g2c, csel, clik,
                                                 hm.n_tr > 0 ? p.g_tr.as<int>() : nullptr, ctx->d_adapt.p, ctx->adapt_stride);
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
