// gmm_tc.cu -- tensor-core scorer (tcgen05 + TMEM + TMA) for sm_100a.  EXPERIMENTAL throughput mode
// (akugpu_set_scorer_variant(ctx, 3)); the default fp32 path stays gmm_diag_f32.
//
// Every Gaussian log-likelihood of the path is an inner product of an expanded feature with an
// expanded parameter vector plus a per-component constant:
//   full covariance (the reference's own exponential form, aku/Distributions.cc:1437-1446, 2664-2672):
//       ll = < [f ; vec(f f^T)] , [P mu ; -1/2 vec(P)] > + normalizer + constant        (length D(D+3)/2)
//   diagonal (aku/Distributions.cc:1041-1062 expanded):
//       ll = < [f^2 ; f] , [-1/2 p ; p mu] > - 1/2 sum p mu^2 + constant                 (length 2D)
//   (a per-dimension [f^2, f, 1] layout that keeps the accumulator near the final value was tried and
//    measured WORSE: the tensor-core accumulate error grows with the number of K steps, not with the
//    partial-sum magnitude)
// so a frame tile x component tile of log-likelihoods is a GEMM.  bf16 tensor cores with one pass
// are off by orders of magnitude; here both operands are split into three
// bf16 terms (x = x1 + x2 + x3) and the six products with i+j <= 4 are laid side by side along K:
//       A' = [A1 | A2 A3 A1 A2 A1] ,  B' = [B1 | B1 B1 B2 B2 B3]      (K' = 6 K, fp32 accumulation;
//       the leading block and the five corrections accumulate into separate TMEM accumulators)
// which gives ~2^-24 relative operand error with plain kind::f16 MMAs.  Features and means are centred
// first so that the cancellation between the quadratic and linear terms stays small.
//
// Kernel (one CTA = 128 frames, loops over 128-component tiles):
//   warp 0    TMA producer: cp.async.bulk.tensor.2d of A' and B' k-blocks (128 x 64 bf16, SWIZZLE_128B)
//             into a 4-stage ring, completion on mbarriers
//   warp 1    MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 N128 K16),
//             accumulators in TMEM (main + correction, 2 x 128 columns), tcgen05.commit frees smem stages
//             and publishes finished accumulators
//   warps 4-7 epilogue: tcgen05.ld 32x32b.x16 (one slot of 16 components per load, thread = frame row),
//             add the component constants, mixture log-sum-exp in registers (carried across the slots
//             of a big state), coalesced stores of state log-likelihoods sll[state][frame]
#include "ctx.hpp"
#include "kernels.hpp"
#include <cuda.h>
#include <cuda_bf16.h>
#include "lna_common.cuh"
#include "tc_common.cuh"
#include <math.h>

namespace akugpu {

namespace tc {
constexpr int BM = 128, BN = 128, BK = 64;        // tile: frames x components x k (bf16 elements)
constexpr int STAGES = 3;                          // 3 x 32 KB: two CTAs (and their 2 x 256 TMEM columns) per SM
constexpr int GR = 16;                             // components per slot
constexpr int SLOTS = BN / GR;
constexpr uint32_t STAGE_BYTES = (BM + BN) * BK * 2;

constexpr uint32_t IDESC = umma_idesc(BM, BN, true);
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  umma_f16(tmem_d, da, db, IDESC, accumulate);
}
}  // namespace tc

// ------------------------------------------------------------------------------------------------
// Expanded, centred, bf16x3-split features:  A'[frame][K'] , K' = 6*L padded to a multiple of 64.
__global__ void tc_expand_feats(const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t nf, int64_t rows,
                                int D, int L, int Lm, int Kp, int full, const double *__restrict__ center,
                                __nv_bfloat16 *__restrict__ A)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const int64_t fr = i / L;
  const int l = (int)(i - fr * L);
  float v = 0.f;
  if (fr < nf) {
    auto x = [&](int d) -> double {
      double t = feats_f64 ? reinterpret_cast<const double *>(feats)[(f_begin + fr) * D + d]
                           : (double)reinterpret_cast<const float *>(feats)[(f_begin + fr) * D + d];
      return t - center[d];
    };
    if (full) {
      if (l < D) v = (float)x(l);
      else {
        int pos = l - D, r = 0;
        while ((r + 1) * (r + 2) / 2 <= pos) r++;
        const int c = pos - r * (r + 1) / 2;
        const double m = x(r) * x(c);
        v = (float)((r == c) ? m : sqrt(2.0) * m);
      }
    } else {
      v = (l < D) ? (float)(x(l) * x(l)) : (float)x(l - D);
    }
  }
  const __nv_bfloat16 a1 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(a1);
  const __nv_bfloat16 a2 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 a3 = __float2bfloat16_rn(r1 - __bfloat162float(a2));
  __nv_bfloat16 *row = A + fr * Kp;
  // [ A1 | pad to Lm ][ A2 A3 A1 A2 A1 | pad to Kp ]   against   [ B1 ][ B1 B1 B2 B2 B3 ]
  row[l] = a1;
  __nv_bfloat16 *cr = row + Lm;
  cr[l] = a2; cr[L + l] = a3; cr[2 * L + l] = a1; cr[3 * L + l] = a2; cr[4 * L + l] = a1;
  if (l == 0) {
    for (int k = L; k < Lm; k++) row[k] = __float2bfloat16_rn(0.f);
    for (int k = Lm + 5 * L; k < Kp; k++) row[k] = __float2bfloat16_rn(0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// ARES = true : the frame tile's whole A' row block (kblocks x 16 KB) stays resident in shared memory and only B'
//               streams through the ring; one CTA per SM, 512 TMEM columns = two (main, correction) accumulator
//               pairs, so the epilogue of tile n overlaps the MMAs of tile n+1.  Used when A' fits (diagonal pools).
// ARES = false: A' and B' both stream (full-covariance pools, K' = 4928); two CTAs per SM with one accumulator
//               pair each.
template <bool ARES>
__global__ void __launch_bounds__(384, ARES ? 1 : 2)
gmm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int kblocks, int kb_main,
              const int *__restrict__ range_begin, const float *__restrict__ bias, const int *__restrict__ meta,
              float *__restrict__ sll, int64_t ldF, float2 *__restrict__ norm)
{
  using namespace tc;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2], a_full;
  __shared__ uint32_t tmem_base_smem;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  constexpr int TMEM_COLS = ARES ? 512 : 256;
  unsigned char *ring = ARES ? smem + (size_t)kblocks * A_BYTES : smem;      // B' ring (ARES) or A'+B' ring
  constexpr uint32_t RING_STAGE = ARES ? B_BYTES : STAGE_BYTES;
  __shared__ __align__(16) float sbias[2][BN];
  __shared__ int smeta[2][SLOTS];
  __shared__ float sM[8][32];
  __shared__ double sR[8][32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n_begin = range_begin[blockIdx.y], n_end = range_begin[blockIdx.y + 1];

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    mbar_init(&a_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: per accumulator pair 256 columns = main [0,128) + correction [128,256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0 && elect_one()) {
    // ===== TMA producer (elect.sync: see tc_common.cuh) =====
    int it = 0;
    if (ARES) {
      mbar_expect_tx(&a_full, (uint32_t)kblocks * A_BYTES);
      for (int kb = 0; kb < kblocks; kb++) tma_load_2d(smem + (size_t)kb * A_BYTES, &mapA, kb * BK, m0, &a_full);
    }
    for (int n = n_begin; n < n_end; n++) {
      for (int kb = 0; kb < kblocks; kb++, it++) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
        mbar_expect_tx(&full_bar[s], RING_STAGE);
        unsigned char *dst = ring + (size_t)s * RING_STAGE;
        if (!ARES) { tma_load_2d(dst, &mapA, kb * BK, m0, &full_bar[s]); dst += A_BYTES; }
        tma_load_2d(dst, &mapB, kb * BK, n * BN, &full_bar[s]);
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===== MMA issuer =====
    int it = 0;
    if (ARES) mbar_wait(&a_full, 0);
    for (int n = n_begin; n < n_end; n++) {
      const int a = ARES ? ((n - n_begin) & 1) : 0, use = ARES ? ((n - n_begin) >> 1) : (n - n_begin);
      if (use > 0) mbar_wait(&tmem_empty[a], (use - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < kblocks; kb++, it++) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ring_addr = smem_u32(ring + (size_t)s * RING_STAGE);
        const uint32_t a_addr = ARES ? smem_u32(smem + (size_t)kb * A_BYTES) : ring_addr;
        const uint32_t b_addr = ARES ? ring_addr : ring_addr + A_BYTES;
        // the leading bf16 terms (A1.B1) go to the main accumulator, the five correction products to a
        // second one: the tensor core's accumulate error scales with the accumulator's magnitude times the
        // number of K steps, and the corrections are 2^-8 smaller
        const bool main_blk = kb < kb_main;
        const uint32_t tmem_d = tmem_base + a * 2 * BN + (main_blk ? 0 : BN);
        const int kfirst = main_blk ? 0 : kb_main;
#pragma unroll
        for (int k = 0; k < BK / 16; k++)
          umma_bf16(tmem_d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), ((kb - kfirst) | k) ? 1u : 0u);
        umma_commit(&empty_bar[s]);          // smem stage reusable once these MMAs have read it
      }
      umma_commit(&tmem_full[a]);            // accumulator complete
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = one frame row of the tile =====
    // 8 epilogue warps: warp w may only touch TMEM lanes 32*(w%4)..+31, so two warps share each lane quarter
    // and split the tile's 8 slots between them (a state's slots never straddle a half tile, see model_pack_tc)
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;                         // which 4 slots of the tile
    const int et = threadIdx.x - 128;                         // 0..255 among the epilogue threads
    const int64_t frame = (int64_t)m0 + q * 32 + lane;
    float run_a = 0.f, run_s = 0.f;
    // fused first pass of the LNA epilogue (aku/phone_probs.cc:227-232): running maximum of the float-cast
    // state likelihoods of this frame and the sum of all the other terms relative to it
    float nMx = -INFINITY, nR = 0.f;   // fp32 is enough: the sum is relative to the maximum (error ~1e-6 of lognorm)
    // component constants / slot table are fetched one tile ahead into a register (software pipelining): the
    // global-load latency hides behind the current tile's epilogue instead of sitting in front of it
    auto fetch = [&](int n) -> uint32_t {
      if (n >= n_end) return 0u;
      if (et < BN) return __float_as_uint(__ldg(bias + (size_t)n * BN + et));
      if (et < BN + SLOTS) return (uint32_t)__ldg(meta + n * SLOTS + et - BN);
      return 0u;
    };
    uint32_t pre = fetch(n_begin);
    for (int n = n_begin; n < n_end; n++) {
      const int a = ARES ? ((n - n_begin) & 1) : 0, use = ARES ? ((n - n_begin) >> 1) : (n - n_begin);
      // component constants and slot table of this tile -> shared, issued BEFORE waiting for the
      // accumulator so that the global-load latency hides under the tile's MMAs
      const int sb = (n - n_begin) & 1;                          // constants are double buffered by tile parity
      if (et < BN) sbias[sb][et] = __uint_as_float(pre);
      else if (et < BN + SLOTS) smeta[sb][et - BN] = (int)pre;
      pre = fetch(n + 1);
      mbar_wait(&tmem_full[a], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");          // the 8 epilogue warps only
#pragma unroll 2
      for (int sl = half * (SLOTS / 2); sl < (half + 1) * (SLOTS / 2); sl++) {
        uint32_t r[16], rc[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * 2 * BN + sl * GR;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            "tcgen05.wait::ld.sync.aligned;\n"
            : "=r"(rc[0]), "=r"(rc[1]), "=r"(rc[2]), "=r"(rc[3]), "=r"(rc[4]), "=r"(rc[5]), "=r"(rc[6]), "=r"(rc[7]), "=r"(rc[8]),
              "=r"(rc[9]), "=r"(rc[10]), "=r"(rc[11]), "=r"(rc[12]), "=r"(rc[13]), "=r"(rc[14]), "=r"(rc[15])
            : "r"(taddr + BN)
            : "memory");
        const int mt = smeta[sb][sl];
        const float4 *bp = reinterpret_cast<const float4 *>(&sbias[sb][sl * GR]);
        float v[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
          const float4 b = bp[c4];
          v[4 * c4 + 0] = (__uint_as_float(r[4 * c4 + 0]) + __uint_as_float(rc[4 * c4 + 0])) + b.x;
          v[4 * c4 + 1] = (__uint_as_float(r[4 * c4 + 1]) + __uint_as_float(rc[4 * c4 + 1])) + b.y;
          v[4 * c4 + 2] = (__uint_as_float(r[4 * c4 + 2]) + __uint_as_float(rc[4 * c4 + 2])) + b.z;
          v[4 * c4 + 3] = (__uint_as_float(r[4 * c4 + 3]) + __uint_as_float(rc[4 * c4 + 3])) + b.w;
        }
        const bool first = (mt & 2) != 0, last = (mt & 1) != 0;
        float mx = v[0];
#pragma unroll
        for (int c = 1; c < 16; c++) mx = fmaxf(mx, v[c]);
        if (!first) mx = fmaxf(mx, run_a);
        const float ml = mx * LOG2E;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          s0 += ex2f(fmaf(v[c], LOG2E, -ml));
          s1 += ex2f(fmaf(v[c + 1], LOG2E, -ml));
        }
        float sum = s0 + s1;
        if (!first) sum = fmaf(run_s, ex2f(fmaf(run_a, LOG2E, -ml)), sum);
        run_a = mx;
        run_s = sum;
        if (last && mt >= 0) {
          const float res = fmaf(lg2f(sum), LN2, mx);
          sll[(int64_t)(mt >> 2) * ldF + frame] = res;
          if (norm) {
            const float Lc = log_of_float_cast(res);
            if (Lc > nMx) {
              nR = (nMx == -INFINITY) ? 0.f : (nR + 1.f) * ex2f((nMx - Lc) * LOG2E);
              nMx = Lc;
            } else {
              nR += ex2f((Lc - nMx) * LOG2E);      // exp2(-inf) = 0 covers flushed states
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[a]);
    }
    if (norm) {   // the two warps of a lane quarter merge their halves; one float2 per frame
      sM[warp - 4][lane] = nMx;
      sR[warp - 4][lane] = (double)nR;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0) {
        const float oM = sM[q + 4][lane];
        const double oR = sR[q + 4][lane];
        float gM = nMx;
        double Rt = (double)nR;
        if (oM > nMx) { gM = oM; Rt = oR + ((nMx == -INFINITY) ? 0.0 : (1.0 + (double)nR) * exp((double)(nMx - oM))); }
        else if (oM != -INFINITY) Rt = (double)nR + (1.0 + oR) * exp((double)(oM - nMx));
        norm[frame] = make_float2(gM, (float)log1p(Rt));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void tc_make_map(CUtensorMap *map, void *base, uint64_t rows, uint64_t cols, bool fp16, int box_cols, int box_rows)
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void *p = nullptr;
    AKU_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw Error(AKUGPU_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {cols, rows};                  // innermost first
  cuuint64_t strides[1] = {cols * 2};                 // bytes, dims 1..
  if (box_cols != tc::BK && box_cols != tc::BK / 2) throw Error(AKUGPU_E_ARG, "tc_make_map: box of 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B) columns");
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols == tc::BK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(AKUGPU_E_CUDA, fmt("cuTensorMapEncodeTiled failed (%d)", (int)r));
}

static inline uint16_t bf16_bits(float v)   // round to nearest even
{
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_val(uint16_t b) { uint32_t u = (uint32_t)b << 16; float v; memcpy(&v, &u, 4); return v; }

// Slots (16 component positions of one state) in state order; tiles of 8 slots.  The epilogue warps of a TMEM lane
// quarter each own `group` consecutive slots of a tile: padding slots (state -1) keep a state from straddling a group.
void tc_build_slots(const HostModel &hm, int group, std::vector<int> &slot_state, std::vector<int> &slot_k0, std::vector<int> &slot_flags,
                    const std::vector<char> *skip_state)
{
  const int HALF = group;        // slots handled by one epilogue warp: no state straddles such a group
  slot_state.clear(); slot_k0.clear(); slot_flags.clear();
  for (int s = 0; s < hm.S; s++) {
    if (skip_state && (*skip_state)[s]) continue;
    const int K = hm.mix_off[s + 1] - hm.mix_off[s];
    const int ns = std::max(1, (K + tc::GR - 1) / tc::GR);
    if (ns > HALF) throw Error(AKUGPU_E_MODEL, fmt("the tensor-core scorer handles at most %d components per state (state %d has %d)", HALF * tc::GR, s, K));
    while ((int)slot_state.size() % HALF + ns > HALF) { slot_state.push_back(-1); slot_k0.push_back(0); slot_flags.push_back(3); }
    for (int i = 0; i < ns; i++) { slot_state.push_back(s); slot_k0.push_back(i * tc::GR); slot_flags.push_back(((i == 0) << 1) | (i == ns - 1)); }
  }
}

// Expanded parameters of every Gaussian for CENTRED features (double): ll = <expand(x - c), theta_g> + gconst_g.
//   diagonal (L = 2D):        theta = [-p/2 ; p m],  gconst = log sqrt(prod p) - 1/2 sum p m^2        (m = mu - c)
//   full (L = D(D+3)/2):      theta = [P m ; -1/2 vec(P)] (lower triangle, off-diagonals x sqrt 2; the reference's own
//                             exponential form, aku/Distributions.cc:1530-1547), gconst = log sqrt(det P) - 1/2 m'Pm
// The expanded form adds terms of order q_g = 1/2 sum_d p_gd (mu_gd - c_d)^2 that cancel against the constant, and its
// fp32 accumulation error is ~4e-7 * q_g (measured, scripts/shape_check.py).  The feature centre c is the
// precision-weighted mean of the means per dimension (diagonal of the covariance for full Gaussians): it keeps q small
// for the bulk of the Gaussians -- the outliers are what the caller hands to the direct-form kernel (a minimax centre
// lowers the worst q but pushes many more states over the limit).  Returns max_g q_g, and q per Gaussian on request.
double tc_expanded_params(const HostModel &hm, bool full, int L, std::vector<double> &cen, std::vector<double> &theta,
                          std::vector<double> &gconst, std::vector<double> *q_of_gauss)
{
  const int G = hm.G, D = hm.D;
  cen.assign(D, 0.0);
  for (int d = 0; d < D; d++) {
    double sw = 0, swm = 0, sm = 0;
    for (int g = 0; g < G; g++) {
      const double cv = full ? hm.full_cov[((size_t)hm.full_index[g] * D + d) * D + d] : hm.cov[(size_t)g * D + d];
      const double pr = cv > 0 ? 1 / cv : 0;
      sw += pr;
      swm += pr * hm.mean[(size_t)g * D + d];
      sm += hm.mean[(size_t)g * D + d];
    }
    cen[d] = sw > 0 && std::isfinite(swm / sw) ? swm / sw : (G ? sm / G : 0);
  }
  double q_max = 0;
  if (q_of_gauss) q_of_gauss->assign(G, 0.0);
  // per-Gaussian expanded parameters for CENTRED features (double)
  theta.assign((size_t)G * L, 0.0);
  gconst.assign(G, 0.0);
  std::vector<double> P, chol, Pm(D), mu(D);
  for (int g = 0; g < G; g++) {
    for (int d = 0; d < D; d++) mu[d] = hm.mean[(size_t)g * D + d] - cen[d];
    double *th = &theta[(size_t)g * L];
    if (!full) {
      double c = 1, q = 0;
      for (int d = 0; d < D; d++) {
        const double cv = hm.cov[(size_t)g * D + d], pr = cv > 0 ? 1 / cv : 0;
        c *= pr;
        th[d] = -0.5 * pr;
        th[D + d] = pr * mu[d];
        q += pr * mu[d] * mu[d];
      }
      if (c > 0) c = log(sqrt(c));
      gconst[g] = c - 0.5 * q;
      q_max = std::max(q_max, 0.5 * q);
      if (q_of_gauss) (*q_of_gauss)[g] = 0.5 * q;
    } else {
      std::vector<double> cov(hm.full_cov.begin() + (size_t)hm.full_index[g] * D * D,
                              hm.full_cov.begin() + (size_t)(hm.full_index[g] + 1) * D * D);
      if (!host_cholesky(cov, D, chol)) { gconst[g] = 0; continue; }
      host_lu_inverse(cov, D, P);
      if (!host_cholesky(P, D, chol)) { gconst[g] = 0; continue; }
      double det = 1;
      for (int i = 0; i < D; i++) det *= chol[(size_t)i * D + i];
      det *= det;
      double dot = 0;
      for (int i = 0; i < D; i++) { double s = 0; for (int j = 0; j < D; j++) s += P[(size_t)i * D + j] * mu[j]; Pm[i] = s; }
      for (int i = 0; i < D; i++) dot += Pm[i] * mu[i];
      for (int i = 0; i < D; i++) th[i] = Pm[i];
      int pos = D;
      for (int i = 0; i < D; i++)
        for (int j = 0; j <= i; j++, pos++) th[pos] = -0.5 * ((i == j) ? P[(size_t)i * D + j] : sqrt(2.0) * P[(size_t)i * D + j]);
      gconst[g] = log(sqrt(det)) - 0.5 * dot;
      q_max = std::max(q_max, 0.5 * dot);
      if (q_of_gauss) (*q_of_gauss)[g] = 0.5 * dot;
    }
  }
  return q_max;
}

// Builds B' (slot-ordered components x K'), the per-component constants and the slot table.
void model_pack_tc(akugpu_ctx *ctx)
{
  const HostModel &hm = ctx->hm;
  PackedTC &p = ctx->ptc;
  const int S = hm.S, G = hm.G, D = hm.D;
  p.full = hm.n_full > 0;
  if (p.full && hm.n_full != G) throw Error(AKUGPU_E_MODEL, "the tensor-core scorer needs an all-diagonal or an all-full pool");
  p.L = p.full ? D * (D + 3) / 2 : 2 * D;
  p.Lm = (p.L + tc::BK - 1) / tc::BK * tc::BK;                       // leading terms, padded to whole k-blocks
  p.Kp = p.Lm + (5 * p.L + tc::BK - 1) / tc::BK * tc::BK;          // + the five correction products
  std::vector<double> cen, theta, gconst;
  p.q_max = tc_expanded_params(hm, p.full, p.L, cen, theta, gconst, nullptr);
  std::vector<int> slot_state, slot_k0, slot_flags;
  tc_build_slots(hm, tc::SLOTS / 2, slot_state, slot_k0, slot_flags, nullptr);
  const int n_slots = (int)slot_state.size();
  p.n_tiles = (n_slots + tc::SLOTS - 1) / tc::SLOTS;
  const size_t rows = (size_t)p.n_tiles * tc::BN;
  std::vector<uint16_t> B(rows * p.Kp, 0);
  std::vector<float> bias(rows, -1.0e30f);
  std::vector<int32_t> meta((size_t)p.n_tiles * tc::SLOTS, -1);
  p.clean.assign(p.n_tiles, 1);
  for (int sl = 0; sl < n_slots; sl++) {
    const int s = slot_state[sl];
    if (s < 0) continue;                                   // padding slot: meta stays -1, bias -1e30
    meta[sl] = (s << 2) | slot_flags[sl];
    if (sl % tc::SLOTS == 0 && !(slot_flags[sl] & 2)) p.clean[sl / tc::SLOTS] = 0;
    const int K = hm.mix_off[s + 1] - hm.mix_off[s];
    for (int j = 0; j < tc::GR && slot_k0[sl] + j < K; j++) {
      const int k = hm.mix_off[s] + slot_k0[sl] + j, g = hm.mix_gauss[k];
      const double w = hm.mix_w[k];
      const size_t row = (size_t)sl * tc::GR + j;
      bias[row] = w > 0 ? (float)(log(w) + gconst[g]) : -1.0e30f;
      uint16_t *br = &B[row * p.Kp];
      for (int l = 0; l < p.L; l++) {
        const float v = (float)theta[(size_t)g * p.L + l];
        const uint16_t b1 = bf16_bits(v);
        const float r1 = v - bf16_val(b1);
        const uint16_t b2 = bf16_bits(r1);
        const uint16_t b3 = bf16_bits(r1 - bf16_val(b2));
        br[l] = b1;
        uint16_t *bc = br + p.Lm;
        bc[l] = b1; bc[p.L + l] = b1; bc[2 * p.L + l] = b2; bc[3 * p.L + l] = b2; bc[4 * p.L + l] = b3;
      }
    }
  }
  auto up = [&](DevBuf &b, const void *src, size_t bytes) {
    b.reserve(std::max<size_t>(bytes, 16));
    AKU_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  };
  up(p.B, B.data(), B.size() * 2);
  up(p.bias, bias.data(), bias.size() * 4);
  up(p.meta, meta.data(), meta.size() * 4);
  up(p.center, cen.data(), cen.size() * 8);
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  p.ranges.clear();
  p.ready = true;
}

// Splits the component tiles into `want` contiguous ranges that start on "clean" tiles (a tile that begins a new
// state), for launches with fewer frame tiles than SMs.  Cached per split count on the device.
const int *tc_tile_ranges(akugpu_ctx *ctx, int n_tiles, const std::vector<char> &clean,
                          std::map<int, std::pair<int, std::shared_ptr<DevBuf>>> &ranges, int want, int &got)
{
  auto it = ranges.find(want);
  if (it == ranges.end()) {
    std::vector<int> r(1, 0);
    for (int k = 1; k < want; ++k) {
      int target = (int)((int64_t)n_tiles * k / want);
      while (target < n_tiles && !clean[target]) target++;
      if (target > r.back() && target < n_tiles) r.push_back(target);
    }
    r.push_back(n_tiles);
    auto buf = std::make_shared<DevBuf>();
    buf->reserve(r.size() * sizeof(int));
    AKU_CUDA(cudaMemcpyAsync(buf->p, r.data(), r.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
    it = ranges.emplace(want, std::make_pair((int)r.size() - 1, buf)).first;
  }
  got = it->second.first;
  return it->second.second->as<int>();
}

// A' resident when its row block plus the B' ring fits the 227 KB of shared memory.
static bool gmm_tc_a_resident(const PackedTC &p)
{
  // Measured on B200 (profiles/r01_tc_*): streaming both operands with two CTAs per SM (6 x 32 KB in flight)
  // beats the resident-A variant, whose 3-deep 16 KB B' ring cannot cover the L2 latency.  Resident A stays
  // available behind AKUGPU_TC_ARES=1 for experiments.
  static const char *force = getenv("AKUGPU_TC_ARES");
  if (!force) return false;
  if (force) return atoi(force) != 0 && (size_t)(p.Kp / tc::BK) * tc::BM * tc::BK * 2 + (size_t)tc::STAGES * tc::BN * tc::BK * 2 + 2048 <= 227 * 1024;
  return (size_t)(p.Kp / tc::BK) * tc::BM * tc::BK * 2 + (size_t)tc::STAGES * tc::BN * tc::BK * 2 + 2048 <= 227 * 1024;
}
int64_t gmm_tc_wave_frames(akugpu_ctx *ctx) { return (int64_t)ctx->sm_count * (gmm_tc_a_resident(ctx->ptc) ? 1 : 2) * tc::BM; }

bool launch_gmm_tc(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, float *sll, int64_t ldF,
                   float2 *norm)
{
  PackedTC &p = ctx->ptc;
  const HostModel &hm = ctx->hm;
  const int64_t nf = f_end - f_begin;
  if (nf <= 0) return false;
  const int64_t rows = (nf + tc::BM - 1) / tc::BM * tc::BM;
  ctx->d_fe[4].reserve((size_t)rows * p.Kp * 2);
  __nv_bfloat16 *A = ctx->d_fe[4].as<__nv_bfloat16>();
  const int64_t ne = rows * p.L;
  {
    StageScope sc(ctx, 0);   // feature expansion is accounted to the front-end stage
    tc_expand_feats<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(feats, feats_f64, f_begin, nf, rows, hm.D, p.L, p.Lm,
                                                                         p.Kp, p.full ? 1 : 0, p.center.as<double>(), A);
    ctx->launches++;
  }
  CUtensorMap mapA, mapB;
  tc_make_map(&mapA, A, (uint64_t)rows, (uint64_t)p.Kp, false);
  tc_make_map(&mapB, p.B.p, (uint64_t)p.n_tiles * tc::BN, (uint64_t)p.Kp, false);
  const int ftiles = (int)(rows / tc::BM);
  const int kblocks = p.Kp / tc::BK;
  const bool ares = gmm_tc_a_resident(p);
  const int per_sm = ares ? 1 : 2;
  int want = 1;
  if (ftiles < per_sm * ctx->sm_count) want = std::min(p.n_tiles, std::max(1, per_sm * ctx->sm_count / ftiles));
  int ysplit = 1;
  const int *ranges = tc_tile_ranges(ctx, p.n_tiles, p.clean, p.ranges, want, ysplit);
  dim3 grid(ftiles, ysplit);
  if (ysplit != 1) norm = nullptr;        // a frame's states are spread over several CTAs: the LNA kernel does both passes
  StageScope sc(ctx, 1);
  if (ares) {
    const size_t smem = (size_t)kblocks * tc::BM * tc::BK * 2 + (size_t)tc::STAGES * tc::BN * tc::BK * 2 + 1024;
    ensure_dynamic_smem(ctx, (const void *)gmm_tc_kernel<true>, smem);
    gmm_tc_kernel<true><<<grid, 384, smem, ctx->stream>>>(mapA, mapB, kblocks, p.Lm / tc::BK, ranges, p.bias.as<float>(),
                                                          p.meta.as<int>(), sll, ldF, norm);
  } else {
    const size_t smem = (size_t)tc::STAGES * tc::STAGE_BYTES + 1024;
    ensure_dynamic_smem(ctx, (const void *)gmm_tc_kernel<false>, smem);
    gmm_tc_kernel<false><<<grid, 384, smem, ctx->stream>>>(mapA, mapB, kblocks, p.Lm / tc::BK, ranges, p.bias.as<float>(),
                                                           p.meta.as<int>(), sll, ldF, norm);
  }
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
  return norm != nullptr;
}

}  // namespace akugpu
