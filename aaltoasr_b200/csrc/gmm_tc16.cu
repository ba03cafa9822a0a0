// gmm_tc16.cu -- default throughput scorer: fp16 hi/lo-split tensor-core GEMM (tcgen05 + TMEM + TMA).  Diagonal pools:
// the frame tile's expanded features are built in the kernel and stay resident in shared memory (gmm_tc16_kernel<NCH>);
// full-covariance pools (K = D(D+3)/2 + 2 terms): the same kernel streaming A' next to B' (gmm_tc16_kernel<0>).
// Which states it serves is decided at pack time from the conditioning of the expanded form (model_pack_tc16): states with
// a component too sharp / too far from the feature centre go to the direct-form FP32-pipe kernel in the same pass.
//
// Reference arithmetic (aku/Distributions.cc:1041-1062, 2079-2086): for every Gaussian of the pool
//     ll = -1/2 sum_d p_d (f_d - mu_d)^2 + log sqrt(prod p_d),        state likelihood = sum_k w_k exp(ll_k).
// Expanded around a feature centre c (x = f - c, m = mu - c):
//     log w + ll = < [x^2 ; x ; 1 ; 1] , [-p/2 ; p m ; b_coarse ; b_fine] >,     b = log w + const - 1/2 sum p m^2
// so a frame tile x component tile of log-likelihoods is a GEMM with K = 2D+2 (80 for D = 39, exactly five K16 steps).
//
// Precision.  One fp16 pass is useless.  Each operand is split in two fp16 terms,
//     a = ah + al * 2^-11,   b = bh + bl * 2^-11      (al, bl stored scaled by 2^11, so nothing is denormal),
// and three products are accumulated in fp32:  MAIN = ah.bh  and  CORR = ah.bl + al.bh  in a second TMEM accumulator
// (the tensor core's accumulate error scales with accumulator magnitude x K steps, and CORR is 2^-11 smaller); the
// epilogue forms MAIN + 2^-11 CORR.  The dropped al.bl term is <= 2^-24 relative, the same order as the terms the
// previous bf16x3 kernel (gmm_tc.cu, six products, K' = 576) dropped -- at 15 K16 steps per tile instead of 36.
// fp16 range: term k is scaled by a power of two, A'_k = A_k 2^e_k, B'_k = B_k 2^-e_k, chosen from the model so that
// max|B'_k| is in [2^11, 2^12); a frame whose A'_k would leave the fp16 range raises a flag and the call is redone
// with the bf16x3 kernel (never seen on real features: it needs a single term of order 2^27).
// The component constant rides in two extra K terms (A = 1): b_coarse = fp16(b) exactly, b_fine = b - b_coarse.
//
// Kernel: one CTA per SM, 128 frames, loops over 128-component tiles.
//   epilogue warps (16) first stage the tile's features, centre them, expand, scale, split and store A' = [Ah | Al]
//             into shared memory in the canonical K-major SWIZZLE_128B layout (what TMA would have written)
//   warp 0    TMA producer (one elected lane): a component tile of B' = 3 k-blocks (128 components x 64 fp16,
//             SWIZZLE_128B) per ring slot, 3 slots, one mbarrier per slot
//   warp 1    MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16 M128 N128 K16, per tile
//             CORR = sum_j Ah_j.Bl_j + Al_j.Bh_j, MAIN = sum_j Ah_j.Bh_j (A' is never duplicated, B' is K = 160 wide);
//             descriptors are compile-time offsets from two bases, so the 15 MMAs issue back to back;
//             accumulators double buffered in TMEM (2 x (128 + 128) columns)
//   warps 4-19 epilogue in two groups of 8 that take alternate tiles (group = accumulator set): two warps per TMEM
//             lane quarter and group, four slots (4 x 16 components) each; tcgen05.ld 32x32b.x16, FFMA2/FADD2-packed
//             mixture log-sum-exp in registers, state log-likelihoods sll[state][frame], and the first pass of the LNA
//             normalisation fused in
// What bounds it (ncu, profiles/): the 16 exp2 per (frame, slot) keep the XU pipe 72 % busy, the tensor pipe is 57 %
// busy (15 MMAs x 67 clk per tile), issue slots 52 %.  History of the issue path, all measured: a `lane == 0` branch
// made nvcc wrap every tcgen05.mma in an elect/uniformisation loop and the ring index used runtime divisions: ~200 clk
// per MMA issue (that, not shared-memory or L2 bandwidth, bound the first version and the bf16x3 kernel); elect.sync +
// compile-time descriptor offsets brought it to the pipe's rate.  All 16 epilogue warps on the same tile left the XU
// pipe idle during every wait + TMEM-load phase; two groups on alternate tiles overlap them.
// Shared-memory traffic per tile: 15 x (4 KB A + 4 KB B) operand reads + 48 KB TMA writes = 168 KB, against 576 KB
// for the streaming bf16x3 kernel.
#include "ctx.hpp"
#include "kernels.hpp"
#include <cuda.h>
#include <cuda_fp16.h>
#include "lna_common.cuh"
#include "tc_common.cuh"
#include <math.h>
#include <string.h>

namespace akugpu {

namespace tc16 {
constexpr int BM = 128, BN = 128, BK = 64;        // tile: frames x components; k-block = 64 fp16 = one 128 B swizzle row
constexpr int GR = 16, SLOTS = BN / GR;
constexpr int EPI_WARPS = 16, EPI_THREADS = EPI_WARPS * 32, THREADS = 128 + EPI_THREADS;
constexpr int EPI_GROUPS = 2;                              // warp groups taking alternate component tiles
constexpr int SLOTS_PER_WARP = SLOTS / (EPI_WARPS / EPI_GROUPS / 4);   // 4: a state's slots must not straddle this group
constexpr int MAX_TSLOTS = 8, MAX_NCH = 8;
constexpr uint32_t BLOCK_BYTES = BM * BK * 2;      // 16 KB: one k-block of A' or B'
constexpr uint32_t IDESC = tc::umma_idesc(BM, BN, false);
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;
constexpr float BIAS_OFF = -65504.f;               // bias of absent components: exp() underflows to exactly 0
}  // namespace tc16

// One step of the fused first pass of the LNA normalisation (aku/phone_probs.cc:227-232): running maximum nMx of the
// float-cast state log-likelihoods of a frame and nR = sum of all the other terms relative to it.  Explicit roundings:
// gmm_tc16_kernel's epilogue and tc16_norm_replay must produce the same bits.
__device__ __forceinline__ void norm_update(float res, float &nMx, float &nR)
{
  const float Lc = log_of_float_cast(res);
  if (Lc > nMx) {
    nR = (nMx == -INFINITY) ? 0.f : __fmul_rn(__fadd_rn(nR, 1.f), tc::ex2f(__fmul_rn(__fsub_rn(nMx, Lc), tc::LOG2E)));
    nMx = Lc;
  } else {
    nR = __fadd_rn(nR, tc::ex2f(__fmul_rn(__fsub_rn(Lc, nMx), tc::LOG2E)));      // exp2(-inf) = 0 covers flushed states
  }
}
// Merge of the four partial (maximum, sum) pairs of a frame, in the epilogue's order.
__device__ __forceinline__ float2 norm_merge(const float (&M)[4], const double (&R)[4])
{
  float gM = M[0];
  double Rt = R[0];
  for (int o = 1; o < 4; o++) {
    const float oM = M[o];
    const double oR = R[o];
    if (oM > gM) { Rt = oR + ((gM == -INFINITY) ? 0.0 : (1.0 + Rt) * exp((double)(gM - oM))); gM = oM; }
    else if (oM != -INFINITY) Rt += (1.0 + oR) * exp((double)(oM - gM));
  }
  return make_float2(gM, (float)log1p(Rt));
}

// The normaliser of a frame must not depend on how the call was cut into launches: when the component tiles of a
// (short) launch are spread over grid.y the epilogue cannot produce it, and this kernel REPLAYS the epilogue's
// accumulation from the stored scores -- the same four partial streams (tiles of one parity x slots of one half tile, in
// tile and slot order), the same operations, the same merge -- so a frame's LNA bytes are identical whichever launch
// shape scored it (per-utterance checksums of differently batched runs are compared, multigpu.cu).
// Four threads per frame, one per partial stream; two tiles (up to eight scores) are loaded before their updates run.
__global__ void __launch_bounds__(256)
tc16_norm_replay(const float *__restrict__ sll, int64_t ldF, int64_t nf, const int *__restrict__ meta, int n_tiles,
                 float2 *__restrict__ norm)
{
  const int o = threadIdx.x & 3;                                   // partial o as the epilogue merges them:
  const int group = o >> 1, half = o & 1;                          // 0 (group 0, half 0), 1 (0, 1), 2 (1, 0), 3 (1, 1)
  const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2);
  const bool valid = f < nf;
  const float *col = sll + (valid ? f : 0);
  float nMx = -INFINITY, nR = 0.f;
  const int4 *meta4 = reinterpret_cast<const int4 *>(meta);
  for (int n = group; n < n_tiles; n += 4) {
    const int4 ma = __ldg(meta4 + (size_t)n * 2 + half);
    const bool second = n + 2 < n_tiles;
    const int4 mb = second ? __ldg(meta4 + (size_t)(n + 2) * 2 + half) : make_int4(-1, -1, -1, -1);
    const int m8[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (m8[k] >= 0 && (m8[k] & 1)) ? __ldg(col + (int64_t)(m8[k] >> 2) * ldF) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (m8[k] >= 0 && (m8[k] & 1)) norm_update(v[k], nMx, nR);
  }
  // the four partials of a frame sit in neighbouring lanes
  float M[4];
  double R[4];
  const int base = (threadIdx.x & 31) & ~3;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    M[k] = __shfl_sync(0xffffffffu, nMx, base + k);
    R[k] = (double)__shfl_sync(0xffffffffu, nR, base + k);
  }
  if (valid && o == 0) norm[f] = norm_merge(M, R);
}

// NCH > 0: A' resident (built in the kernel), B' streamed one component tile per ring slot.
// NCH == 0: streaming variant for wide expansions (full covariance: K = D(D+3)/2 + 2 = 821): A' = [Ah | Al] comes
//           from an expansion kernel through TMA like B' = [Bh | Bl]; a ring stage = the same 64-wide k-block of the
//           four halves (64 KB); `nkb` k-blocks per half, `tslots` stages.
template <int NCH, bool PAIR = true>
__global__ void __launch_bounds__(tc16::THREADS, 1)
gmm_tc16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int nkb, int tslots,
                const int *__restrict__ range_begin,
                const int *__restrict__ meta, const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t nf, int D,
                const double *__restrict__ center, const float *__restrict__ escale, float *__restrict__ sll, int64_t ldF,
                float2 *__restrict__ norm, int *__restrict__ ovf_flag)
{
  using namespace tc;
  using namespace tc16;
  constexpr bool STREAM = NCH == 0;
  constexpr int KB = STREAM ? 4 : (2 * NCH + 3) / 4;          // 64-wide k-blocks of A' and of B' (STREAM: per ring stage)
  constexpr uint32_t SLOT_BYTES = KB * BLOCK_BYTES;           // one component tile of B' / one stage of {Ah, Al, Bh, Bl}
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[MAX_TSLOTS], empty_bar[MAX_TSLOTS], tmem_full[2], tmem_empty[2], a_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float sM[EPI_WARPS][32];
  __shared__ double sR[EPI_WARPS][32];
  unsigned char *ring = STREAM ? smem : smem + SLOT_BYTES;    // resident A' occupies the first KB blocks
  float *stage_x = reinterpret_cast<float *>(ring + (size_t)tslots * SLOT_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n_begin = range_begin[blockIdx.y], n_end = range_begin[blockIdx.y + 1];

  // cluster of CL frame tiles (launch attribute; 1 = none): every CTA fetches 1 / CL of each component tile of B' and
  // multicasts it to all of them -- the same tile order in every CTA of a cluster (same blockIdx.y)
  const uint32_t cl_size = STREAM ? 1u : cluster_nctarank(), cl_rank = STREAM ? 0u : cluster_ctarank();
  const uint16_t cl_mask = (uint16_t)((1u << cl_size) - 1u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < tslots; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], cl_size); }
    for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS / EPI_GROUPS); }
    mbar_init(&a_full, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // 512 TMEM columns = two accumulator sets of MAIN [0,128) + CORR [128,256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  if (cl_size > 1) cluster_sync_all();                          // every CTA's barriers exist before a peer's copy or commit reaches them

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: one component tile of B' (KB k-blocks, one barrier) per ring slot =====
      int slot = 0;
      uint32_t ph = 0;
      if (STREAM) {
        const int half = nkb * BK;                           // columns of Ah / Bh
        for (int n = n_begin; n < n_end; n++)
          for (int kb = 0; kb < nkb; kb++) {
            mbar_wait(&empty_bar[slot], ph ^ 1);
            mbar_expect_tx(&full_bar[slot], SLOT_BYTES);
            unsigned char *dst = ring + (size_t)slot * SLOT_BYTES;
            tma_load_2d(dst, &mapA, kb * BK, m0, &full_bar[slot]);
            tma_load_2d(dst + BLOCK_BYTES, &mapA, half + kb * BK, m0, &full_bar[slot]);
            tma_load_2d(dst + 2 * BLOCK_BYTES, &mapB, kb * BK, n * BN, &full_bar[slot]);
            tma_load_2d(dst + 3 * BLOCK_BYTES, &mapB, half + kb * BK, n * BN, &full_bar[slot]);
            if (++slot == tslots) { slot = 0; ph ^= 1; }
          }
      } else
      if (cl_size > 1) {
        // this CTA's rows of every k-block, to every CTA of the cluster: the slot is free when ALL of them have read it
        // (their MMA commits arrive on every CTA's barrier), and every CTA's barrier counts the whole tile
        const int rows = BN / (int)cl_size;
        const uint32_t part = (uint32_t)cl_rank * (uint32_t)rows * 128u;
        for (int n = n_begin; n < n_end; n++) {
          mbar_wait(&empty_bar[slot], ph ^ 1);
          mbar_expect_tx(&full_bar[slot], SLOT_BYTES);
          unsigned char *dst = ring + (size_t)slot * SLOT_BYTES + part;
#pragma unroll
          for (int kb = 0; kb < KB; kb++)
            tma_load_2d_mc(dst + kb * BLOCK_BYTES, &mapA, kb * BK, n * BN + (int)cl_rank * rows, &full_bar[slot], cl_mask);
          if (++slot == tslots) { slot = 0; ph ^= 1; }
        }
      } else
      for (int n = n_begin; n < n_end; n++) {
        mbar_wait(&empty_bar[slot], ph ^ 1);                 // a fresh barrier passes the wait on the previous phase
        mbar_expect_tx(&full_bar[slot], SLOT_BYTES);
        unsigned char *dst = ring + (size_t)slot * SLOT_BYTES;
#pragma unroll
        for (int kb = 0; kb < KB; kb++) tma_load_2d(dst + kb * BLOCK_BYTES, &mapB, kb * BK, n * BN, &full_bar[slot]);
        if (++slot == tslots) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      // descriptors differ only in the start-address field (bits 0-13, address >> 4): chunk c of a K-major SWIZZLE_128B
      // operand sits at block c/4, byte offset (c%4)*32 of the 128 B swizzle row
      auto chunk_off = [](int c) -> uint32_t { return ((uint32_t)(c >> 2) * BLOCK_BYTES + (uint32_t)(c & 3) * 32u) >> 4; };
      const uint64_t a_desc0 = umma_desc(smem_u32(smem));
      int slot = 0;
      uint32_t ph = 0;
      if (STREAM) {
        constexpr uint32_t BLK = BLOCK_BYTES >> 4;           // descriptor units
        for (int n = n_begin; n < n_end; n++) {
          const int i = n - n_begin, a = i & 1;
          mbar_wait(&tmem_empty[a], ((i >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_main = tmem_base + a * 2 * BN, d_corr = d_main + BN;
          for (int kb = 0; kb < nkb; kb++) {
            mbar_wait(&full_bar[slot], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ah = umma_desc(smem_u32(ring) + (uint32_t)slot * SLOT_BYTES);     // Ah | Al | Bh | Bl
            const uint64_t al = ah + BLK, bh = ah + 2 * BLK, bl = ah + 3 * BLK;
#pragma unroll
            for (int k = 0; k < BK / 16; k++) {
              const uint32_t first = (kb | k) ? 1u : 0u;
              umma_f16(d_corr, ah + 2 * k, bl + 2 * k, IDESC, first);
              umma_f16(d_corr, al + 2 * k, bh + 2 * k, IDESC, 1u);
              umma_f16(d_main, ah + 2 * k, bh + 2 * k, IDESC, first);
            }
            umma_commit(&empty_bar[slot]);
            if (++slot == tslots) { slot = 0; ph ^= 1; }
          }
          umma_commit(&tmem_full[a]);
        }
      } else {
      mbar_wait(&a_full, 0);                        // A' written by the epilogue warps (generic proxy, fenced)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int n = n_begin; n < n_end; n++) {
        const int i = n - n_begin, a = i & 1;
        mbar_wait(&tmem_empty[a], ((i >> 1) & 1) ^ 1);
        mbar_wait(&full_bar[slot], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_main = tmem_base + a * 2 * BN, d_corr = d_main + BN;
        const uint64_t b_desc0 = umma_desc(smem_u32(ring) + (uint32_t)slot * SLOT_BYTES);
        // CORR = Ah.Bl + Al.Bh, then MAIN = Ah.Bh (grouped by accumulator)
#pragma unroll
        for (int j = 0; j < NCH; j++) umma_f16(d_corr, a_desc0 + chunk_off(j), b_desc0 + chunk_off(NCH + j), IDESC, j > 0 ? 1u : 0u);
#pragma unroll
        for (int j = 0; j < NCH; j++) umma_f16(d_corr, a_desc0 + chunk_off(NCH + j), b_desc0 + chunk_off(j), IDESC, 1u);
#pragma unroll
        for (int j = 0; j < NCH; j++) umma_f16(d_main, a_desc0 + chunk_off(j), b_desc0 + chunk_off(j), IDESC, j > 0 ? 1u : 0u);
        if (cl_size > 1) umma_commit_mc(&empty_bar[slot], cl_mask);      // ... in every CTA of the cluster: they all refill it
        else umma_commit(&empty_bar[slot]);    // ring slot reusable once these MMAs have read it
        umma_commit(&tmem_full[a]);            // both accumulators of this tile complete
        if (++slot == tslots) { slot = 0; ph ^= 1; }
      }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    const int half = ((warp - 4) >> 2) & 1;                   // which SLOTS_PER_WARP slots of the tile
    const int group = (warp - 4) >> 3;                        // this warp takes the tiles i with (i & 1) == group, accumulator set = group
    const int et = threadIdx.x - 128;                         // 0..511 among the epilogue threads
    // ===== A' = [Ah | Al]: stage, centre, expand, scale, split, store swizzled =====
    if (!STREAM) {
      const int64_t valid = nf - m0;                          // frames of this tile that exist
      const int64_t g0 = (f_begin + m0) * (int64_t)D;
      const int total = BM * D;
      for (int idx = et; idx < total; idx += EPI_THREADS) {
        const int r = idx / D, d = idx - r * D;
        float v = 0.f;
        if (r < valid) {
          const double t = feats_f64 ? reinterpret_cast<const double *>(feats)[g0 + idx]
                                     : (double)reinterpret_cast<const float *>(feats)[g0 + idx];
          v = (float)(t - center[d]);
        }
        stage_x[idx] = v;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      const int r = et >> 2;
      const float *x = stage_x + r * D;
      const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
      bool ovf = false;
      for (int v = et & 3; v < 2 * NCH; v += 4) {            // vector = 8 consecutive K terms
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float val[2];
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const int k = v * 8 + i * 2 + j;
            float t = 0.f;
            if (k < D) t = x[k] * x[k];
            else if (k < 2 * D) t = x[k - D];
            else if (k < 2 * D + 2) t = 1.f;
            val[j] = t * __ldg(escale + k);
            if (!(fabsf(val[j]) <= 65504.f)) ovf = true;
          }
          const __half2 h2 = __floats2half2_rn(val[0], val[1]);
          const float2 hf = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn((val[0] - hf.x) * LO_SCALE, (val[1] - hf.y) * LO_SCALE);
          hw[i] = *reinterpret_cast<const uint32_t *>(&h2);
          lw[i] = *reinterpret_cast<const uint32_t *>(&l2);
        }
        const int ch = v >> 1, cl = NCH + ch;                // chunk of the hi / lo half
        const uint32_t jh = (uint32_t)((ch & 3) * 2 + (v & 1)), jl = (uint32_t)((cl & 3) * 2 + (v & 1));
        unsigned char *ph = smem + (size_t)(ch >> 2) * BLOCK_BYTES + row_off + ((jh ^ (uint32_t)(r & 7)) << 4);
        unsigned char *pl = smem + (size_t)(cl >> 2) * BLOCK_BYTES + row_off + ((jl ^ (uint32_t)(r & 7)) << 4);
        *reinterpret_cast<uint4 *>(ph) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4 *>(pl) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      if (ovf) atomicOr(ovf_flag, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full);
    }
    // ===== epilogue: thread = one frame row of the tile =====
    const int64_t frame = (int64_t)m0 + q * 32 + lane;
    float run_a = 0.f, run_s = 0.f;
    // fused first pass of the LNA epilogue (aku/phone_probs.cc:227-232): running maximum of the float-cast
    // state likelihoods of this frame and the sum of all the other terms relative to it
    float nMx = -INFINITY, nR = 0.f;
    float *const sll_f = sll + frame;
    auto process = [&](const uint32_t (&rm)[16], const uint32_t (&rc)[16], int mt) {
      // packed fp32 pairs (FFMA2 / FADD2) wherever two components take the same operation: the epilogue is
      // issue-bound next to the MUFU pipe (16 exp2 per slot)
      float2 v[8];
#pragma unroll
      for (int c = 0; c < 8; c++)
        v[c] = __ffma2_rn(make_float2(__uint_as_float(rc[2 * c]), __uint_as_float(rc[2 * c + 1])), make_float2(LO_INV, LO_INV),
                          make_float2(__uint_as_float(rm[2 * c]), __uint_as_float(rm[2 * c + 1])));
      const bool first = (mt & 2) != 0, last = (mt & 1) != 0;
      float m8[8], m4[4];
#pragma unroll
      for (int c = 0; c < 8; c++) m8[c] = fmaxf(v[c].x, v[c].y);
#pragma unroll
      for (int c = 0; c < 4; c++) m4[c] = fmaxf(m8[c], m8[c + 4]);
      float mx = fmaxf(fmaxf(m4[0], m4[2]), fmaxf(m4[1], m4[3]));
      if (!first) mx = fmaxf(mx, run_a);
      const float nml = -mx * LOG2E;
      float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const float2 e0 = __ffma2_rn(v[c], make_float2(LOG2E, LOG2E), make_float2(nml, nml));
        const float2 e1 = __ffma2_rn(v[c + 1], make_float2(LOG2E, LOG2E), make_float2(nml, nml));
        s0 = __fadd2_rn(s0, make_float2(ex2f(e0.x), ex2f(e0.y)));
        s1 = __fadd2_rn(s1, make_float2(ex2f(e1.x), ex2f(e1.y)));
      }
      float sum = (s0.x + s0.y) + (s1.x + s1.y);
      if (!first) sum = fmaf(run_s, ex2f(fmaf(run_a, LOG2E, nml)), sum);
      run_a = mx;
      run_s = sum;
      if (last && mt >= 0) {
        const float res = fmaf(lg2f(sum), LN2, mx);
        sll_f[(int64_t)(mt >> 2) * ldF] = res;
        if (norm) norm_update(res, nMx, nR);
      }
    };
    // Two slots at once: the same arithmetic in the same order as two process() calls, but both slots' maxima and
    // arguments come first and the 32 exp2 are issued in ONE burst before either slot's tail (sum, carry, log, store,
    // normaliser -- the branchy part).  process() leaves 16-wide bursts separated by ~200 instructions of tail and
    // prologue; with four epilogue warps per scheduler that kept the XU pipe 67 % busy (profiles/r02_gmm_tc16_ncu_full.txt).
    auto tail = [&](const float2 (&x)[8], float mx, float nml, bool first, bool last, int mt) {
      float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        s0 = __fadd2_rn(s0, x[c]);
        s1 = __fadd2_rn(s1, x[c + 1]);
      }
      float sum = (s0.x + s0.y) + (s1.x + s1.y);
      if (!first) sum = fmaf(run_s, ex2f(fmaf(run_a, LOG2E, nml)), sum);
      run_a = mx;
      run_s = sum;
      if (last && mt >= 0) {
        const float res = fmaf(lg2f(sum), LN2, mx);
        sll_f[(int64_t)(mt >> 2) * ldF] = res;
        if (norm) norm_update(res, nMx, nR);
      }
    };
    auto process2 = [&](const uint32_t (&rmA)[16], const uint32_t (&rcA)[16], int mtA, const uint32_t (&rmB)[16],
                        const uint32_t (&rcB)[16], int mtB) {
      float2 vA[8], vB[8];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        vA[c] = __ffma2_rn(make_float2(__uint_as_float(rcA[2 * c]), __uint_as_float(rcA[2 * c + 1])), make_float2(LO_INV, LO_INV),
                           make_float2(__uint_as_float(rmA[2 * c]), __uint_as_float(rmA[2 * c + 1])));
        vB[c] = __ffma2_rn(make_float2(__uint_as_float(rcB[2 * c]), __uint_as_float(rcB[2 * c + 1])), make_float2(LO_INV, LO_INV),
                           make_float2(__uint_as_float(rmB[2 * c]), __uint_as_float(rmB[2 * c + 1])));
      }
      const bool firstA = (mtA & 2) != 0, lastA = (mtA & 1) != 0, firstB = (mtB & 2) != 0, lastB = (mtB & 1) != 0;
      auto max16 = [](const float2 (&v)[8]) {
        float m8[8], m4[4];
#pragma unroll
        for (int c = 0; c < 8; c++) m8[c] = fmaxf(v[c].x, v[c].y);
#pragma unroll
        for (int c = 0; c < 4; c++) m4[c] = fmaxf(m8[c], m8[c + 4]);
        return fmaxf(fmaxf(m4[0], m4[2]), fmaxf(m4[1], m4[3]));
      };
      float mxA = max16(vA), mxB = max16(vB);
      if (!firstA) mxA = fmaxf(mxA, run_a);
      if (!firstB) mxB = fmaxf(mxB, mxA);                       // what run_a is once slot A is done
      const float nmlA = -mxA * LOG2E, nmlB = -mxB * LOG2E;
      float2 xA[8], xB[8];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        xA[c] = __ffma2_rn(vA[c], make_float2(LOG2E, LOG2E), make_float2(nmlA, nmlA));
        xB[c] = __ffma2_rn(vB[c], make_float2(LOG2E, LOG2E), make_float2(nmlB, nmlB));
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        xA[c] = make_float2(ex2f(xA[c].x), ex2f(xA[c].y));
        xB[c] = make_float2(ex2f(xB[c].x), ex2f(xB[c].y));
      }
      tail(xA, mxA, nmlA, firstA, lastA, mtA);
      tail(xB, mxB, nmlB, firstB, lastB, mtB);
    };
    static_assert(SLOTS_PER_WARP == 4 && EPI_GROUPS == 2, "the epilogue below handles four slots per warp and tile, two warp groups");
    // The two warp groups work on alternate tiles: while one group waits for its accumulators and loads them, the
    // other one keeps the MUFU / FMA pipes busy (all 16 warps on one tile left those pipes idle during every
    // wait + tcgen05.ld phase).
    const int4 *meta4 = reinterpret_cast<const int4 *>(meta);
    for (int n = n_begin + group; n < n_end; n += EPI_GROUPS) {
      const int use = (n - n_begin) >> 1;
      const int4 mt = __ldg(meta4 + (size_t)n * 2 + half);
      mbar_wait(&tmem_full[group], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + group * 2 * BN + half * SLOTS_PER_WARP * GR;
      uint32_t r0[16], c0[16], r1[16], c1[16];
      AKU_TMEM_LD16(r0, taddr);
      AKU_TMEM_LD16(c0, taddr + BN);
      AKU_TMEM_LD16(r1, taddr + GR);
      AKU_TMEM_LD16(c1, taddr + GR + BN);
      AKU_TMEM_LD_WAIT();
      if (PAIR) process2(r0, c0, mt.x, r1, c1, mt.y);
      else { process(r0, c0, mt.x); process(r1, c1, mt.y); }
      AKU_TMEM_LD16(r0, taddr + 2 * GR);
      AKU_TMEM_LD16(c0, taddr + 2 * GR + BN);
      AKU_TMEM_LD16(r1, taddr + 3 * GR);
      AKU_TMEM_LD16(c1, taddr + 3 * GR + BN);
      AKU_TMEM_LD_WAIT();
      // every TMEM read of this warp is complete: hand the accumulator set back before the arithmetic
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[group]);
      if (PAIR) process2(r0, c0, mt.z, r1, c1, mt.w);
      else { process(r0, c0, mt.z); process(r1, c1, mt.w); }
    }
    if (norm) {   // the four warps of a lane quarter merge their parts; one float2 per frame
      sM[warp - 4][lane] = nMx;
      sR[warp - 4][lane] = (double)nR;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (warp - 4 < 4) {
        static_assert(EPI_WARPS == 16, "four partial streams per frame: (group, half)");
        float M[4];
        double R[4];
#pragma unroll
        for (int o = 0; o < 4; o++) { M[o] = sM[q + 4 * o][lane]; R[o] = sR[q + 4 * o][lane]; }
        norm[frame] = norm_merge(M, R);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cl_size > 1) cluster_sync_all();      // no CTA leaves while a peer's copy or commit may still be on its way to it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
static inline uint16_t half_bits(float v)
{
  const __half h = __float2half_rn(v);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline float half_val(uint16_t b)
{
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}

static size_t tc16_stage_x_bytes(int D) { return ((size_t)tc16::BM * D * sizeof(float) + 1023) / 1024 * 1024; }
// ring slots (one component tile of B' each) that fit beside A' and the feature staging area
static int tc16_tslots(int KB, int D, bool leave_room = false)
{
  const size_t budget = 227 * 1024 - 8192 /* static shared memory */ - 1024 /* alignment */;
  const size_t fixed = (size_t)KB * tc16::BLOCK_BYTES + tc16_stage_x_bytes(D);
  if (fixed >= budget) return 0;
  int n = (int)std::min<size_t>(tc16::MAX_TSLOTS, (budget - fixed) / ((size_t)KB * tc16::BLOCK_BYTES));
  // leave_room: the LNA epilogue of the previous chunk shares the SM (33 KB of shared memory per CTA, api.cu); two ring
  // slots cover the L2 latency of a 48 KB tile at the ~1 us a tile takes
  if (leave_room) while (n > 2 && fixed + (size_t)n * KB * tc16::BLOCK_BYTES + 40 * 1024 > budget) n--;
  static const char *force = getenv("AKUGPU_TC16_SLOTS");
  if (force && atoi(force) >= 2) n = std::min(n, atoi(force));
  return n;
}
constexpr int TC16_STREAM_STAGES = 3;
constexpr int TC16_DEFAULT_CLUSTER = 1;    // frame tiles per cluster sharing the fetch of B' (see launch_gmm_tc16)      // 3 x 64 KB stages of {Ah, Al, Bh, Bl}

// Expansion width (terms incl. the two constant terms) and whether A' can stay resident in shared memory.
static void tc16_shape(const HostModel &hm, int &L, int &NCH, bool &stream)
{
  const bool full = hm.n_full > 0;
  L = (full ? hm.D * (hm.D + 3) / 2 : 2 * hm.D) + 2;
  NCH = (L + 15) / 16;
  stream = full || NCH > tc16::MAX_NCH || tc16_tslots((2 * NCH + 3) / 4, hm.D) < 2;
}

bool tc16_supported(const HostModel &hm)
{
  if (hm.S <= 0 || hm.G <= 0 || (hm.n_full != 0 && hm.n_full != hm.G)) return false;   // all diagonal or all full
  for (int s = 0; s < hm.S; s++)
    if (hm.mix_off[s + 1] - hm.mix_off[s] > tc16::SLOTS_PER_WARP * tc16::GR) return false;
  return true;
}

// Streaming variant: A' = [Ah | Al] (rows x 2*half fp16) from the features: centre, expand, scale, split.
// Full covariance: [x ; vec(x x^T)] (lower triangle row-wise, off-diagonals x sqrt 2, aku/LinearAlgebra.cc:220-238);
// diagonal: [x^2 ; x]; then the two constant terms.
__global__ void tc16_expand_feats(const void *__restrict__ feats, int feats_f64, int64_t f_begin, int64_t nf, int64_t rows,
                                  int D, int L, int half, int full, const double *__restrict__ center,
                                  const float *__restrict__ escale, __half *__restrict__ A, int *__restrict__ ovf_flag)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * half) return;
  const int64_t fr = i / half;
  const int l = (int)(i - fr * half);
  float v = 0.f;
  if (fr < nf && l < L) {
    auto x = [&](int d) -> double {
      const double t = feats_f64 ? reinterpret_cast<const double *>(feats)[(f_begin + fr) * D + d]
                                 : (double)reinterpret_cast<const float *>(feats)[(f_begin + fr) * D + d];
      return t - center[d];
    };
    const int L0 = L - 2;
    if (l >= L0) v = 1.f;
    else if (full) {
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
    v *= escale[l];
    if (!(fabsf(v) <= 65504.f)) atomicOr(ovf_flag, 1);
  }
  const __half h = __float2half_rn(v);
  A[fr * (2 * (int64_t)half) + l] = h;
  A[fr * (2 * (int64_t)half) + half + l] = __float2half_rn((v - __half2float(h)) * tc16::LO_SCALE);
}

// Builds B' = [Bh | Bl] (slot-ordered component rows, fp16), the slot table and the per-term scales.
void model_pack_tc16(akugpu_ctx *ctx)
{
  const HostModel &hm = ctx->hm;
  PackedTC16 &p = ctx->ptc16;
  p.ready = false;
  p.q_max = 0;
  if (!tc16_supported(hm)) return;
  const int G = hm.G, D = hm.D;
  const bool full = hm.n_full > 0;
  p.D = D;
  p.full = full;
  tc16_shape(hm, p.L, p.NCH, p.stream);
  const int L0 = p.L - 2;                        // data terms; then b_coarse, b_fine
  if (p.stream) {
    p.KB = (p.NCH + 3) / 4;                      // k-blocks per half
    p.half = p.KB * tc16::BK;
    p.Kp = 2 * p.half;
  } else {
    p.KB = (2 * p.NCH + 3) / 4;                  // k-blocks of the whole row
    p.half = p.NCH * 16;
    p.Kp = p.KB * tc16::BK;
  }
  const int half = p.half, Kp = p.Kp;
  std::vector<double> cen, theta, gconst;
  std::vector<double> q_of;
  p.q_max = tc_expanded_params(hm, full, L0, cen, theta, gconst, &q_of);
  // States with a component that is ill-conditioned for the expanded form (sharp Gaussians far from the centre) are
  // left to the direct-form FP32-pipe kernel (diagonal pools: model_pack packs exactly those states into the FP32
  // image and both kernels run per chunk); when half of the states or more are like that, or the pool is full
  // covariance, the tensor-core scorer is not used at all.  akugpu_set_scorer_variant(ctx, 3) switches the check off.
  p.hybrid = false;
  p.n_bad = 0;
  p.bad_state.assign(hm.S, 0);
  if (ctx->scorer_variant != 3 && p.q_max > TC_Q_MAX) {
    for (int st = 0; st < hm.S; st++)
      for (int k = hm.mix_off[st]; k < hm.mix_off[st + 1]; k++)
        if (hm.mix_w[k] > 0 && q_of[hm.mix_gauss[k]] > TC_Q_MAX) { p.bad_state[st] = 1; break; }
    for (int st = 0; st < hm.S; st++) p.n_bad += p.bad_state[st];
    if (full || 2 * p.n_bad >= hm.S) return;
    p.hybrid = p.n_bad > 0;
  }
  std::vector<double> tmax(L0, 0.0);
  for (int g = 0; g < G; g++)
    for (int l = 0; l < L0; l++) tmax[l] = std::max(tmax[l], fabs(theta[(size_t)g * L0 + l]));
  // power-of-two scale of every term: max |B'_k| in [2^11, 2^12); the two constant terms stay unscaled
  std::vector<float> escale(half, 0.f);
  std::vector<int> ek(L0, 0);
  for (int l = 0; l < L0; l++) {
    if (tmax[l] > 0 && std::isfinite(tmax[l])) ek[l] = ilogb(tmax[l]) - 11;
    ek[l] = std::max(-100, std::min(100, ek[l]));
    escale[l] = (float)ldexp(1.0, ek[l]);
  }
  escale[L0] = escale[L0 + 1] = 1.f;

  std::vector<int> slot_state, slot_k0, slot_flags;
  tc_build_slots(hm, tc16::SLOTS_PER_WARP, slot_state, slot_k0, slot_flags, p.hybrid ? &p.bad_state : nullptr);
  const int n_slots = (int)slot_state.size();
  p.n_tiles = (n_slots + tc16::SLOTS - 1) / tc16::SLOTS;
  const size_t rows = (size_t)p.n_tiles * tc16::BN;
  std::vector<uint16_t> B(rows * Kp, 0);
  std::vector<int32_t> meta((size_t)p.n_tiles * tc16::SLOTS, -1);
  p.clean.assign(p.n_tiles, 1);
  const uint16_t off_bits = half_bits(tc16::BIAS_OFF);
  auto put = [&](uint16_t *br, int l, double val) {        // hi at l, lo (scaled by 2^11) at half + l
    const float v = (float)val;
    const uint16_t h = half_bits(v);
    br[l] = h;
    br[half + l] = half_bits((float)((val - (double)half_val(h)) * (double)tc16::LO_SCALE));
  };
  for (size_t row = 0; row < rows; row++) B[row * Kp + L0] = off_bits;     // absent components: constant = -65504
  for (int sl = 0; sl < n_slots; sl++) {
    const int s = slot_state[sl];
    if (s < 0) continue;                                   // padding slot: meta stays -1
    meta[sl] = (s << 2) | slot_flags[sl];
    if (sl % tc16::SLOTS == 0 && !(slot_flags[sl] & 2)) p.clean[sl / tc16::SLOTS] = 0;
    const int K = hm.mix_off[s + 1] - hm.mix_off[s];
    for (int j = 0; j < tc16::GR && slot_k0[sl] + j < K; j++) {
      const int k = hm.mix_off[s] + slot_k0[sl] + j, g = hm.mix_gauss[k];
      const double w = hm.mix_w[k];
      if (!(w > 0)) continue;                              // weight 0: contributes nothing (bias stays -65504)
      const double b = log(w) + gconst[g];
      if (!(fabs(b) < 60000.0)) { p.ready = false; return; }   // constant outside the fp16 range: keep the bf16x3 kernel
      uint16_t *br = &B[((size_t)sl * tc16::GR + j) * Kp];
      for (int l = 0; l < L0; l++) put(br, l, ldexp(theta[(size_t)g * L0 + l], -ek[l]));
      const uint16_t bc = half_bits((float)b);
      br[L0] = bc;                                         // coarse part, exact in fp16
      br[half + L0] = 0;
      put(br, L0 + 1, b - (double)half_val(bc));           // fine part, hi/lo
    }
  }
  auto up = [&](DevBuf &buf, const void *src, size_t bytes) {
    buf.reserve(std::max<size_t>(bytes, 16));
    AKU_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  };
  up(p.B, B.data(), B.size() * 2);
  up(p.meta, meta.data(), meta.size() * 4);
  up(p.center, cen.data(), cen.size() * 8);
  p.h_center = cen;
  p.map_ready = false;
  up(p.escale, escale.data(), escale.size() * 4);
  p.flag.reserve(16);
  AKU_CUDA(cudaMemsetAsync(p.flag.p, 0, 16, ctx->stream));
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  p.ranges.clear();
  p.ready = true;
}

int64_t gmm_tc16_wave_frames(akugpu_ctx *ctx) { return (int64_t)ctx->sm_count * tc16::BM; }

bool gmm_tc16_overflowed(akugpu_ctx *ctx)
{
  PackedTC16 &p = ctx->ptc16;
  if (!p.ready) return false;
  int h = 0;
  AKU_CUDA(cudaMemcpyAsync(&h, p.flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h) {
    AKU_CUDA(cudaMemsetAsync(p.flag.p, 0, sizeof(int), ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return h != 0;
}

bool launch_gmm_tc16(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t f_begin, int64_t f_end, float *sll, int64_t ldF,
                     float2 *norm)
{
  PackedTC16 &p = ctx->ptc16;
  const int64_t nf = f_end - f_begin;
  if (nf <= 0) return false;
  CUtensorMap mapA, mapB;
  tc_make_map(&mapB, p.B.p, (uint64_t)p.n_tiles * tc16::BN, (uint64_t)p.Kp, true);
  const int ftiles = (int)((nf + tc16::BM - 1) / tc16::BM);
  if (p.stream) {      // expansion kernel (accounted to the front-end stage, like the bf16x3 path's)
    const int64_t rows = (int64_t)ftiles * tc16::BM;
    ctx->d_fe[4].reserve((size_t)rows * p.Kp * 2);
    StageScope sc(ctx, 0);
    const int64_t ne = rows * p.half;
    tc16_expand_feats<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(feats, feats_f64, f_begin, nf, rows, p.D, p.L, p.half,
                                                                           p.full ? 1 : 0, p.center.as<double>(), p.escale.as<float>(),
                                                                           ctx->d_fe[4].as<__half>(), p.flag.as<int>());
    AKU_CUDA(cudaGetLastError());
    ctx->launches++;
    tc_make_map(&mapA, ctx->d_fe[4].p, (uint64_t)rows, (uint64_t)p.Kp, true);
  } else {
    mapA = mapB;       // unused by the resident variant
  }
  int want = 1;
  if (ftiles < ctx->sm_count) want = std::min(p.n_tiles, std::max(1, ctx->sm_count / ftiles));
  int ysplit = 1;
  const int *ranges = tc_tile_ranges(ctx, p.n_tiles, p.clean, p.ranges, want, ysplit);
  float2 *replay_norm = nullptr;          // a frame's states are spread over several CTAs: the normaliser is replayed afterwards
  if (ysplit != 1) { replay_norm = norm; norm = nullptr; }
  const int tslots = p.stream ? TC16_STREAM_STAGES : tc16_tslots(p.KB, p.D, ctx->overlap_lna);
  const size_t smem = p.stream ? 1024 + (size_t)tslots * 4 * tc16::BLOCK_BYTES
                               : 1024 + (size_t)(1 + tslots) * p.KB * tc16::BLOCK_BYTES + tc16_stage_x_bytes(p.D);
  StageScope sc(ctx, 1);
  // cluster of frame tiles sharing the fetch of B' (tma multicast): AKUGPU_TC16_CLUSTER = 1 / 2 / 4
  static const int want_cl = getenv("AKUGPU_TC16_CLUSTER") ? atoi(getenv("AKUGPU_TC16_CLUSTER")) : TC16_DEFAULT_CLUSTER;
  int cl = (!p.stream && (want_cl == 2 || want_cl == 4) && ftiles % want_cl == 0) ? want_cl : 1;
  if (cl > 1) tc_make_map(&mapA, p.B.p, (uint64_t)p.n_tiles * tc16::BN, (uint64_t)p.Kp, true, tc16::BK, tc16::BN / cl);
  auto launch = [&](auto kernel) {
    ensure_dynamic_smem(ctx, (const void *)kernel, smem);
    const int nkb = p.KB;
    if (cl == 1) {
      kernel<<<dim3(ftiles, ysplit), tc16::THREADS, smem, ctx->stream>>>(mapA, mapB, nkb, tslots, ranges, p.meta.as<int>(), feats, feats_f64,
                                                                       f_begin, nf, p.D, p.center.as<double>(), p.escale.as<float>(), sll, ldF,
                                                                       norm, p.flag.as<int>());
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ftiles, ysplit);
      cfg.blockDim = dim3(tc16::THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      AKU_CUDA(cudaLaunchKernelEx(&cfg, kernel, mapA, mapB, nkb, tslots, ranges, (const int *)p.meta.as<int>(), feats, feats_f64, f_begin, nf, p.D,
                                  (const double *)p.center.as<double>(), (const float *)p.escale.as<float>(), sll, ldF, norm, p.flag.as<int>()));
    }
  };
  if (p.stream) launch(gmm_tc16_kernel<0>);
  else switch (p.NCH) {
    case 1: launch(gmm_tc16_kernel<1>); break;
    case 2: launch(gmm_tc16_kernel<2>); break;
    case 3: launch(gmm_tc16_kernel<3>); break;
    case 4: launch(gmm_tc16_kernel<4>); break;
    case 5: if (getenv("AKUGPU_TC16_EPI_OLD")) launch(gmm_tc16_kernel<5, false>); else launch(gmm_tc16_kernel<5>); break;
    case 6: launch(gmm_tc16_kernel<6>); break;
    case 7: launch(gmm_tc16_kernel<7>); break;
    case 8: launch(gmm_tc16_kernel<8>); break;
    default: throw Error(AKUGPU_E_STATE, "gmm_tc16: unsupported feature dimension");
  }
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
  if (replay_norm) {
    tc16_norm_replay<<<(unsigned)((nf + 63) / 64), 256, 0, ctx->stream>>>(sll, ldF, nf, p.meta.as<int>(), p.n_tiles, replay_norm);
    AKU_CUDA(cudaGetLastError());
    ctx->launches++;
    return true;
  }
  return norm != nullptr;
}

}  // namespace akugpu
