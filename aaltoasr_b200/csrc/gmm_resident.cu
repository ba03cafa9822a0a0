// gmm_resident.cu -- resident scorer of the decoder's per-frame feed: ONE kernel that stays on the device between
// calls, with the model's parameter image held in the shared memory of the 148 SMs.
//
// The reference scores one frame per call for its stream decoder (decoder/decode-stream.cc:191-207 ->
// Toolbox::set_one_frame -> OneFrameAcoustics::set, decoder/src/OneFrameAcoustics.cc:23-30; the scoring itself is
// HmmSet::precompute_likelihoods + state_likelihood, aku/HmmSet.cc:485-501).  gmm_stream.cu serves such a call with one
// launch, and its floor is the launch (~5 us), the prologue (TMEM allocation, barriers) and a sweep of the 30 MB image
// from L2 (~6 us).  But 30.7 MB is 207 KB per SM: without the padding of the last k-block (40 KB per 128-component
// tile instead of 48: the half-filled block is fetched as a 32-column box with the 64-byte swizzle) the config-2 model
// FITS the shared memory of the chip, five tiles per CTA.  So:
//   * akugpu_stream_open() starts gmm_resident_kernel once: every CTA loads its tiles of B' by TMA and keeps them;
//     TMEM, barriers and tables are set up once;
//   * a call is a message: the host writes the centred features and a sequence number into pinned, mapped memory; one
//     lane of ONE CTA polls that word over PCIe (many SMs polling one host line are served one after the other, 2.7 us
//     each: scripts/micro_mailbox.cu), fetches the features and relays both through device memory, where the other
//     CTAs poll an L2-resident line; the epilogue warps build A', the MMA warp runs the tile's 15 MMAs against the
//     resident B', the epilogue (the one of gmm_stream_kernel: same operations, same bits) stores log-likelihoods into
//     mapped host memory, the last CTA publishes the sequence number, the host polls it.  No launch, no CUDA call, no
//     parameter traffic on the path of a call;
//   * tiles that do not fit (config-4 model: 17 per CTA) are streamed through a small ring each call, as before;
//   * the kernel ends on a quit message (any other library call sends one first: the GPU is whole again before anything
//     else runs) or by itself after `idle_ms` without a call (the polling CTA decides and relays the decision, so the
//     grid ends as one; a call that crosses that moment is detected by the host -- the stream has drained, the
//     sequence number is missing -- and repeated after a relaunch).
#include "ctx.hpp"
#include "kernels.hpp"
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "stream_common.cuh"
#include <math.h>
#include <string.h>
#include <chrono>
#include <atomic>
#include <mutex>
#include <map>
#include <algorithm>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace akugpu {

namespace tcr {
constexpr int NF = 16;                             // frames per call (N side of the MMA)
constexpr int MAX_SLOTS = 28;                      // 227 KB / 8 KB (NCH = 1)
constexpr uint32_t A_BLOCK = NF * tcs::BK * 2;     // one 64-wide k-block of A'
constexpr unsigned CMD_QUIT = 1u << 16, CMD_FLOOR = 1u << 8, CMD_BIG = 1u << 17;
constexpr int PKT_WORDS = 128, PKT_SECTORS = 16;   // a CTA's packet: 512 B = 16 sectors of {7 payload words, tag}
constexpr int MAX_GRID = 256;
}  // namespace tcr

struct ResidentCmd { unsigned int seq; int nf, use_floor, quit, direct; float floor_at; unsigned long long t_seen; };

__device__ __forceinline__ uint4 ld_sys_v4(const unsigned int *p)
{
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f32(const float *p)
{
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// grid = (1, number of tile ranges); one CTA per SM.
template <int NCH>
__global__ void __launch_bounds__(tcs::THREADS, 1)
gmm_resident_kernel(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapB64, int nslots,
                    const int *__restrict__ range_begin, const int *__restrict__ meta, int D, int S, const float *__restrict__ escale,
                    const unsigned int *mailbox, const unsigned int *pkts, float *gx, unsigned int *gcmd, float *__restrict__ out, unsigned int *__restrict__ cnt,
                    volatile unsigned int *__restrict__ flags, unsigned int seq0, unsigned long long idle_ns)
{
  using namespace tc;
  using namespace tcs;
  using namespace tcr;
  constexpr int NCHUNK = 2 * NCH;                             // K16 chunks of a row of B' = [Bh | Bl]
  constexpr int KBF = NCHUNK / 4;                             // full 64-wide k-blocks (SWIZZLE_128B)
  constexpr bool HALF = (NCHUNK & 2) != 0;                    // + one 32-wide block (SWIZZLE_64B)
  constexpr int KBA = (NCHUNK + 3) / 4;                       // k-blocks of A' (padded, SWIZZLE_128B: it is 6 KB)
  constexpr uint32_t SLOT_BYTES = KBF * B_BLOCK + (HALF ? B_BLOCK / 2 : 0);
  constexpr uint32_t IDESC = umma_idesc(BM, NF, false);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // accumulator sets: MAIN + CORR of one tile = 2 NF columns; all 512 columns = 16 sets, so that the MMA warp runs through the
  // tiles of a call without waiting for the epilogue (set = running tile number mod 16, across calls)
  constexpr int NSETS = 512 / (2 * NF);
  __shared__ uint64_t full_bar[MAX_SLOTS], empty_bar[MAX_SLOTS], tmem_full[NSETS], tmem_empty[NSETS], a_full, call_bar, done_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ ResidentCmd s_cmd;
  __shared__ float2 part[2][2][SLOTS][NF];                    // [group][use parity][slot][frame] = {max, sum of exp}
  __shared__ int pmeta[2][2][SLOTS];
  unsigned char *ring = smem + KBA * A_BLOCK;
  float *stage_x = reinterpret_cast<float *>(ring + (size_t)nslots * SLOT_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_begin = range_begin[blockIdx.y], n_end = range_begin[blockIdx.y + 1];
  const int T = n_end - n_begin;
  // resident tiles: the first R of the range, one slot each, loaded once; the other T - R tiles pass through NR ring slots
  // (ring depth: a slot is ~1 us in flight, 40 KB; three of them keep the SM's ~120 GB/s L2 port busy)
  const int NR = T <= nslots ? 0 : (nslots >= 5 ? 3 : nslots >= 3 ? 2 : 1);
  const int R = T <= nslots ? T : nslots - NR;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < NSETS; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    mbar_init(&a_full, EPI_THREADS / 32);
    mbar_init(&call_bar, 1);
    mbar_init(&done_bar, EPI_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  auto load_tile = [&](int n, int slot) {
    mbar_expect_tx(&full_bar[slot], SLOT_BYTES);
    unsigned char *dst = ring + (size_t)slot * SLOT_BYTES;
#pragma unroll
    for (int kb = 0; kb < KBF; kb++) tma_load_2d(dst + kb * B_BLOCK, &mapB, kb * BK, n * BM, &full_bar[slot]);
    if (HALF) tma_load_2d(dst + KBF * B_BLOCK, &mapB64, KBF * BK, n * BM, &full_bar[slot]);
  };

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: the resident tiles once; the others every call =====
      for (int i = 0; i < R; i++) load_tile(n_begin + i, i);
      if (NR > 0) {
        int rs = 0;
        uint32_t rph = 0, cph = 0;
        for (;;) {
          mbar_wait(&call_bar, cph);
          cph ^= 1;
          if (s_cmd.quit) break;
          for (int i = R; i < T; i++) {
            mbar_wait(&empty_bar[R + rs], rph ^ 1);
            load_tile(n_begin + i, R + rs);
            if (++rs == NR) { rs = 0; rph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer: the component tile is the M-side operand, the frames the N-side =====
      auto aoff = [](int c) -> uint32_t { return ((uint32_t)(c >> 2) * A_BLOCK + (uint32_t)(c & 3) * 32u) >> 4; };
      const uint64_t x_desc0 = umma_desc(smem_u32(smem));
      int rs = 0;
      uint32_t rph = 0, cph = 0, aph = 0, g = 0;                  // g: running tile number
      for (;;) {
        mbar_wait(&call_bar, cph);
        cph ^= 1;
        if (s_cmd.quit) {
          for (int i = 0; i < R; i++) mbar_wait(&full_bar[i], 0);      // no bulk copy may still be in flight when the CTA ends
          break;
        }
        mbar_wait(&a_full, aph);
        aph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int i = 0; i < T; i++, g++) {
          const uint32_t a = g % NSETS, u = g / NSETS;
          mbar_wait(&tmem_empty[a], (u & 1) ^ 1);
          int slot;
          if (i < R) { slot = i; mbar_wait(&full_bar[slot], 0); }        // completed once, for good
          else { slot = R + rs; mbar_wait(&full_bar[slot], rph); }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_main = tmem_base + a * 2 * NF, d_corr = d_main + NF;
          const uint32_t sbase = smem_u32(ring) + (uint32_t)slot * SLOT_BYTES;
          const uint64_t p_desc0 = umma_desc(sbase), p_desc64 = umma_desc_sw64(sbase + KBF * B_BLOCK);
          auto bdesc = [&](int c) -> uint64_t {
            if (!HALF || c < KBF * 4) return p_desc0 + (((uint32_t)(c >> 2) * B_BLOCK + (uint32_t)(c & 3) * 32u) >> 4);
            return p_desc64 + (((uint32_t)(c - KBF * 4) * 32u) >> 4);
          };
#pragma unroll
          for (int j = 0; j < NCH; j++) umma_f16(d_corr, bdesc(NCH + j), x_desc0 + aoff(j), IDESC, j > 0 ? 1u : 0u);   // Bl . Ah
#pragma unroll
          for (int j = 0; j < NCH; j++) umma_f16(d_corr, bdesc(j), x_desc0 + aoff(NCH + j), IDESC, 1u);                // Bh . Al
#pragma unroll
          for (int j = 0; j < NCH; j++) umma_f16(d_main, bdesc(j), x_desc0 + aoff(j), IDESC, j > 0 ? 1u : 0u);         // Bh . Ah
          if (i >= R) {
            umma_commit(&empty_bar[slot]);
            if (++rs == NR) { rs = 0; rph ^= 1; }
          }
          umma_commit(&tmem_full[a]);
        }
      }
    }
  } else if (warp == 2) {
    // ===== mailbox.  Every CTA polls ITS OWN packet in mapped host memory: 32-byte sectors of {7 payload words, tag},
    // the tag (= sequence number) written last by the host, so a sector whose tag is new carries new payload and the
    // whole command -- header and the features of up to two frames -- arrives with ONE read over PCIe.  (148 CTAs
    // polling one shared host line are served one after the other, 2.7 us each = 400 us per call; own lines proceed
    // in parallel: scripts/micro_mailbox.cu, micro_mailbox2.cu.)  Calls too large for a packet go to CTA 0 alone, which
    // fetches the features from the shared area and relays them through device memory (gx, gcmd: polled in L2). =====
    const unsigned int *mine = pkts + (size_t)blockIdx.y * PKT_WORDS;
    const int ns1 = (2 + D + 6) / 7;                            // sectors of a one-frame command
    unsigned int last_pkt = seq0, last_g = seq0;
    uint32_t dph = 0;
    unsigned long long t_last = globaltimer_ns();
    for (;;) {
      uint4 v = make_uint4(0, 0, 0, 0), g = make_uint4(seq0, 0, 0, 0);
      unsigned int seq = 0, y = 0, z = 0;
      int ns = ns1;
      bool direct = false, quit = false;
      for (;;) {
        if (lane < 2 * ns) v = ld_sys_v4(mine + lane * 4);
        if (lane == 31) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(g.x), "=r"(g.y), "=r"(g.z), "=r"(g.w) : "l"(gcmd) : "memory");
        const unsigned int tag = __shfl_sync(0xffffffffu, v.w, lane | 1);
        const unsigned int t0 = __shfl_sync(0xffffffffu, tag, 1);
        const bool whole = __all_sync(0xffffffffu, lane >= 2 * ns || tag == t0);
        if (t0 != last_pkt && whole) {
          y = __shfl_sync(0xffffffffu, v.x, 0);
          const int need = (y & (CMD_QUIT | CMD_BIG)) ? 1 : (2 + (int)(y & 0xffu) * D + 6) / 7;
          if (need <= ns) { seq = t0; z = __shfl_sync(0xffffffffu, v.y, 0); direct = true; break; }
          ns = need;                                             // a two-frame command: read its other sectors too
          continue;
        }
        const unsigned int gs = __shfl_sync(0xffffffffu, g.x, 31);
        if (gs != last_g) {                                      // relayed by CTA 0
          seq = gs; y = __shfl_sync(0xffffffffu, g.y, 31); z = __shfl_sync(0xffffffffu, g.z, 31);
          __threadfence();                                       // the features were written before the command word
          break;
        }
        const unsigned long long idle = globaltimer_ns() - t_last;
        if (__shfl_sync(0xffffffffu, idle > idle_ns ? 1 : 0, 0)) { quit = true; break; }
        if (idle > 50000ull) __nanosleep(300);                   // a quiet stream: fewer reads over PCIe
      }
      if (!quit && direct) last_pkt = seq;
      if (!quit && !direct) last_g = seq;
      if (!quit && (y & CMD_QUIT)) quit = true;
      const int nf = (int)(y & 0xffu), have = nf * D;
      if (!quit && direct && (y & CMD_BIG)) {
        // CTA 0 only: fetch the features from the shared area, hand them and the command to the grid
        const float *mail_x = reinterpret_cast<const float *>(mailbox) + STREAM_X_BYTE / 4;
        float x[(NF * 64 + 31) / 32];
#pragma unroll
        for (int i = 0; i < (NF * 64 + 31) / 32; i++) if (i * 32 < have) x[i] = (i * 32 + lane < have) ? ld_sys_f32(mail_x + i * 32 + lane) : 0.f;
#pragma unroll
        for (int i = 0; i < (NF * 64 + 31) / 32; i++) if (i * 32 + lane < have) gx[i * 32 + lane] = x[i];
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(gcmd), "r"(seq), "r"(y), "r"(z), "r"(0u) : "memory");
        last_g = seq;
        direct = false;
      } else if (!quit && direct) {
        // payload word p = 7 * sector + j; words 0, 1 = header, word 2 + i = feature i: straight into the staging row
        const int k = lane >> 1, p0 = 7 * k + ((lane & 1) ? 4 : 0) - 2;
        if (lane < 2 * ns) {
          if (p0 >= 0 && p0 < have) stage_x[p0] = __uint_as_float(v.x);
          if (p0 + 1 >= 0 && p0 + 1 < have) stage_x[p0 + 1] = __uint_as_float(v.y);
          if (p0 + 2 >= 0 && p0 + 2 < have) stage_x[p0 + 2] = __uint_as_float(v.z);
          if (!(lane & 1) && p0 + 3 >= 0 && p0 + 3 < have) stage_x[p0 + 3] = __uint_as_float(v.w);
        }
      }
      __syncwarp();
      if (lane == 0) {
        s_cmd.seq = seq;
        s_cmd.nf = nf;
        s_cmd.use_floor = (y & CMD_FLOOR) ? 1 : 0;
        s_cmd.floor_at = __uint_as_float(z);
        s_cmd.quit = quit ? 1 : 0;
        s_cmd.direct = direct ? 1 : 0;
        s_cmd.t_seen = globaltimer_ns();
        mbar_arrive(&call_bar);
      }
      if (quit) break;
      if (lane == 0) mbar_wait(&done_bar, dph);                  // the call has left this CTA: s_cmd, the staging row and A' may change
      dph ^= 1;
      __syncwarp();
      t_last = globaltimer_ns();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                                   // TMEM lane quarter = components [32q, 32q + 32) of the tile
    const int group = (warp - 4) >> 2;                        // tiles i with (i & 1) == group, accumulator set = group
    const int et = threadIdx.x - 128;                         // 0..255 among the epilogue threads
    const int tg = q * 32 + lane;                             // 0..127 within the group
    const int sl = q * 2 + (lane >> 4);                       // this lane's slot within the tile
    const int bar_id = 2 + group;
    uint32_t cph = 0, use = 0, g0 = 0;                          // g0: running number of the call's first tile
    for (;;) {
      mbar_wait(&call_bar, cph);
      cph ^= 1;
      if (s_cmd.quit) break;
      const int nf = s_cmd.nf, use_floor = s_cmd.use_floor;
      const float floor_at = s_cmd.floor_at;
      const unsigned int seq = s_cmd.seq;
      // ===== A' = [Ah | Al] for the NF frame rows: the features of the packet (or the relayed ones, L2), expand, scale, split, store swizzled =====
      {
        const int total = NF * D, have = nf * D;
        const int direct = s_cmd.direct;                         // the mailbox warp has put the features there itself
        for (int idx = et; idx < total; idx += EPI_THREADS)
          if (idx >= have) stage_x[idx] = 0.f;
          else if (!direct) stage_x[idx] = __ldcg(gx + idx);
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        bool ovf = false;
        // (only the rows of the call's frames: the other rows of A' are columns of D nobody reads)
        for (int task = et; task < nf * 2 * NCH; task += EPI_THREADS) {
          const int r = task / (2 * NCH), v = task - r * (2 * NCH);   // vector = 8 consecutive K terms of frame r
          const float *x = stage_x + r * D;
          const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
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
          const int ch = v >> 1, cl = NCH + ch;
          const uint32_t jh = (uint32_t)((ch & 3) * 2 + (v & 1)), jl = (uint32_t)((cl & 3) * 2 + (v & 1));
          unsigned char *ph = smem + (size_t)(ch >> 2) * A_BLOCK + row_off + ((jh ^ (uint32_t)(r & 7)) << 4);
          unsigned char *pl = smem + (size_t)(cl >> 2) * A_BLOCK + row_off + ((jl ^ (uint32_t)(r & 7)) << 4);
          *reinterpret_cast<uint4 *>(ph) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4 *>(pl) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        if (ovf) flags[1] = 1u;                                  // any CTA, same value: a plain store
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full);
      }
      const unsigned long long t_a = globaltimer_ns();
      // ===== epilogue (gmm_stream_kernel's): lane = component; a slot = 16 consecutive lanes =====
      for (int n = n_begin + group; n < n_end; n += 2, use++) {
        const uint32_t g = g0 + (uint32_t)(n - n_begin), set = g % NSETS;
        if (tg < SLOTS) pmeta[group][use & 1][tg] = __ldg(meta + (size_t)n * SLOTS + tg);     // off the critical path
        mbar_wait(&tmem_full[set], (g / NSETS) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * 2 * NF;
        float2(*pt)[NF] = part[group][use & 1];
        const int *pm = pmeta[group][use & 1];
        {
          uint32_t rm[16], rc[16];
          AKU_TMEM_LD16(rm, taddr);
          AKU_TMEM_LD16(rc, taddr + NF);
          AKU_TMEM_LD_WAIT();
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[set]);
#pragma unroll
          for (int c = 0; c < 16; c++) {
            if (c < nf) {                                        // kernel-uniform
              const float v = fmaf(__uint_as_float(rc[c]), LO_INV, __uint_as_float(rm[c]));
              float m = v;
#pragma unroll
              for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
              float s = ex2f((v - m) * LOG2E);
#pragma unroll
              for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
              if ((lane & 15) == c) pt[sl][c] = make_float2(m, s);
            }
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GROUP_THREADS) : "memory");
        // merge the slots of a state, floor, store transposed: thread = (frame, slot)
        for (int e = tg; e < NF * SLOTS; e += GROUP_THREADS) {
          const int f = e >> 3, k = e & 7;
          const int mt = pm[k];
          if (f < nf && mt >= 0 && (mt & 1)) {
            float2 a = pt[k][f];
            if (!(mt & 2)) {
              for (int j = k - 1; j >= 0; j--) {                // a state's slots are consecutive and never leave the tile
                const float2 b = pt[j][f];
                const float M = fmaxf(a.x, b.x);
                a.y = a.y * ex2f((a.x - M) * LOG2E) + b.y * ex2f((b.x - M) * LOG2E);
                a.x = M;
                if (pm[j] & 2) break;
              }
            }
            float res = fmaf(lg2f(a.y), LN2, a.x);
            if (use_floor) res = fmaxf(res, floor_at);           // (float) log(max(likelihood, tiny)); NaN -> floor
            out[(size_t)f * S + (mt >> 2)] = res;
          }
        }
      }
      g0 += (uint32_t)T;
      // ===== completion.  Every CTA has stored its results straight into mapped host memory.  A system-scope fence per
      // CTA costs 4.6 us when 148 execute one together, carrying the rows through device memory with a few copying
      // CTAs 3.7 us (both measured here); a device-scope fence + count per CTA and ONE system-scope fence by the last
      // one is the cheapest order that still puts every row before the sequence number (micro_mailbox2.cu). =====
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (et == 0) {
        const unsigned long long t_epi = globaltimer_ns();
        __threadfence();
        const unsigned int old = atomicAdd(cnt, 1u);
        if (old == gridDim.y - 1) {                              // last CTA: re-arm the counter, publish the sequence number
          *cnt = 0u;
          const unsigned long long t_seen = s_cmd.t_seen, t_all = globaltimer_ns();
          __threadfence_system();
          const unsigned long long t_now = globaltimer_ns();
          flags[4] = (unsigned int)(t_a - t_seen);               // this CTA's view of the call, nanoseconds: command seen -> A' built
          flags[5] = (unsigned int)(t_now - t_seen);             // -> every row fenced
          flags[6] = (unsigned int)(t_epi - t_seen);             // -> this CTA's results stored
          flags[7] = (unsigned int)(t_all - t_seen);             // -> counted in as the last
          flags[0] = seq;
          __threadfence_system();                                // push the words out now: nothing else of this kernel follows them
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&done_bar);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
static size_t resident_slot_bytes(int NCH)
{
  const int nchunk = 2 * NCH;
  return (size_t)(nchunk / 4) * tcs::B_BLOCK + ((nchunk & 2) ? tcs::B_BLOCK / 2 : 0);
}
static size_t resident_fixed_bytes(int NCH, int D)
{
  return 1024 + (size_t)((2 * NCH + 3) / 4) * tcr::A_BLOCK + ((size_t)tcr::NF * D * 4 + 1023) / 1024 * 1024;
}
static int resident_slots(int NCH, int D)
{
  const size_t budget = 227 * 1024 - 7 * 1024 /* static shared memory */;
  const size_t fixed = resident_fixed_bytes(NCH, D), slot = resident_slot_bytes(NCH);
  if (fixed + slot > budget) return 0;
  return (int)std::min<size_t>(tcr::MAX_SLOTS, (budget - fixed) / slot);
}

bool session_applicable(akugpu_ctx *ctx, int precision, int64_t n_frames, const void *feats, const void *out)
{
  if (!ctx->stream_state.session_want) return false;
  if (!stream_applicable(ctx, precision, n_frames)) return false;
  // the two pointer queries cost ~1 us of a ~11 us call: a per-frame loop hands in the same two host buffers every time
  // (under unified addressing a host address never turns into a device address within the life of the process)
  StreamState &st = ctx->stream_state;
  if (feats != st.known_host[0] && feats != st.known_host[1]) {
    if (is_device_ptr(feats)) return false;
    st.known_host[0] = feats;
  }
  if (out != st.known_host[0] && out != st.known_host[1]) {
    if (is_device_ptr(out)) return false;
    st.known_host[1] = out;
  }
  return resident_slots(ctx->ptc16.NCH, ctx->ptc16.D) >= 1;
}

// Writes one command into the packets of CTAs [0, n_ctas): payload words {y, z, x[0 .. nx)} in sectors of 7 words + tag;
// a sector's tag (= seq) is stored after its payload, so the device may trust any sector whose tag is new.  At least
// `min_sectors` sectors are written (the kernel always reads the sectors of a one-frame command).
static void session_send(StreamState &st, int n_ctas, unsigned int seq, unsigned int y, unsigned int z, const float *x, int nx, int min_sectors)
{
  alignas(64) unsigned int tmpl[tcr::PKT_WORDS];
  const int ns = std::max(min_sectors, (2 + nx + 6) / 7);
  for (int k = 0; k < ns; k++) {
    for (int j = 0; j < 7; j++) {
      const int p = 7 * k + j;
      unsigned int w = 0;
      if (p == 0) w = y;
      else if (p == 1) w = z;
      else if (p - 2 < nx) memcpy(&w, x + (p - 2), 4);
      tmpl[8 * k + j] = w;
    }
    tmpl[8 * k + 7] = seq;
  }
  unsigned int *base = reinterpret_cast<unsigned int *>(st.pkt_host);
#if defined(__x86_64__)
  _mm_sfence();       // whatever the caller stored before (the features of a relayed call) is ordered before the packets
#endif
  // CTAs with the most component tiles first: the last packet written is on the critical path of the call
  const bool ordered = (int)st.session_order.size() == n_ctas;
  for (int i = 0; i < n_ctas; i++) {
    unsigned int *pk = base + (size_t)(ordered ? st.session_order[i] : i) * tcr::PKT_WORDS;
#if defined(__x86_64__)
    // non-temporal 16-byte stores in address order: a line leaves the write-combining buffer whole (or as a prefix of what
    // was written), so a sector's tag never precedes its payload, and the host does not fight the polling device for
    // ownership of 444 cache lines (2.2 -> 1.7 us for 148 packets, scripts/micro_mailbox2.cu)
    for (int c = 0; c < 2 * ns; c++) _mm_stream_si128(reinterpret_cast<__m128i *>(pk + 4 * c), _mm_load_si128(reinterpret_cast<const __m128i *>(tmpl + 4 * c)));
#else
    for (int k = 0; k < ns; k++) {
      memcpy(pk + 8 * k, tmpl + 8 * k, 28);
      std::atomic_thread_fence(std::memory_order_release);
      reinterpret_cast<volatile unsigned int *>(pk)[8 * k + 7] = seq;
    }
#endif
  }
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

// One resident kernel per device: a second one could not become resident beside the first (each takes all of the SMs'
// shared memory) and the two would hand the device back and forth at the pace of their idle timers.
static std::mutex g_session_mu;
static std::map<int, akugpu_ctx *> g_session_owner;
static void session_claim_device(akugpu_ctx *ctx)
{
  std::lock_guard<std::mutex> lock(g_session_mu);
  auto it = g_session_owner.find(ctx->device);
  if (it != g_session_owner.end() && it->second != ctx)
    throw Error(AKUGPU_E_STATE, "another context of this process holds the resident scorer of this device (akugpu_stream_close it first)");
  g_session_owner[ctx->device] = ctx;
}
void session_release_device(akugpu_ctx *ctx)
{
  std::lock_guard<std::mutex> lock(g_session_mu);
  auto it = g_session_owner.find(ctx->device);
  if (it != g_session_owner.end() && it->second == ctx) g_session_owner.erase(it);
}

void session_launch(akugpu_ctx *ctx)
{
  PackedTC16 &p = ctx->ptc16;
  StreamState &st = ctx->stream_state;
  const int S = ctx->hm.S;
  session_claim_device(ctx);
  stream_buffers(ctx, S);
  stream_map_ready(ctx);
  if (!st.session_stream) AKU_CUDA(cudaStreamCreateWithFlags(&st.session_stream, cudaStreamNonBlocking));
  int ysplit = 1;
  const int *ranges = tc_tile_ranges(ctx, p.n_tiles, p.clean, p.ranges, std::min(p.n_tiles, ctx->sm_count), ysplit);
  const int nslots = resident_slots(p.NCH, p.D);
  const size_t smem = resident_fixed_bytes(p.NCH, p.D) + (size_t)nslots * resident_slot_bytes(p.NCH);
  unsigned char *db = reinterpret_cast<unsigned char *>(st.dev_view);
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));      // model uploads, range tables
  if (!st.relay.p) st.relay.reserve(64 + (size_t)tcr::NF * 64 * sizeof(float));
  if (ysplit > tcr::MAX_GRID) throw Error(AKUGPU_E_STATE, "gmm_resident: more CTAs than packets");
  if (!st.pkt_host) {
    AKU_CUDA(cudaHostAlloc(&st.pkt_host, (size_t)tcr::MAX_GRID * tcr::PKT_WORDS * 4, cudaHostAllocMapped | cudaHostAllocPortable));
    AKU_CUDA(cudaHostGetDevicePointer(&st.pkt_dev, st.pkt_host, 0));
  }
  {   // every sector's tag = the sequence number the kernel starts from: nothing is new
    unsigned int *pk = reinterpret_cast<unsigned int *>(st.pkt_host);
    for (int i = 0; i < tcr::MAX_GRID * tcr::PKT_WORDS; i++) pk[i] = (i & 7) == 7 ? st.seq : 0u;
  }
  st.session_grid = ysplit;
  {   // the order in which a call's packets are written: CTAs by falling number of component tiles
    std::vector<int> rb((size_t)ysplit + 1);
    AKU_CUDA(cudaMemcpy(rb.data(), ranges, rb.size() * sizeof(int), cudaMemcpyDeviceToHost));
    st.session_order.resize((size_t)ysplit);
    for (int i = 0; i < ysplit; i++) st.session_order[(size_t)i] = i;
    std::stable_sort(st.session_order.begin(), st.session_order.end(), [&](int a, int b) { return rb[a + 1] - rb[a] > rb[b + 1] - rb[b]; });
  }
  AKU_CUDA(cudaMemsetAsync(st.cnt.p, 0, 16, st.session_stream));
  const unsigned int cmd0[4] = {st.seq, 0u, 0u, 0u};           // the relayed command word the CTAs wait to see CHANGE
  AKU_CUDA(cudaMemcpyAsync(st.relay.p, cmd0, 16, cudaMemcpyHostToDevice, st.session_stream));
  const unsigned long long idle_ns = (unsigned long long)(std::max(1.0, st.session_idle_ms) * 1e6);
  float *d_out = reinterpret_cast<float *>(db + STREAM_X_BYTE) + STREAM_MAX_FRAMES * 64;
  auto launch = [&](auto kernel) {
    ensure_dynamic_smem(ctx, (const void *)kernel, smem);
    kernel<<<dim3(1, ysplit), tcs::THREADS, smem, st.session_stream>>>(
        *reinterpret_cast<const CUtensorMap *>(p.map_b), *reinterpret_cast<const CUtensorMap *>(p.map_b64), nslots, ranges, p.meta.as<int>(),
        p.D, S, p.escale.as<float>(), reinterpret_cast<const unsigned int *>(db), reinterpret_cast<const unsigned int *>(st.pkt_dev),
        reinterpret_cast<float *>(st.relay.as<unsigned char>() + 64), st.relay.as<unsigned int>(), d_out, st.cnt.as<unsigned int>(),
        reinterpret_cast<volatile unsigned int *>(db), st.seq, idle_ns);
  };
  switch (p.NCH) {
    case 1: launch(gmm_resident_kernel<1>); break;
    case 2: launch(gmm_resident_kernel<2>); break;
    case 3: launch(gmm_resident_kernel<3>); break;
    case 4: launch(gmm_resident_kernel<4>); break;
    case 5: launch(gmm_resident_kernel<5>); break;
    case 6: launch(gmm_resident_kernel<6>); break;
    case 7: launch(gmm_resident_kernel<7>); break;
    case 8: launch(gmm_resident_kernel<8>); break;
    default: throw Error(AKUGPU_E_STATE, "gmm_resident: unsupported feature dimension");
  }
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
  st.session_launches++;
  st.session_live = true;
  st.session_last = std::chrono::steady_clock::now();
}

// Ends the resident kernel (if one is running) so that the device is free for whatever comes next.
void session_quiesce(akugpu_ctx *ctx)
{
  StreamState &st = ctx->stream_state;
  if (!st.session_live) return;
  const unsigned int seq = ++st.seq ? st.seq : ++st.seq;
  session_send(st, st.session_grid, seq, tcr::CMD_QUIT, 0u, nullptr, 0, (2 + ctx->ptc16.D + 6) / 7);
  st.session_live = false;
  AKU_CUDA(cudaStreamSynchronize(st.session_stream));
}

const float *session_rows(akugpu_ctx *ctx)
{
  unsigned char *hb = reinterpret_cast<unsigned char *>(ctx->stream_state.host);
  return hb ? reinterpret_cast<const float *>(hb + STREAM_X_BYTE) + STREAM_MAX_FRAMES * 64 : nullptr;
}

void session_destroy(akugpu_ctx *ctx)
{
  StreamState &st = ctx->stream_state;
  try { session_quiesce(ctx); } catch (...) {}
  session_release_device(ctx);
  if (st.session_stream) { cudaStreamDestroy(st.session_stream); st.session_stream = nullptr; }
  if (st.pkt_host) { cudaFreeHost(st.pkt_host); st.pkt_host = nullptr; st.pkt_dev = nullptr; }
}

// One call through the resident kernel: host features in, host log-likelihoods out.  false = a feature left the fp16
// range of the scaled terms (the caller quiesces and takes the general path).
bool session_score(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, float *out, int use_floor, float floor_at)
{
  PackedTC16 &p = ctx->ptc16;
  StreamState &st = ctx->stream_state;
  const int S = ctx->hm.S, D = p.D, nf = (int)n_frames;
  const auto t_call = std::chrono::steady_clock::now();
  if (st.session_live && std::chrono::duration<double, std::milli>(t_call - st.session_last).count() > 0.5 * st.session_idle_ms &&
      cudaStreamQuery(st.session_stream) == cudaSuccess)
    st.session_live = false;                                    // it has ended on its idle timer
  if (!st.session_live) session_launch(ctx);
  unsigned char *hb = reinterpret_cast<unsigned char *>(st.host);
  volatile unsigned int *hflags = reinterpret_cast<volatile unsigned int *>(hb);
  float *h_x = reinterpret_cast<float *>(hb + STREAM_X_BYTE), *h_out = h_x + STREAM_MAX_FRAMES * 64;
  const int ns1 = (2 + D + 6) / 7;
  const bool direct = 2 + nf * D <= 7 * tcr::PKT_SECTORS;      // the whole command fits a CTA's packet
  // centre on the host exactly as the kernels do: (float)((double) f - c)
  const double *c = p.h_center.data();
  if (feats_f64) { const double *f = (const double *)feats; for (int i = 0; i < nf; i++) for (int d = 0; d < D; d++) h_x[i * D + d] = (float)(f[i * D + d] - c[d]); }
  else { const float *f = (const float *)feats; for (int i = 0; i < nf; i++) for (int d = 0; d < D; d++) h_x[i * D + d] = (float)((double)f[i * D + d] - c[d]); }
  unsigned int floor_bits;
  memcpy(&floor_bits, &floor_at, 4);
  for (int attempt = 0;; attempt++) {
    const unsigned int seq = ++st.seq ? st.seq : ++st.seq;      // never 0
    const unsigned int y = (unsigned int)nf | (use_floor ? tcr::CMD_FLOOR : 0u);
    if (direct) session_send(st, st.session_grid, seq, y, floor_bits, h_x, nf * D, ns1);
    else session_send(st, 1, seq, y | tcr::CMD_BIG, floor_bits, nullptr, 0, ns1);   // features in the shared area: CTA 0 relays
    uint64_t spins = 0;
    bool lost = false;
    const auto t0 = std::chrono::steady_clock::now();
    while (hflags[0] != seq) {
#if defined(__x86_64__)
      _mm_pause();
#endif
      if (++spins == 16384 || (spins & 0x3FFFF) == 0) {           // is the kernel still there?
        cudaError_t e = cudaStreamQuery(st.session_stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) { st.session_live = false; throw Error(AKUGPU_E_CUDA, std::string("gmm_resident_kernel: ") + cudaGetErrorString(e)); }
        if (e == cudaSuccess && hflags[0] != seq) { lost = true; break; }       // it ended on its idle timer around this call
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 10.0) {
          st.session_live = false;
          throw Error(AKUGPU_E_CUDA, "gmm_resident_kernel does not answer (10 s)");
        }
      }
    }
    if (!lost) break;
    st.session_live = false;
    if (attempt >= 2) throw Error(AKUGPU_E_CUDA, "gmm_resident_kernel ended three times under one call");
    session_launch(ctx);
  }
  st.session_calls++;
  st.session_last = std::chrono::steady_clock::now();
  if (hflags[1]) { hflags[1] = 0; return false; }
  if (out) memcpy(out, h_out, (size_t)nf * S * sizeof(float));        // out == NULL: the caller reads the pinned rows themselves
  return true;
}

}  // namespace akugpu
