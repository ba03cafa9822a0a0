// gmm_stream.cu -- streaming-regime scorer: a handful of frames against the WHOLE model in one launch.
//
// The decoder's in-process feed scores ONE frame per call (decoder/decode-stream.cc:191-207 -> Toolbox::set_one_frame ->
// OneFrameAcoustics::set, decoder/src/OneFrameAcoustics.cc:23-30; in the reference HmmSet::precompute_likelihoods,
// aku/HmmSet.cc:485-501, then state_likelihood for every state).  In this regime the work is a sweep of the parameter
// image (config-2 model: 30 MB of fp16 hi/lo-split expanded parameters, L2-resident from the second call on), not
// arithmetic: SURVEY.md section 8(d) "streaming regime".  The batch kernel (gmm_tc16.cu) spends its time per call on
// several launches, an fp16-overflow flag read-back and two copies: 41 us whatever F <= 128.
//
// Here: ONE kernel, the component tiles spread over all SMs (grid.y), and the GEMM turned around --
//     D[component (M = 128), frame (N = 16)] = B'[128 x K] . A'[frames x K]^T
// so the tensor-core time per tile is an eighth of the batch kernel's M = 128 frames (a tile then costs what its 48 KB
// of parameters cost to fetch).  The epilogue reads TMEM with component = lane: the mixture log-sum-exp of a slot is a
// 16-lane shuffle reduction; multi-slot states are merged through a tiny shared-memory table; the floor of the decoder
// feed (log(max(l, tiny))) and the [frame][state] transpose are fused into the store.
// Host-facing latency: features of a small call travel as kernel parameters (no H2D copy), results are stored
// straight into pinned, mapped host memory (no D2H copy), and the last CTA to finish publishes a sequence number there,
// together with the fp16-range flag, which the host polls -- no stream synchronise, no flag read-back.
#include "ctx.hpp"
#include "kernels.hpp"
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "stream_common.cuh"
#include <math.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace akugpu {


template <int NF> struct StreamX { float v[NF * tcs::XS_DIM]; };      // 2.5 KB (NF = 16) / 5 KB (NF = 32: large-parameter launch)

// feats == nullptr: the frames are xs.v (already centred, [nf][D]); otherwise raw features in global / mapped memory,
// centred here (center != nullptr) or taken as they are (center == nullptr: centred by the host into mapped memory).
template <int NCH, int NF>
__global__ void __launch_bounds__(tcs::THREADS, 1)
gmm_stream_kernel(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ StreamX<NF> xs, int tslots,
                  const int *__restrict__ range_begin, const int *__restrict__ meta, const void *__restrict__ feats, int feats_f64,
                  int nf, int D, int S, const double *__restrict__ center, const float *__restrict__ escale,
                  float *__restrict__ out, int use_floor, float floor_at, unsigned int *__restrict__ cnt,
                  volatile unsigned int *__restrict__ flags, unsigned int seq)
{
  using namespace tc;
  using namespace tcs;
  constexpr int KB = (2 * NCH + 3) / 4;                       // 64-wide k-blocks of a row of A' / B'
  constexpr uint32_t A_BLOCK = NF * BK * 2;                   // one k-block of A' (NF frame rows)
  constexpr uint32_t SLOT_BYTES = KB * B_BLOCK;
  constexpr uint32_t IDESC = umma_idesc(BM, NF, false);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[MAX_TSLOTS], empty_bar[MAX_TSLOTS], tmem_full[2], tmem_empty[2], a_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float2 part[2][2][SLOTS][NF];                    // [group][tile parity][slot][frame] = {max, sum of exp}
  __shared__ int pmeta[2][2][SLOTS];                          // the tile's slot table, fetched while its MMAs run
  unsigned char *ring = smem + KB * A_BLOCK;
  float *stage_x = reinterpret_cast<float *>(ring + (size_t)tslots * SLOT_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_begin = range_begin[blockIdx.y], n_end = range_begin[blockIdx.y + 1];

  if (threadIdx.x == 0) {
    for (int s = 0; s < tslots; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    mbar_init(&a_full, EPI_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // two accumulator sets of MAIN [0, NF) + CORR [NF, 2 NF)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(4 * NF) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: one component tile of B' per ring slot; the first tslots tiles are in flight before A' exists =====
      int slot = 0;
      uint32_t ph = 0;
      for (int n = n_begin; n < n_end; n++) {
        mbar_wait(&empty_bar[slot], ph ^ 1);
        mbar_expect_tx(&full_bar[slot], SLOT_BYTES);
        unsigned char *dst = ring + (size_t)slot * SLOT_BYTES;
#pragma unroll
        for (int kb = 0; kb < KB; kb++) tma_load_2d(dst + kb * B_BLOCK, &mapB, kb * BK, n * BM, &full_bar[slot]);
        if (++slot == tslots) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer: the component tile is the M-side operand, the frames the N-side =====
      auto boff = [](int c) -> uint32_t { return ((uint32_t)(c >> 2) * B_BLOCK + (uint32_t)(c & 3) * 32u) >> 4; };
      auto aoff = [](int c) -> uint32_t { return ((uint32_t)(c >> 2) * A_BLOCK + (uint32_t)(c & 3) * 32u) >> 4; };
      const uint64_t x_desc0 = umma_desc(smem_u32(smem));
      int slot = 0;
      uint32_t ph = 0;
      mbar_wait(&a_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int n = n_begin; n < n_end; n++) {
        const int i = n - n_begin, a = i & 1;
        mbar_wait(&tmem_empty[a], ((i >> 1) & 1) ^ 1);
        mbar_wait(&full_bar[slot], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_main = tmem_base + a * 2 * NF, d_corr = d_main + NF;
        const uint64_t p_desc0 = umma_desc(smem_u32(ring) + (uint32_t)slot * SLOT_BYTES);
#pragma unroll
        for (int j = 0; j < NCH; j++) umma_f16(d_corr, p_desc0 + boff(NCH + j), x_desc0 + aoff(j), IDESC, j > 0 ? 1u : 0u);   // Bl . Ah
#pragma unroll
        for (int j = 0; j < NCH; j++) umma_f16(d_corr, p_desc0 + boff(j), x_desc0 + aoff(NCH + j), IDESC, 1u);                // Bh . Al
#pragma unroll
        for (int j = 0; j < NCH; j++) umma_f16(d_main, p_desc0 + boff(j), x_desc0 + aoff(j), IDESC, j > 0 ? 1u : 0u);         // Bh . Ah
        umma_commit(&empty_bar[slot]);
        umma_commit(&tmem_full[a]);
        if (++slot == tslots) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                                   // TMEM lane quarter = components [32q, 32q + 32) of the tile
    const int group = (warp - 4) >> 2;                        // tiles i with (i & 1) == group, accumulator set = group
    const int et = threadIdx.x - 128;                         // 0..255 among the epilogue threads
    const int tg = q * 32 + lane;                             // 0..127 within the group
    // ===== A' = [Ah | Al] for the NF frame rows: centre, expand, scale, split, store swizzled =====
    {
      const int total = NF * D;
      for (int idx = et; idx < total; idx += EPI_THREADS) {
        const int r = idx / D, d = idx - r * D;
        float v = 0.f;
        if (r < nf) {
          if (!feats) v = xs.v[idx];
          else {
            const double t = feats_f64 ? reinterpret_cast<const double *>(feats)[idx] : (double)reinterpret_cast<const float *>(feats)[idx];
            v = center ? (float)(t - center[d]) : (float)t;
          }
        }
        stage_x[idx] = v;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      bool ovf = false;
      for (int task = et; task < NF * 2 * NCH; task += EPI_THREADS) {
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
    // ===== epilogue: lane = component; a slot = 16 consecutive lanes =====
    const int sl = q * 2 + (lane >> 4);                       // this lane's slot within the tile
    const int bar_id = 2 + group;
    for (int n = n_begin + group; n < n_end; n += 2) {
      const int use = (n - n_begin) >> 1;
      if (tg < SLOTS) pmeta[group][use & 1][tg] = __ldg(meta + (size_t)n * SLOTS + tg);     // off the critical path
      mbar_wait(&tmem_full[group], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + group * 2 * NF;
      float2(*pt)[NF] = part[group][use & 1];
      const int *pm = pmeta[group][use & 1];
#pragma unroll
      for (int h = 0; h < NF / 16; h++) {
        uint32_t rm[16], rc[16];
        AKU_TMEM_LD16(rm, taddr + h * 16);
        AKU_TMEM_LD16(rc, taddr + NF + h * 16);
        AKU_TMEM_LD_WAIT();
        if (h == NF / 16 - 1) {   // every TMEM read of this warp is complete: hand the accumulator set back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[group]);
        }
#pragma unroll
        for (int c = 0; c < 16; c++) {
          const int f = h * 16 + c;
          if (f < nf) {                                        // kernel-uniform
            const float v = fmaf(__uint_as_float(rc[c]), LO_INV, __uint_as_float(rm[c]));
            float m = v;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s = ex2f((v - m) * LOG2E);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((lane & 15) == (c & 15)) pt[sl][f] = make_float2(m, s);
          }
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GROUP_THREADS) : "memory");
      // merge the slots of a state, floor, store transposed: thread = (frame, slot), eight slots = eight neighbouring
      // states = one 32-byte store per frame
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
      // the table of the other parity is written next; this one again only after the next barrier
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(4 * NF) : "memory");
  }
  if (threadIdx.x == 0) {
    // ONE system-scope fence per CTA: the barrier above orders every thread's result stores (possibly into mapped host
    // memory) before it, and the fence is cumulative (a fence per writing thread cost more than the whole sweep)
    __threadfence_system();
    const unsigned int old = atomicAdd(cnt, 1u);
    if (old == gridDim.x * gridDim.y - 1) {                    // last CTA: re-arm the counter, publish the sequence number
      *cnt = 0u;
      __threadfence_system();
      flags[0] = seq;
      __threadfence_system();
    }
  }
}

// Plain read sweep of a buffer by every SM (uint4 loads): the L2 / HBM read rate the parameter sweep is measured against.
__global__ void __launch_bounds__(512)
sweep_probe_kernel(const uint4 *__restrict__ p, size_t n, unsigned int *__restrict__ sink)
{
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(p + i);
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) *sink = 1u;     // never true in practice; keeps the loads
}
// SM clock in MHz right now: cycles of one SM against the global nanosecond timer over a ~20 us spin.
__global__ void sm_clock_kernel(double *mhz)
{
  unsigned long long t0, t1, g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  t0 = clock64();
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); } while (g1 - g0 < 20000ull);
  t1 = clock64();
  *mhz = (double)(t1 - t0) / (double)(g1 - g0) * 1e3;
}
__global__ void flush_fill_kernel(uint4 *__restrict__ p, size_t n, unsigned int v)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(v, v, v, v);
}

// ------------------------------------------------------------------------------------------------
// host side
static int stream_tslots(int KB, int NF, int D)
{
  const size_t budget = 227 * 1024 - 12 * 1024 /* static shared memory, alignment */;
  const size_t fixed = (size_t)KB * NF * tcs::BK * 2 + ((size_t)NF * D * 4 + 1023) / 1024 * 1024;
  const size_t slot = (size_t)KB * tcs::B_BLOCK;
  if (fixed + slot > budget) return 0;
  return (int)std::min<size_t>(tcs::MAX_TSLOTS, (budget - fixed) / slot);
}

void stream_buffers(akugpu_ctx *ctx, int S)
{
  StreamState &st = ctx->stream_state;
  const size_t need = 256 + (size_t)STREAM_MAX_FRAMES * 64 * sizeof(float) + (size_t)STREAM_MAX_FRAMES * S * sizeof(float);
  if (st.host && st.bytes >= need) return;
  if (st.host) { cudaFreeHost(st.host); st.host = nullptr; }
  AKU_CUDA(cudaHostAlloc(&st.host, need, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(st.host, 0, 256);
  AKU_CUDA(cudaHostGetDevicePointer(&st.dev_view, st.host, 0));
  st.bytes = need;
  if (!st.cnt.p) { st.cnt.reserve(16); AKU_CUDA(cudaMemsetAsync(st.cnt.p, 0, 16, ctx->stream)); AKU_CUDA(cudaStreamSynchronize(ctx->stream)); }
}

void stream_map_ready(akugpu_ctx *ctx)
{
  PackedTC16 &p = ctx->ptc16;
  if (p.map_ready) return;
  tc_make_map(reinterpret_cast<CUtensorMap *>(p.map_b), p.B.p, (uint64_t)p.n_tiles * tcs::BM, (uint64_t)p.Kp, true);
  // the last, half-filled k-block of a row (NCH odd) as a 32-column box with the 64-byte swizzle (gmm_resident.cu)
  tc_make_map(reinterpret_cast<CUtensorMap *>(p.map_b64), p.B.p, (uint64_t)p.n_tiles * tcs::BM, (uint64_t)p.Kp, true, 32);
  p.map_ready = true;
}

bool stream_applicable(akugpu_ctx *ctx, int precision, int64_t n_frames)
{
  const PackedTC16 &p = ctx->ptc16;
  if (!ctx->streaming_enabled || precision != AKUGPU_F32 || n_frames < 1 || n_frames > STREAM_MAX_FRAMES) return false;
  if (!p.ready || p.stream || p.hybrid || ctx->tc16_suspended || ctx->hm.cmllr_on) return false;
  if (ctx->hm.use_clustering && ctx->hm.n_clusters > 0) return false;
  if (p.NCH < 1 || p.NCH > 8 || p.D > 63) return false;
  if (n_frames > 8 && p.n_tiles > 800) return false;      // the result stores (32 bytes per tile and frame over PCIe) outweigh the sweep
  return stream_tslots(p.KB, 16, p.D) >= 1;
}

template <int NF>
static void stream_launch(akugpu_ctx *ctx, const StreamX<NF> &xs, const void *feats, int feats_f64, int nf, const double *center, float *out,
                          int use_floor, float floor_at, unsigned int seq, const int *ranges, int ysplit, int tslots)
{
  PackedTC16 &p = ctx->ptc16;
  StreamState &st = ctx->stream_state;
  const size_t smem = 1024 + (size_t)p.KB * NF * tcs::BK * 2 + (size_t)tslots * p.KB * tcs::B_BLOCK + ((size_t)NF * p.D * 4 + 1023) / 1024 * 1024;
  const CUtensorMap &map = *reinterpret_cast<const CUtensorMap *>(p.map_b);
  unsigned int *flags = reinterpret_cast<unsigned int *>(st.dev_view);
  auto launch = [&](auto kernel) {
    ensure_dynamic_smem(ctx, (const void *)kernel, smem);
    kernel<<<dim3(1, ysplit), tcs::THREADS, smem, ctx->stream>>>(map, xs, tslots, ranges, p.meta.as<int>(), feats, feats_f64, nf, p.D, ctx->hm.S,
                                                                center, p.escale.as<float>(), out, use_floor, floor_at, st.cnt.as<unsigned int>(),
                                                                flags, seq);
  };
  switch (p.NCH) {
    case 1: launch(gmm_stream_kernel<1, NF>); break;
    case 2: launch(gmm_stream_kernel<2, NF>); break;
    case 3: launch(gmm_stream_kernel<3, NF>); break;
    case 4: launch(gmm_stream_kernel<4, NF>); break;
    case 5: launch(gmm_stream_kernel<5, NF>); break;
    case 6: launch(gmm_stream_kernel<6, NF>); break;
    case 7: launch(gmm_stream_kernel<7, NF>); break;
    case 8: launch(gmm_stream_kernel<8, NF>); break;
    default: throw Error(AKUGPU_E_STATE, "gmm_stream: unsupported feature dimension");
  }
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

// Scores n_frames <= STREAM_MAX_FRAMES frames into out [n_frames x S] (float natural-log state likelihoods, floored at
// floor_at when use_floor).  feats / out: host or device.  Returns false when a feature left the fp16 range of the
// scaled terms (the caller then takes the general path, which redoes the call with the bf16x3 kernel).
static bool stream_score_impl(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, float *out, int use_floor, float floor_at,
                              bool wait);
bool stream_score(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, float *out, int use_floor, float floor_at)
{
  return stream_score_impl(ctx, feats, feats_f64, n_frames, out, use_floor, floor_at, true);
}
static bool stream_score_impl(akugpu_ctx *ctx, const void *feats, int feats_f64, int64_t n_frames, float *out, int use_floor, float floor_at,
                              bool wait)
{
  PackedTC16 &p = ctx->ptc16;
  StreamState &st = ctx->stream_state;
  const int S = ctx->hm.S, D = p.D, nf = (int)n_frames;
  stream_buffers(ctx, S);
  stream_map_ready(ctx);
  int ysplit = 1;
  const int *ranges = tc_tile_ranges(ctx, p.n_tiles, p.clean, p.ranges, std::min(p.n_tiles, ctx->sm_count), ysplit);
  constexpr int NF = 16;
  const int tslots = stream_tslots(p.KB, NF, D);
  unsigned char *hb = reinterpret_cast<unsigned char *>(st.host), *db = reinterpret_cast<unsigned char *>(st.dev_view);
  volatile unsigned int *hflags = reinterpret_cast<volatile unsigned int *>(hb);
  float *h_x = reinterpret_cast<float *>(hb + 256), *h_out = h_x + STREAM_MAX_FRAMES * 64;
  const float *d_x = reinterpret_cast<const float *>(db + 256);
  float *d_out = reinterpret_cast<float *>(db + 256) + STREAM_MAX_FRAMES * 64;

  const bool feats_dev = is_device_ptr(feats), out_dev = is_device_ptr(out);
  StreamX<16> xs16;
  float *xs_v = xs16.v;
  const void *k_feats = nullptr;
  int k_f64 = 0;
  const double *k_center = nullptr;
  if (feats_dev) { k_feats = feats; k_f64 = feats_f64; k_center = p.center.as<double>(); }
  else {
    // centre on the host exactly as the kernels do: (float)((double) f - c)
    float *dst = (D <= tcs::XS_DIM) ? xs_v : h_x;
    const double *c = p.h_center.data();
    if (feats_f64) { const double *f = (const double *)feats; for (int i = 0; i < nf; i++) for (int d = 0; d < D; d++) dst[i * D + d] = (float)(f[i * D + d] - c[d]); }
    else { const float *f = (const float *)feats; for (int i = 0; i < nf; i++) for (int d = 0; d < D; d++) dst[i * D + d] = (float)((double)f[i * D + d] - c[d]); }
    if (dst == h_x) k_feats = d_x;      // too wide for the parameter block: the kernel reads the mapped staging area
  }
  float *k_out = out_dev ? out : d_out;
  const unsigned int seq = ++st.seq ? st.seq : ++st.seq;      // never 0
  {
    StageScope sc(ctx, 1);
    stream_launch<16>(ctx, xs16, k_feats, k_f64, nf, k_center, k_out, use_floor, floor_at, seq, ranges, ysplit, tslots);
  }
  if (!wait) return true;                      // (probe: a train of launches, completion by event)
  // completion: the last CTA stores seq into mapped host memory after every result is visible
  uint64_t spins = 0;
  while (hflags[0] != seq) {
#if defined(__x86_64__)
    _mm_pause();
#endif
    if ((++spins & 0xFFFFF) == 0) {           // every ~1M polls: has the launch failed?
      cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) throw Error(AKUGPU_E_CUDA, std::string("gmm_stream_kernel: ") + cudaGetErrorString(e));
      if (e == cudaSuccess && hflags[0] != seq) throw Error(AKUGPU_E_CUDA, "gmm_stream_kernel finished without publishing its sequence number");
    }
  }
  if (hflags[1]) { hflags[1] = 0; return false; }
  if (!out_dev) memcpy(out, h_out, (size_t)nf * S * sizeof(float));
  return true;
}

// Rates the streaming regime is reported against (akugpu_stream_probe): out[0] = bytes of the parameter image,
// out[1] / out[2] = seconds per stream-kernel launch with the image L2-resident / after an L2 flush (CUDA events around
// the kernel only, one frame), out[3] / out[4] = seconds of a plain read sweep of the image by all SMs, L2-resident /
// after a flush; out[5] = SM clock (MHz) measured right after those isolated launches (an idle GPU clocks down between
// them); out[6] = seconds per launch in a train of 200 back-to-back launches (GPU busy, image L2-resident), out[7] = SM
// clock (MHz) at the end of that train.
void stream_probe(akugpu_ctx *ctx, double out[8])
{
  PackedTC16 &p = ctx->ptc16;
  for (int i = 0; i < 8; i++) out[i] = 0;
  if (!stream_applicable(ctx, AKUGPU_F32, 1)) throw Error(AKUGPU_E_STATE, "stream_probe: the streaming scorer does not serve the loaded model");
  const int S = ctx->hm.S, D = p.D;
  const size_t img = (size_t)p.n_tiles * tcs::BM * p.Kp * 2;
  out[0] = (double)img;
  std::vector<float> x(D, 0.f), res(S);
  for (int d = 0; d < D; d++) x[d] = (float)p.h_center[d];
  DevBuf flush, sink, dres;
  const size_t flush_bytes = (size_t)512 << 20;
  flush.reserve(flush_bytes);
  sink.reserve(16);
  dres.reserve((size_t)S * 4);
  cudaEvent_t e0, e1;
  AKU_CUDA(cudaEventCreate(&e0));
  AKU_CUDA(cudaEventCreate(&e1));
  auto time_it = [&](bool flush_first, bool kernel, int reps) {
    double best = 1e30, sum = 0;
    for (int r = 0; r < reps; r++) {
      if (flush_first) flush_fill_kernel<<<ctx->sm_count * 8, 512, 0, ctx->stream>>>(flush.as<uint4>(), flush_bytes / 16, (unsigned)r);
      AKU_CUDA(cudaEventRecord(e0, ctx->stream));
      if (kernel) {
        // device-resident output so that the probe times the sweep, not PCIe
        stream_score(ctx, x.data(), 0, 1, dres.as<float>(), 0, 0.f);
      } else {
        sweep_probe_kernel<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>(p.B.as<uint4>(), img / 16, sink.as<unsigned int>());
      }
      AKU_CUDA(cudaEventRecord(e1, ctx->stream));
      AKU_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      AKU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      best = std::min(best, (double)ms);
      sum += ms;
    }
    (void)sum;
    return best * 1e-3;
  };
  time_it(false, true, 3);
  out[1] = time_it(false, true, 20);
  out[2] = time_it(true, true, 10);
  DevBuf mhz;
  mhz.reserve(16);
  auto clock_now = [&]() {
    double h = 0;
    sm_clock_kernel<<<1, 1, 0, ctx->stream>>>(mhz.as<double>());
    AKU_CUDA(cudaMemcpyAsync(&h, mhz.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    AKU_CUDA(cudaStreamSynchronize(ctx->stream));
    return h;
  };
  out[5] = clock_now();
  time_it(false, false, 3);
  out[3] = time_it(false, false, 20);
  out[4] = time_it(true, false, 10);
  {   // a train of launches: the GPU stays busy and at its working clock
    const int train = 200;
    for (int w = 0; w < 2; w++) {
      AKU_CUDA(cudaEventRecord(e0, ctx->stream));
      for (int r = 0; r < train; r++) stream_score_impl(ctx, x.data(), 0, 1, dres.as<float>(), 0, 0.f, false);
      AKU_CUDA(cudaEventRecord(e1, ctx->stream));
      AKU_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      AKU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      out[6] = (double)ms * 1e-3 / train;
    }
    out[7] = clock_now();
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

}  // namespace akugpu
